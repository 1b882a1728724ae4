"""Data-parallel engine logic over gloo on CPU (world size 2): bucket layout, boundary-only communication, gradient
averaging == mean of per-rank gradients, lock-step with parameters that got no gradient, checkpoint round trip.
The engine is model-agnostic, so a small torch module stands in for the GPU-only DB1 module here; the same code path
runs on the B200s with NCCL (bench.py --gpus N)."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.word_embedding = nn.Embedding(11, 8)
        self.h = nn.ModuleList([nn.Linear(8, 8) for _ in range(3)])
        self.unused = nn.Linear(8, 8)  # e.g. the vision encoder on a text-only rank

    def forward(self, tok, use_unused=False):
        x = self.word_embedding(tok)
        for l in self.h:
            x = torch.tanh(l(x))
        if use_unused:
            x = x + self.unused(x)
        return None, x.pow(2).mean()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, ga):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import src.mpu as mpu
        from db1_sm100.engine import DB1Engine
        mpu.initialize_model_parallel()
        assert mpu.get_data_parallel_world_size() == world and mpu.get_tensor_model_parallel_world_size() == 1
        torch.manual_seed(0)
        model = _Toy()
        ref = _Toy()
        ref.load_state_dict(model.state_dict())
        eng = DB1Engine(model, mpu=mpu, gradient_accumulation_steps=ga, loss_scale=8.0)
        assert [b.key for b in eng.buckets] == ["rest", "h.0", "h.1", "h.2"]
        assert eng.gradient_accumulation_steps() == ga
        # per-rank batches; rank 1 also exercises the otherwise unused branch
        batches = [[torch.randint(0, 11, (4, 5), generator=torch.Generator().manual_seed(100 * r + i))
                    for i in range(ga)] for r in range(world)]
        for i in range(ga):
            _, loss = eng(batches[rank][i], use_unused=(rank == 1))
            eng.backward(loss)
            if i + 1 < ga:  # not a boundary: nothing may have been communicated yet
                assert all(b.work is None for b in eng.buckets)
        eng.step()
        # expectation: mean over ranks of (sum over micro-steps of grad(loss * scale / ga))
        exp = {n: torch.zeros_like(p) for n, p in ref.named_parameters()}
        for r in range(world):
            ref.zero_grad()
            for i in range(ga):
                _, loss = ref(batches[r][i], use_unused=(r == 1))
                (loss * (8.0 / ga)).backward()
            for n, p in ref.named_parameters():
                if p.grad is not None:
                    exp[n] += p.grad / world
        for n, p in model.named_parameters():
            assert torch.allclose(p.grad, exp[n], rtol=1e-5, atol=1e-7), n
        # checkpoint round trip through the DeepSpeed-shaped layout
        eng.save_checkpoint(tmp, tag="latest_model", client_state={"iteration": 7})
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
        path, client = eng.load_checkpoint(tmp, "latest_model")
        assert path.endswith("mp_rank_00_model_states.pt") and client["iteration"] == 7
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            assert torch.equal(p, q), n
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ga", [1, 2])
def test_engine_gloo_world2(ga):
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp, ga), nprocs=2, join=True)


def test_load_deepspeed_layout_checkpoint_cpu(tmp_path):
    """A mp_rank_00_model_states.pt in DeepSpeed 0.6.7's key layout (FP16_Optimizer dict under `optimizer`, or None)
    loads into the engine: weights restored, DeepSpeed bookkeeping consumed, client state returned, no crash on the
    foreign optimizer state; the engine's own checkpoints still round-trip their optimizer state."""
    import warnings
    import torch
    from db1_sm100.engine import DB1Engine
    from tests.ds_ckpt import write_deepspeed_style_checkpoint
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    src = {k: torch.randn_like(v) for k, v in net.state_dict().items()}
    for fp16_opt in (True, False):
        d = tmp_path / ("ds_%d" % fp16_opt)
        write_deepspeed_style_checkpoint(str(d), src, fp16_optimizer=fp16_opt)
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
        eng = DB1Engine(net, optimizer=opt, loss_scale=128.0)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            path, client = eng.load_checkpoint(str(d))
        assert path.endswith("global_step1234/mp_rank_00_model_states.pt")
        assert client == {"iteration": 1234, "args": {"n_layer": 24}}
        assert eng.global_steps == 1234 and eng.skipped_steps == 3
        assert eng.loss_scale == (32768.0 if fp16_opt else 128.0)
        assert bool(w) == fp16_opt  # foreign optimizer state: warned about, not loaded
        for k, v in net.state_dict().items():
            assert torch.equal(v, src[k])
    # own format: optimizer state round trip
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    eng = DB1Engine(net, optimizer=opt, loss_scale=64.0)
    loss = net(torch.randn(3, 8)).pow(2).mean()
    eng.backward(loss)
    eng.step()
    d = tmp_path / "own"
    eng.save_checkpoint(str(d), tag="t1", client_state={"iteration": 7})
    st = torch.load(str(d / "t1" / "mp_rank_00_model_states.pt"), weights_only=False)
    for key in ("module", "buffer_names", "optimizer", "lr_scheduler", "sparse_tensor_module_names", "skipped_steps",
                "global_steps", "global_samples", "dp_world_size", "mp_world_size", "ds_config", "ds_version"):
        assert key in st, key
    opt2 = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    eng2 = DB1Engine(net, optimizer=opt2, loss_scale=1.0)
    _path, client = eng2.load_checkpoint(str(d))
    assert client == {"iteration": 7} and eng2.loss_scale == 64.0 and eng2.global_steps == 1
    assert len(opt2.state_dict()["state"]) == len(opt.state_dict()["state"]) > 0
