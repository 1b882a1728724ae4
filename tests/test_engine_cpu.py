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
