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
        assert [b.key for b in eng.buckets] == ["emb", "h.0", "h.1", "h.2", "rest"]
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


class _TiedHead(torch.autograd.Function):
    """logits = x @ W^T with the weight gradient written through the gradient sink (as HeadLossFn does on the GPU)."""

    @staticmethod
    def forward(ctx, x, W):
        ctx.save_for_backward(x, W)
        ctx.param = W
        return x @ W.t()

    @staticmethod
    def backward(ctx, dy):
        from db1_sm100 import functions as F_
        x, W = ctx.saved_tensors
        g = F_._Grad(ctx.param)
        val = dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
        if g.acc:
            g.buf.add_(val)
        else:
            g.buf.copy_(val)
        return dy @ W, g.ret()


class _Lookup(torch.autograd.Function):
    """e = W[tok] with the scatter-add of the gradient done through the sink (as EmbedFn does on the GPU)."""

    @staticmethod
    def forward(ctx, tok, W):
        from db1_sm100 import functions as F_
        F_._sink.note_tokens(tok)
        ctx.save_for_backward(tok)
        ctx.param = W
        return W[tok]

    @staticmethod
    def backward(ctx, de):
        from db1_sm100 import functions as F_
        (tok,) = ctx.saved_tensors
        g = F_._Grad(ctx.param, zero=True)
        if g.direct and not g.acc:
            g.buf.zero_()
        g.buf.index_add_(0, tok.reshape(-1), de.reshape(-1, de.shape[-1]).to(g.buf.dtype))
        return None, g.ret()


class _TiedToy(nn.Module):
    def __init__(self):
        super().__init__()
        self.word_embedding = nn.Embedding(23, 8)
        self.h = nn.ModuleList([nn.Linear(8, 8) for _ in range(2)])
        self.vision_encoder = nn.Linear(8, 8)

    def forward(self, toks, use_vision=False):
        """toks: list of token tensors (task segments), each embedded by its own lookup."""
        W = self.word_embedding.weight
        x = torch.cat([_Lookup.apply(t, W) for t in toks], dim=0)
        if use_vision:
            from db1_sm100 import functions as F_
            F_._sink.note_vision()
            x = x + self.vision_encoder(x)
        for l in self.h:
            x = torch.tanh(l(x))
        return None, _TiedHead.apply(x, W).pow(2).mean()

    def plain(self, toks, use_vision=False):
        W = self.word_embedding.weight
        x = torch.cat([W[t] for t in toks], dim=0)
        if use_vision:
            x = x + self.vision_encoder(x)
        for l in self.h:
            x = torch.tanh(l(x))
        return (x @ W.t()).pow(2).mean()


def _worker_tied(rank, world, port, ga, vision):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from db1_sm100 import functions as F_
        from db1_sm100.engine import DB1Engine
        torch.manual_seed(0)
        model = _TiedToy()
        ref = _TiedToy()
        ref.load_state_dict(model.state_dict())
        eng = DB1Engine(model, gradient_accumulation_steps=ga, loss_scale=4.0, direct_grads="force")
        assert [b.key for b in eng.buckets] == ["emb", "h.0", "h.1", "vision"]
        assert eng._emb is not None and eng._vision is not None
        sent = []
        orig = eng._launch_allreduce
        eng._launch_allreduce = lambda b: (sent.append(b.key), orig(b))[1]
        for window in range(2):  # second window: the side buffer must have been left all-zero
            los = [3 + window, 9 + window]  # rank r's tokens lie in [los[r], los[r] + 7): agreed range = union
            batches = [[[torch.randint(los[r], los[r] + 4, (3, 5), generator=torch.Generator().manual_seed(100 * r + 10 * window + i)),
                         torch.randint(los[r] + 2, los[r] + 7, (2, 5), generator=torch.Generator().manual_seed(7 + 100 * r + i))]
                        for i in range(ga)] for r in range(world)]
            del sent[:]
            for i in range(ga):
                _, loss = eng(batches[rank][i], use_vision=(vision and rank == 1))
                eng.backward(loss)
            assert sent[0] == "emb", sent  # the dense (head) part leaves first, from inside backward
            assert ("vision" in sent) == vision, sent
            assert eng._scatter.abs().max().item() == 0.0
            exp = {n: torch.zeros_like(p) for n, p in ref.named_parameters()}
            for r in range(world):
                ref.zero_grad()
                for i in range(ga):
                    (ref.plain(batches[r][i], use_vision=(vision and r == 1)) * (4.0 / ga)).backward()
                for n, p in ref.named_parameters():
                    if p.grad is not None:
                        exp[n] += p.grad / world
            for n, p in model.named_parameters():
                assert torch.allclose(p.grad, exp[n], rtol=1e-5, atol=1e-6), (n, window)
            eng.step()
        F_.set_grad_sink(None)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ga,vision", [(1, False), (2, True)])
def test_engine_tied_embedding_split_gloo_world2(ga, vision):
    """The tied embedding's gradient: dense head part all-reduced from inside backward, scatter part through the side
    buffer over the agreed row range only; the patch embedder's bucket is skipped when no rank used it."""
    port = _free_port()
    mp.spawn(_worker_tied, args=(2, port, ga, vision), nprocs=2, join=True)


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
