"""Data-parallel gradient exchange on real GPUs over NCCL (needs >= 2 visible GPUs; skipped otherwise): after
DB1Engine.backward every rank holds the MEAN over ranks of the single-GPU gradients of the per-rank batches
(SURVEY section 8e), with the bucket all-reduces launched from inside backward (gradient sink -> done() -> NCCL stream).
Run with `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import src.mpu as mpu
        from db1_sm100 import functions as F_, synth
        from db1_sm100.engine import DB1Engine
        from oracle import db1_oracle as orc
        from src.model import TransformerXL
        mpu.initialize_model_parallel()
        cfg = orc.tiny_config(text_vocab_size=480)
        sd = orc.synth_state_dict(cfg, seed=5)
        model = TransformerXL(cfg)
        model.load_state_dict(sd, strict=True)
        model = model.half().to(dev).eval()
        model.pos_emb.phase_dtype = torch.float32
        SCALE = 1024.0
        batches = [[synth.rl_continuous_batch(cfg, 2, 256, obs_len=5, act_len=2, seed=40 + r),
                    synth.nlp_batch(cfg, 1, 256, seed=50 + r)] for r in range(world)]
        # single-GPU reference: plain autograd on every rank's batch, averaged
        F_.set_grad_sink(None)
        mean = {}
        for r in range(world):
            for p in model.parameters():
                p.grad = None
            _, loss = model([synth.to_device(t, dev) for t in batches[r]])
            (loss * SCALE).backward()
            for n, p in model.named_parameters():
                if p.grad is not None:
                    mean[n] = mean.get(n, 0) + p.grad.float() / world
        for p in model.parameters():
            p.grad = None
        eng = DB1Engine(model, mpu=mpu, loss_scale=SCALE)
        assert eng._world == world and eng._overlap
        for _ in range(2):  # twice: the second window must not see stale bucket contents
            _, loss = eng([synth.to_device(t, dev) for t in batches[rank]])
            eng.backward(loss)
            torch.cuda.synchronize()
            worst, bad = 0.0, []
            for n, p in model.named_parameters():
                if n in mean:
                    d = ((p.grad.float() - mean[n]).norm() / (mean[n].norm() + 1e-20)).item()
                    worst = max(worst, d)
                    if d > 5e-3:
                        bad.append((n, round(d, 4)))
                else:
                    assert p.grad.abs().max().item() == 0, n
            assert not bad, bad[:12]  # tolerance: fp16 sum order / NCCL averaging
        out[rank] = worst
    finally:
        from db1_sm100 import functions as F_
        F_.set_grad_sink(None)
        dist.destroy_process_group()


def test_engine_nccl_allreduce_equals_mean_of_single_gpu_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
