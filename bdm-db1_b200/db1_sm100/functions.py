"""autograd.Function wrappers: each forward/backward is a fixed sequence of C-ABI kernel launches (db1_sm100.ops).

No arithmetic of the path runs as torch ops; torch supplies device buffers, the current stream and the autograd graph,
plus buffer plumbing (zero-fills of gradient accumulators, the loss mask's cast to fp32, concatenation of task segments,
contiguous copies of strided views: ~70 small ATen launches per step, < 1 % of its time, profiles/kernels_r2.md).
Shapes: activations are [rows = B*L, d] fp16 row-major.
"""
import math
import threading

import torch

from . import ops

_ws_lock = threading.Lock()
_ws = {}


def _workspace(key, shape, dtype, device, zero=True, tag=None):
    """Process-wide cached scratch (attention-backward score buffers, dlogits). Zero-filled at creation: the causal
    GEMMs read whole 128-row blocks, so tiles the recompute kernel does not visit (above the diagonal, outside the
    attention window) must stay zero. The set of visited tiles depends on the shape AND on `tag` (the window): a buffer
    reused under another tag is zero-filled again. At most one buffer per (key, device) is kept, so a change of shape
    frees the old one instead of pinning 3*B*H*L*L*2 bytes per distinct (B, L) for the life of the process."""
    k = (key, device.index)
    with _ws_lock:
        ent = _ws.get(k)
        if ent is not None and (ent[0].shape != tuple(shape) or ent[0].dtype != dtype):
            ent = None
        if ent is None:
            t = torch.zeros(shape, dtype=dtype, device=device) if zero else torch.empty(shape, dtype=dtype, device=device)
            _ws[k] = (t, tag)
            return t
        t, old_tag = ent
        if old_tag != tag:
            if zero:
                t.zero_()
            _ws[k] = (t, tag)
    return t


def clear_workspaces():
    with _ws_lock:
        _ws.clear()


class _Seeds:
    """Counter-based dropout seeds: (torch.initial_seed(), call counter) -> 64-bit seed, one per dropout site per call."""

    def __init__(self):
        self.counter = 0

    def next(self):
        self.counter += 1
        s = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self.counter * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        return s


seeds = _Seeds()


# ---------------------------------------------------------------------------------------------------------------------
# Gradient sink: when a training engine registers one (DB1Engine does), weight-gradient kernels write straight into the
# engine's flat all-reduce buckets (first write of an accumulation window overwrites, later ones accumulate) and the
# engine is told when a parameter's gradient is complete - no autograd `.grad +=` kernels, no bucket zero-fill.
# Without a sink (plain module use, parity tests) the Functions return gradients to autograd as usual.
# ---------------------------------------------------------------------------------------------------------------------
_sink = None


def set_grad_sink(sink):
    """sink.target(param) -> (fp16 view shaped like param, accumulate: bool) or None;  sink.done(param)."""
    global _sink
    _sink = sink


class _Grad:
    """Destination of one parameter gradient inside a backward: the sink's bucket view, or a fresh tensor for autograd."""
    __slots__ = ("param", "buf", "acc", "direct")

    def __init__(self, param, shape=None, zero=False):
        t = _sink.target(param) if (_sink is not None and param is not None) else None
        self.param = param
        if t is not None:
            self.buf, self.acc = t
            self.direct = True
        else:
            shape = tuple(param.shape) if shape is None else shape
            mk = torch.zeros if zero else torch.empty
            self.buf, self.acc, self.direct = mk(shape, dtype=torch.float16, device=param.device), False, False

    def ret(self):
        """Value to hand back to autograd (None when the gradient already sits in the engine's bucket)."""
        if self.direct:
            _sink.done(self.param)
            return None
        return self.buf


# ---------------------------------------------------------------------------------------------------------------------
# Weight-gradient GEMMs on a second stream - OPT-IN (DB1_WGRAD_STREAM=1), measured and rejected as a default.
# Every large GEMM of the step covers the 74 CTA pairs with 64 / 128 / 192 / 256 / 384 pair-tiles - a last wave that is
# 86 % full at best, and a persistent launch cannot start before the previous one has drained. The weight gradients are
# off the critical chain of backward (nothing but the all-reduce consumes them), so they can be enqueued on a side
# stream where their CTAs take the SMs the data-gradient GEMM's last wave leaves idle, and vice versa. Inputs are
# fenced by an event of the main stream, lifetimes by record_stream; the engine fences the bucket all-reduces and the
# end of backward on both streams. Measured on a B200 (profiles/README.md, round 2): 42.2 ms per step against 34.6 ms
# on one stream - two persistent tcgen05 GEMMs sharing the SMs lose the lock-step in which the CTAs of one launch
# re-use each other's operand tiles in L2 (the same effect that made the stream-K tail lose), which costs far more
# than the idle tail waves.
# ---------------------------------------------------------------------------------------------------------------------
_side_streams = {}


def wgrad_stream(device):
    """The side stream of `device` when weight-gradient GEMMs run concurrently (engine sink registered), else None."""
    import os
    if _sink is None or not getattr(_sink, "wgrad_side_stream", False) or os.environ.get("DB1_WGRAD_STREAM", "0") != "1":
        return None
    st = _side_streams.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[device.index] = st
    return st


class _Side:
    """with _Side(dev, inputs...): <weight-gradient GEMM launches>"""
    __slots__ = ("st", "ctx")

    def __init__(self, device, *tensors):
        self.st = wgrad_stream(device)
        self.ctx = None
        if self.st is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self.st.wait_event(ev)
            for t in tensors:
                if t is not None:
                    t.record_stream(self.st)

    def __enter__(self):
        if self.st is not None:
            self.ctx = torch.cuda.stream(self.st)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


# ---------------------------------------------------------------------------------------------------------------------
# Small fp32 gradient accumulators (du, dv, dgamma, dbeta, biases). Under an engine one zero-filled arena per backward
# replaces the per-block torch.zeros launches (one memset instead of 48), and the feed-forward block's conversion to
# fp16 rides along with the attention block's of the same layer (one db1_f32_to_f16_multi launch per layer instead of
# two): both accumulators are views of the same arena, so one launch can address them through offsets.
# ---------------------------------------------------------------------------------------------------------------------
class _Arena:
    def __init__(self):
        self.buf = None
        self.off = 0
        self.active = False
        self.want = 0      # floats requested during the last backward (sizes the arena for the next one)
        self.pending = []  # deferred conversions: (offset of the block's accumulator in the arena, plan, grads)


_arena = _Arena()


def begin_backward(device):
    """Called by the engine before autograd runs: (re)arms the arena. The first backward only measures demand."""
    a = _arena
    need = a.want + a.want // 4
    if need > 0 and (a.buf is None or a.buf.numel() < need or a.buf.device != device):
        a.buf = torch.empty(need, dtype=torch.float32, device=device)
    a.active = a.buf is not None and a.buf.device == device
    if a.active:
        a.buf.zero_()
    a.off = 0
    a.want = 0
    a.pending = []


def end_backward():
    """Called by the engine after autograd returned: converts whatever is still deferred and disarms the arena."""
    _flush_small()
    _arena.active = False


def _f32zeros(n, dev):
    a = _arena
    n4 = (n + 3) // 4 * 4  # 16-byte aligned slices (vector reductions)
    a.want += n4
    if a.active and a.off + n4 <= a.buf.numel():
        v = a.buf[a.off:a.off + n]
        a.off += n4
        return v
    return torch.zeros(n, dtype=torch.float32, device=dev)


def _arena_offset(small):
    a = _arena
    if not a.active or small.untyped_storage().data_ptr() != a.buf.untyped_storage().data_ptr():
        return None
    return small.storage_offset()


def _flush_small():
    a = _arena
    if not a.pending:
        return
    segs, fin = [], []
    for base, plan, grads in a.pending:
        for g, (_p, o, n) in zip(grads, plan):
            segs.append((g.buf, base + o, n, g.acc))
            fin.append(g)
    a.pending = []
    ops.f32_to_f16_multi(a.buf, segs)
    for g in fin:
        g.ret()


def _scatter_small(small, plan, defer=False):
    """plan: list of (param, offset, n). Converts slices of the fp32 accumulator `small` to fp16 gradients with one
    launch per 8 parameters; returns the autograd return values in plan order. defer=True (feed-forward block): when
    every gradient goes straight into an engine bucket and the accumulator lives in the arena, the conversion is left
    to the next non-deferred call (the attention block of the same layer) and the returns are None."""
    grads = [_Grad(p, (n,)) for p, _o, n in plan]
    base = _arena_offset(small)
    direct = base is not None and all(g.direct for g in grads)
    if direct:
        _arena.pending.append((base, plan, grads))
        if not defer:
            _flush_small()
        return [None] * len(plan)
    if _arena.pending:
        _flush_small()
    ops.f32_to_f16_multi(small, [(g.buf, o, n, g.acc) for g, (_p, o, n) in zip(grads, plan)])
    out = []
    for g, (p, _o, _n) in zip(grads, plan):
        r = g.ret()
        out.append(r.view(p.shape) if r is not None else None)
    return out


def _to_half(x32, shape):
    out = torch.empty(shape, dtype=torch.float16, device=x32.device)
    ops.f32_to_f16(x32, out)
    return out


def _attn_bwd_chunk(B):
    import os
    v = int(os.environ.get("DB1_ATTN_BWD_CHUNK", "0"))
    return B if v <= 0 else min(B, v)


class AttnBlockFn(torch.autograd.Function):
    """RelPartialLearnableMultiHeadAttn.forward, post-LN branch (transformer_xl.py:112-243):
    out = LayerNorm(w + dropout(o_net(rel_attention(qkv_net(w), r_net(r)))))."""

    @staticmethod
    def forward(ctx, w, r, Wqkv, Wr, Wo, u, v, gamma, beta, H, eps, drop_p, window):
        B, L, d = w.shape
        dh = d // H
        rows = B * L
        dev = w.device
        x2 = w.reshape(rows, d)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        uf = u.reshape(d)
        vf = v.reshape(d)
        qkv4 = torch.empty(rows, 4 * d, dtype=torch.float16, device=dev)
        ops.gemm(x2, Wqkv, qkv4, rows, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=uf, v=vf, d_model=d)
        rk = torch.empty(L, d, dtype=torch.float16, device=dev)
        ops.gemm(r, Wr, rk, L, d, d, lda=d, ldb=d, ldc=d)
        o = torch.empty(rows, d, dtype=torch.float16, device=dev)
        lse2 = torch.empty(B, H, L, dtype=torch.float32, device=dev)
        scale = 1.0 / math.sqrt(dh)
        ops.relattn_fwd(qkv4, rk, o, lse2, B, L, H, dh, window, scale)
        seed = seeds.next() if drop_p > 0 else 0
        y = torch.empty(rows, d, dtype=torch.float16, device=dev)
        ops.gemm(o, Wo, y, rows, d, d, lda=d, ldb=d, ldc=d, resid=x2, ldr=d, drop_p=drop_p, seed=seed)
        out = torch.empty(rows, d, dtype=torch.float16, device=dev)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=dev)
        ops.layernorm_fwd(y, gamma, beta, out, stats, eps)
        ctx.save_for_backward(x2, r, Wqkv, Wr, Wo, gamma, qkv4, rk, o, lse2, y, stats)
        ctx.cfg = (B, L, d, H, dh, drop_p, seed, window, scale)
        ctx.params = (Wqkv, Wr, Wo, u, v, gamma, beta)
        return out.view(B, L, d)

    @staticmethod
    def backward(ctx, dout):
        x2, r, Wqkv, Wr, Wo, gamma, qkv4, rk, o, lse2, y, stats = ctx.saved_tensors
        B, L, d, H, dh, drop_p, seed, window, scale = ctx.cfg
        rows = B * L
        dev = x2.device
        f16 = torch.float16
        dout2 = dout.reshape(rows, d)
        if not dout2.is_contiguous():
            dout2 = dout2.contiguous()
        small = _f32zeros(4 * d, dev)  # [du | dv | dgamma | dbeta], converted to fp16 by one launch at the end
        du, dv, dgamma, dbeta = small[0:d], small[d:2 * d], small[2 * d:3 * d], small[3 * d:4 * d]
        dy = torch.empty(rows, d, dtype=f16, device=dev)
        dz = torch.empty(rows, d, dtype=f16, device=dev) if drop_p > 0 else None
        ops.layernorm_bwd(dout2, y, gamma, stats, dy, dz, dgamma, dbeta, None, drop_p, seed)
        dzz = dz if dz is not None else dy
        pWqkv, pWr, pWo, pu, pv, pgamma, pbeta = ctx.params
        # o_net
        gWo = _Grad(pWo)
        with _Side(dev, dzz, o):
            ops.gemm(dzz, o, gWo.buf, d, d, rows, lda=d, ldb=d, ldc=d, a_mn=True, b_mn=True, accumulate=gWo.acc)
        do = torch.empty(rows, d, dtype=f16, device=dev)
        # D = rowsum(dO * O), the softmax-backward row term: from the dO GEMM's own epilogue at model size (head dim 128),
        # else inside the recompute kernel; no separate db1_rowdot pass over dO / O either way
        Drow = None
        if ops.gemm_dot_supported(rows, d, H, dh):
            Drow = torch.empty(B, H, L, dtype=torch.float32, device=dev)
            ops.gemm(dzz, Wo, do, rows, d, d, lda=d, ldb=d, ldc=d, b_mn=True, dot=(o, Drow, L, H))
        else:
            ops.gemm(dzz, Wo, do, rows, d, d, lda=d, ldb=d, ldc=d, b_mn=True)
        # attention core: the recompute kernel writes P and dS once; three kernels with TMEM-resident accumulators consume them
        # (csrc/relattn_bwd.cu): key-outer dV / dK, query-outer dq (+ du, dv), diagonal-outer dR; the adjoint of _rel_shift
        # is done on registers - no re-laid-out copy of dS, no generic batched GEMMs, no separate dq / bias-gradient pass
        # The four kernels run per chunk of `bc` sequences so that the chunk's P + dS scratch (2 * bc * H * nq^2 * 32 KB) can
        # stay in the 126 MB L2 between its producer and its three consumers (DB1_ATTN_BWD_CHUNK, default: whole batch).
        bc = _attn_bwd_chunk(B)
        tshape = ops.score_tiles_shape(bc, L, H)  # tiled scratch: only tiles the recompute kernel wrote are ever read
        P = _workspace("P", tshape, f16, dev, zero=False)
        dS = _workspace("dS", tshape, f16, dev, zero=False)
        dqkv = torch.empty(rows, 3 * d, dtype=f16, device=dev)
        dr32 = _f32zeros(L * d, dev).view(L, d)
        for b0 in range(0, B, bc):
            nb = min(bc, B - b0)
            rs = slice(b0 * L, (b0 + nb) * L)
            q4, do_c, o_c, dq_c = qkv4[rs], do[rs], o[rs], dqkv[rs]
            ops.relattn_bwd_ds_tiled(q4, rk, do_c, lse2[b0:b0 + nb], Drow[b0:b0 + nb] if Drow is not None else None, P, dS,
                                     nb, L, H, dh, window, scale, o=o_c)
            ops.relattn_bwd_dkdv(P, dS, do_c, q4[:, 0:d], dq_c[:, d:2 * d], dq_c[:, 2 * d:], nb, L, H, dh, window)
            ops.relattn_bwd_dq(dS, q4[:, 2 * d:3 * d], rk, dq_c[:, 0:d], du, dv, nb, L, H, dh, window)
            ops.relattn_bwd_dr(dS, q4[:, d:2 * d], dr32, nb, L, H, dh, window)
        drk = _to_half(dr32, (L, d))
        # r_net / qkv_net
        gWr = _Grad(pWr)
        gWqkv = _Grad(pWqkv)
        with _Side(dev, drk, r, dqkv, x2):
            ops.gemm(drk, r, gWr.buf, d, d, L, lda=d, ldb=d, ldc=d, a_mn=True, b_mn=True, accumulate=gWr.acc)
            ops.gemm(dqkv, x2, gWqkv.buf, 3 * d, d, rows, lda=3 * d, ldb=d, ldc=d, a_mn=True, b_mn=True,
                     accumulate=gWqkv.acc)
        dx = torch.empty(rows, d, dtype=f16, device=dev)
        ops.gemm(dqkv, Wqkv, dx, rows, d, 3 * d, lda=3 * d, ldb=d, ldc=d, b_mn=True, resid=dy, ldr=d)
        gu, gv, gg, gb = _scatter_small(small, [(pu, 0, d), (pv, d, d), (pgamma, 2 * d, d), (pbeta, 3 * d, d)])
        return (dx.view(B, L, d), None, gWqkv.ret(), gWr.ret(), gWo.ret(), gu, gv, gg, gb, None, None, None, None)


def attn_block_with_memory(w, mem, r, Wqkv, Wr, Wo, u, v, gamma, beta, H, eps, window):
    """RelPartialLearnableMultiHeadAttn.forward with `mems` (transformer_xl.py:124-133, :141-238), inference only:
    keys / values come from cat(mem, w), queries are the last qlen rows, r has klen = mlen + qlen rows."""
    B, qlen, d = w.shape
    mlen = mem.shape[1]
    K = mlen + qlen
    dh = d // H
    dev = w.device
    f16 = torch.float16
    cat = torch.cat([mem, w], dim=1).reshape(B * K, d)  # buffer plumbing only (the reference concatenates too, :125)
    qkv4 = torch.empty(B * K, 4 * d, dtype=f16, device=dev)
    ops.gemm(cat, Wqkv, qkv4, B * K, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=u.reshape(d),
             v=v.reshape(d), d_model=d)
    rk = torch.empty(K, d, dtype=f16, device=dev)
    ops.gemm(r, Wr, rk, K, d, d, lda=d, ldb=d, ldc=d)
    o = torch.empty(B, K, d, dtype=f16, device=dev)
    lse2 = torch.empty(B, H, K, dtype=torch.float32, device=dev)
    ops.relattn_mem_fwd(qkv4, rk, o.view(B * K, d), lse2, B, K, H, dh, window, 1.0 / math.sqrt(dh), mlen)
    # the query rows as a dense [B*qlen, d] matrix. (For qlen == 1 reshape() returns a strided VIEW with row stride K*d, not a
    # copy - the GEMM below is told lda = d, so the copy must be explicit.)
    oq = o[:, mlen:].contiguous().view(B * qlen, d)
    x2 = w.reshape(B * qlen, d)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    y = torch.empty(B * qlen, d, dtype=f16, device=dev)
    ops.gemm(oq, Wo, y, B * qlen, d, d, lda=d, ldb=d, ldc=d, resid=x2, ldr=d)
    out = torch.empty(B * qlen, d, dtype=f16, device=dev)
    stats = torch.empty(B * qlen, 2, dtype=torch.float32, device=dev)
    ops.layernorm_fwd(y, gamma, beta, out, stats, eps)
    return out.view(B, qlen, d)


class KVMemory:
    """Transformer-XL memory for the decode loop with cached projections. Per layer: ring buffers of the
    layer-input hidden states (what the reference's `mems` hold), their keys and their values, [B, mem_len, d] fp16 each,
    starting from zeros exactly like init_mem (transformer_xl.py:470-485: zero hidden rows ARE attended, their k = v = 0).
    `head` = slot of the oldest row. to_mems() returns the reference-format list (logical order)."""

    def __init__(self, n_layer, batch_size, mem_len, d, device):
        z = lambda: torch.zeros(batch_size, mem_len, d, dtype=torch.float16, device=device)  # noqa: E731
        self.hid = [z() for _ in range(n_layer)]
        self.k = [z() for _ in range(n_layer)]
        self.v = [z() for _ in range(n_layer)]
        self.cap = mem_len
        self.head = 0                                                        # host copy (to_mems)
        self.head_dev = torch.zeros(1, dtype=torch.int32, device=device)    # what the kernels read (CUDA-graph safe)
        self.batch_size = batch_size
        self.rk = {}   # (layer, klen) -> r_net(pos_emb(klen)) [klen, d]
        self.ws = None

    def to_mems(self):
        return [torch.roll(h, -self.head, dims=1) for h in self.hid]

    def advance(self, q):
        """The q rows every layer appended become the newest memory rows. (In-place device update: captured by graphs.)"""
        self.head_dev.add_(q).remainder_(self.cap)
        self.head = (self.head + q) % self.cap

    def __len__(self):
        return len(self.hid)


class DecodeGraph:
    """One decode step of fixed shape (B sequences x q new tokens) captured in a CUDA graph: the ~350 kernel launches of a
    24-layer step replay as one submission (the step is launch-bound from Python otherwise). The ring-buffer head lives
    in device memory, so every replay appends at the right slots.
        g = DecodeGraph(model, mem, q);  logits = g.step(tokens [B, q], position_id [B, q])   # logits: static buffer"""

    def __init__(self, model, mem, q, task_cls=None):
        from src.data.input_specs import RLTaskInput
        dev = mem.head_dev.device
        B = mem.batch_size
        self.mem, self.q = mem, q
        self.tok = torch.zeros(B, q, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(B, q, dtype=torch.int64, device=dev)
        self.inputs = [RLTaskInput(position_id=self.pos, attention_mask=None, loss_mask=None, label=None, text_seq=None,
                                   vision_seq=None, tensor_seq=self.tok)]
        keep = [(t, t.clone()) for lst in (mem.hid, mem.k, mem.v) for t in lst] + [(mem.head_dev, mem.head_dev.clone())]
        head0 = mem.head
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():  # warm-up: workspaces, r-table cache, allocator pools
            for _ in range(2):
                model(self.inputs, compute_loss=False, mems=mem)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.logits, _none, _m = model(self.inputs, compute_loss=False, mems=mem)
        for t, c in keep:  # warm-up and capture bookkeeping must not leave traces in the memory
            t.copy_(c)
        mem.head = head0

    def step(self, tok, pos=None):
        self.tok.copy_(tok, non_blocking=True)
        if pos is not None:
            self.pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        self.mem.head = (self.mem.head + self.q) % self.mem.cap
        return self.logits


def attn_block_cached(w, li, mem, r, Wqkv, Wr, Wo, u, v, gamma, beta, H, eps, window):
    """One attention block of a decode step on a KVMemory: only the B*Q new rows go through qkv_net; the attention runs
    over [cached k / v | new k / v] (db1_relattn_decode); afterwards the new rows replace the oldest ring slots."""
    B, Q, d = w.shape
    dh = d // H
    dev = w.device
    f16 = torch.float16
    K = mem.cap + Q
    x2 = w.reshape(B * Q, d)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    qkv4 = torch.empty(B * Q, 4 * d, dtype=f16, device=dev)
    ops.gemm(x2, Wqkv, qkv4, B * Q, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=u.reshape(d), v=v.reshape(d),
             d_model=d, b_static=True)
    rk = mem.rk.get((li, K))
    if rk is None:  # depends on the layer's r_net and on klen only
        rk = torch.empty(K, d, dtype=f16, device=dev)
        ops.gemm(r, Wr, rk, K, d, d, lda=d, ldb=d, ldc=d)
        mem.rk[(li, K)] = rk
    need = B * Q * H * ops.decode_splits(B, Q, H) * (dh + 2)
    if mem.ws is None or mem.ws.numel() < need:
        mem.ws = torch.empty(need, dtype=torch.float32, device=dev)
    o = torch.empty(B * Q, d, dtype=f16, device=dev)
    ops.relattn_decode(qkv4, mem.k[li], mem.v[li], mem.head, rk, o, mem.ws, B, Q, H, dh, window, 1.0 / math.sqrt(dh),
                       head_dev=mem.head_dev)
    ops.ring_append([(x2, mem.hid[li]), (qkv4[:, 2 * d:3 * d], mem.k[li]), (qkv4[:, 3 * d:], mem.v[li])], mem.head, B, Q,
                    head_dev=mem.head_dev)
    y = torch.empty(B * Q, d, dtype=f16, device=dev)
    ops.gemm(o, Wo, y, B * Q, d, d, lda=d, ldb=d, ldc=d, resid=x2, ldr=d, b_static=True)
    out = torch.empty(B * Q, d, dtype=f16, device=dev)
    stats = torch.empty(B * Q, 2, dtype=torch.float32, device=dev)
    ops.layernorm_fwd(y, gamma, beta, out, stats, eps)
    return out.view(B, Q, d)


def decode_step_fused(model, hidden, mem, pos_rows, window):
    """One cached decode step with rows = B * Q <= 8, all layers and the tied head: every LayerNorm is applied by the
    few-row GEMM that consumes it (db1_gemm_desc.ln_*: normalise-on-load; the normalised rows are written once, by that
    GEMM, for the residual add and the memory append that need them) - 49 launches less than block-by-block execution.
    Same arithmetic and the same fp16 rounding points as attn_block_cached + FFBlockFn + head_logits. Returns the logits."""
    B, Q, d = hidden.shape
    rows = B * Q
    dev = hidden.device
    f16 = torch.float16
    K = mem.cap + Q
    x2 = hidden.reshape(rows, d)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    pend = None  # (pre-norm rows, gamma, beta, eps) of the previous block
    for li, layer in enumerate(model.h):
        at, ff = layer.dec_attn, layer.pos_ff
        H = at.n_head
        dh = d // H
        # ---- attention block: qkv (LayerNorm of the previous block on load), attention over the cache, o_net + residual
        qkv4 = torch.empty(rows, 4 * d, dtype=f16, device=dev)
        if pend is not None:
            xin = torch.empty(rows, d, dtype=f16, device=dev)
            ops.gemm(pend[0], at.qkv_net.weight, qkv4, rows, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV,
                     u=at.r_w_bias.reshape(d), v=at.r_r_bias.reshape(d), d_model=d, b_static=True,
                     ln=(pend[1], pend[2], pend[3], xin))
            x2 = xin
        else:
            ops.gemm(x2, at.qkv_net.weight, qkv4, rows, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV,
                     u=at.r_w_bias.reshape(d), v=at.r_r_bias.reshape(d), d_model=d, b_static=True)
        rk = mem.rk.get((li, K))
        if rk is None:
            rk = torch.empty(K, d, dtype=f16, device=dev)
            ops.gemm(pos_rows, at.r_net.weight, rk, K, d, d, lda=d, ldb=d, ldc=d)
            mem.rk[(li, K)] = rk
        need = rows * H * ops.decode_splits(B, Q, H) * (dh + 2)
        if mem.ws is None or mem.ws.numel() < need:
            mem.ws = torch.empty(need, dtype=torch.float32, device=dev)
        o = torch.empty(rows, d, dtype=f16, device=dev)
        ops.relattn_decode(qkv4, mem.k[li], mem.v[li], mem.head, rk, o, mem.ws, B, Q, H, dh, window, 1.0 / math.sqrt(dh),
                           head_dev=mem.head_dev)
        ops.ring_append([(x2, mem.hid[li]), (qkv4[:, 2 * d:3 * d], mem.k[li]), (qkv4[:, 3 * d:], mem.v[li])], mem.head, B, Q,
                        head_dev=mem.head_dev)
        ya = torch.empty(rows, d, dtype=f16, device=dev)
        ops.gemm(o, at.o_net.weight, ya, rows, d, d, lda=d, ldb=d, ldc=d, resid=x2, ldr=d, b_static=True)
        # ---- feed-forward block: ff1 + GeGLU (the attention block's LayerNorm on load), ff2 + bias + residual
        W1, b1, W2, b2 = ff.CoreNet[0].weight, ff.CoreNet[0].bias, ff.CoreNet[2].weight, ff.CoreNet[2].bias
        F = W1.shape[0] // 2
        x3 = torch.empty(rows, d, dtype=f16, device=dev)
        g = torch.empty(rows, F, dtype=f16, device=dev)
        ops.gemm(ya, W1, g, rows, 2 * F, d, lda=d, ldb=d, ldc=F, epilogue=ops.EPI_GEGLU, bias=b1, F=F, b_static=True,
                 ln=(at.layer_norm.weight, at.layer_norm.bias, at.layer_norm.eps, x3))
        yf = torch.empty(rows, d, dtype=f16, device=dev)
        ops.gemm(g, W2, yf, rows, d, F, lda=F, ldb=F, ldc=d, bias=b2, resid=x3, ldr=d, b_static=True)
        pend = (yf, ff.layer_norm.weight, ff.layer_norm.bias, ff.layer_norm.eps)
    W = model.word_embedding.weight if model.share_input_output_embedding else model.lm_head.weight
    V = W.shape[0]
    Vp = (V + 127) // 128 * 128
    buf = torch.empty(rows, Vp, dtype=f16, device=dev)
    ops.gemm(pend[0], W, buf, rows, V, d, lda=d, ldb=d, ldc=Vp, b_static=True, ln=(pend[1], pend[2], pend[3], None))
    return buf.view(B, Q, Vp)[:, :, :V]


def decode_step_fused_applies(model, rows, d):
    """rows <= 8 on the few-row GEMM path, the released layer type (post-LN, GeGLU)."""
    import os
    if os.environ.get("DB1_DECODE_UNFUSED"):
        return False
    if not ops.few_row_gemm_applies(rows, d, d):
        return False
    l0 = model.h[0]
    return (not l0.dec_attn.pre_lnorm) and getattr(l0.pos_ff, "activation", "geglu") == "geglu" and len(model.h) > 0


class FFBlockFn(torch.autograd.Function):
    """PositionwiseFF.forward, post-LN branch with GeGLU (transformer_xl.py:276-292, activations.py:19-32):
    out = LayerNorm(x + dropout(W2 (a * gelu(g)) + b2)), [a|g] = W1 x + b1."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, gamma, beta, eps, drop_p):
        B, L, d = x.shape
        rows = B * L
        dev = x.device
        F2 = W1.shape[0]
        F = F2 // 2
        x2 = x.reshape(rows, d)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        Hb = torch.empty(rows, F2, dtype=torch.float16, device=dev)
        g = torch.empty(rows, F, dtype=torch.float16, device=dev)
        static = not any(ctx.needs_input_grad)  # inference (no_grad): the weights are not being written by anyone
        ops.gemm(x2, W1, g, rows, F2, d, lda=d, ldb=d, ldc=F, epilogue=ops.EPI_GEGLU, bias=b1, H=Hb, ldh=F2, F=F,
                 b_static=static)
        seed = seeds.next() if drop_p > 0 else 0
        y = torch.empty(rows, d, dtype=torch.float16, device=dev)
        ops.gemm(g, W2, y, rows, d, F, lda=F, ldb=F, ldc=d, bias=b2, resid=x2, ldr=d, drop_p=drop_p, seed=seed,
                 b_static=static)
        out = torch.empty(rows, d, dtype=torch.float16, device=dev)
        stats = torch.empty(rows, 2, dtype=torch.float32, device=dev)
        ops.layernorm_fwd(y, gamma, beta, out, stats, eps)
        ctx.save_for_backward(x2, W1, W2, gamma, Hb, g, y, stats)
        ctx.cfg = (B, L, d, F, drop_p, seed)
        ctx.params = (W1, b1, W2, b2, gamma, beta)
        return out.view(B, L, d)

    @staticmethod
    def backward(ctx, dout):
        x2, W1, W2, gamma, Hb, g, y, stats = ctx.saved_tensors
        B, L, d, F, drop_p, seed = ctx.cfg
        rows = B * L
        dev = x2.device
        f16 = torch.float16
        dout2 = dout.reshape(rows, d)
        if not dout2.is_contiguous():
            dout2 = dout2.contiguous()
        small = _f32zeros(3 * d + 2 * F, dev)  # [dgamma | dbeta | db2 | db1]
        dgamma, dbeta, db2, db1 = small[0:d], small[d:2 * d], small[2 * d:3 * d], small[3 * d:]
        dy = torch.empty(rows, d, dtype=f16, device=dev)
        dz = torch.empty(rows, d, dtype=f16, device=dev) if drop_p > 0 else None
        ops.layernorm_bwd(dout2, y, gamma, stats, dy, dz, dgamma, dbeta, db2, drop_p, seed)
        dzz = dz if dz is not None else dy
        pW1, pb1, pW2, pb2, pgamma, pbeta = ctx.params
        gW2 = _Grad(pW2)
        with _Side(dev, dzz, g):
            ops.gemm(dzz, g, gW2.buf, d, F, rows, lda=d, ldb=F, ldc=F, a_mn=True, b_mn=True, accumulate=gW2.acc)
        dH = torch.empty(rows, 2 * F, dtype=f16, device=dev)
        ops.gemm(dzz, W2, dH, rows, F, d, lda=d, ldb=F, ldc=2 * F, b_mn=True, epilogue=ops.EPI_DGEGLU, H=Hb,
                 ldh=2 * F, F=F)
        ops.colsum(dH, db1, rows, 2 * F)
        gW1 = _Grad(pW1)
        with _Side(dev, dH, x2):
            ops.gemm(dH, x2, gW1.buf, 2 * F, d, rows, lda=2 * F, ldb=d, ldc=d, a_mn=True, b_mn=True, accumulate=gW1.acc)
        dx = torch.empty(rows, d, dtype=f16, device=dev)
        ops.gemm(dH, W1, dx, rows, d, 2 * F, lda=2 * F, ldb=d, ldc=d, b_mn=True, resid=dy, ldr=d)
        gg, gb, gb2, gb1 = _scatter_small(small, [(pgamma, 0, d), (pbeta, d, d), (pb2, 2 * d, d), (pb1, 3 * d, 2 * F)], defer=True)
        return (dx.view(B, L, d), gW1.ret(), gb1, gW2.ret(), gb2, gg, gb, None, None)


def _pad8(n):
    return (n + 7) // 8 * 8


class HeadLossFn(torch.autograd.Function):
    """Tied LM head + masked-mean cross-entropy (transformer_xl.py:593-613). Returns (logits view [B,L,V], loss)."""

    @staticmethod
    def forward(ctx, hidden, W, labels, mask):
        B, L, d = hidden.shape
        rows = B * L
        V = W.shape[0]
        Vp = (V + 127) // 128 * 128
        dev = hidden.device
        h2 = hidden.reshape(rows, d)
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        buf = torch.empty(rows, Vp, dtype=torch.float16, device=dev)
        ops.gemm(h2, W, buf, rows, V, d, lda=d, ldb=d, ldc=Vp)
        lab = labels.reshape(rows).contiguous()
        msk = mask.reshape(rows).to(torch.float32).contiguous()
        row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
        row_lse = torch.empty(rows, dtype=torch.float32, device=dev)
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        ops.ce_fwd(buf, lab, msk, row_loss, row_lse, loss2, V)
        logits = buf.view(B, L, Vp)[:, :, :V]
        ctx.save_for_backward(h2, W, buf, lab, msk, row_lse, loss2)
        ctx.cfg = (B, L, d, V, Vp)
        ctx.params = (W,)
        ctx.mark_non_differentiable(logits)
        return logits, loss2[0]

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        h2, W, buf, lab, msk, row_lse, loss2 = ctx.saved_tensors
        B, L, d, V, Vp = ctx.cfg
        rows = B * L
        dev = h2.device
        gs = dloss.reshape(1).to(torch.float32)
        dl = _workspace("dlogits", (rows, Vp), torch.float16, dev)
        ops.ce_bwd(buf, lab, msk, row_lse, loss2, gs, dl, V)
        gW = _Grad(ctx.params[0])
        with _Side(dev, dl, h2):
            ops.gemm(dl, h2, gW.buf, V, d, rows, lda=Vp, ldb=d, ldc=d, a_mn=True, b_mn=True, accumulate=gW.acc)
        dh = torch.empty(rows, d, dtype=torch.float16, device=dev)
        ops.gemm(dl, W, dh, rows, d, V, lda=Vp, ldb=d, ldc=d, b_mn=True)
        return dh.view(B, L, d), gW.ret(), None, None


def head_logits(hidden, W):
    """Inference-only tied head (no autograd)."""
    B, L, d = hidden.shape
    rows = B * L
    V = W.shape[0]
    Vp = (V + 127) // 128 * 128
    h2 = hidden.reshape(rows, d).contiguous()
    buf = torch.empty(rows, Vp, dtype=torch.float16, device=hidden.device)
    ops.gemm(h2, W, buf, rows, V, d, lda=d, ldb=d, ldc=Vp, b_static=True)
    return buf.view(B, L, Vp)[:, :, :V]


class EmbedFn(torch.autograd.Function):
    """Embedding assembly for one task segment (transformer_xl.py:627-649 / :665): word lookup, image-patch slots,
    RL local-timestep embedding, embedding dropout (:545)."""

    @staticmethod
    def forward(ctx, tok, pos, W, T, vis, drop_p):
        B, L = tok.shape
        V, d = W.shape
        dev = W.device
        tok = tok.contiguous()
        pos = pos.contiguous() if pos is not None else None
        if vis is not None and not vis.is_contiguous():
            vis = vis.contiguous()
        out = torch.empty(B, L, d, dtype=torch.float16, device=dev)
        slot = torch.empty(B, L, dtype=torch.int32, device=dev)
        seed = seeds.next() if drop_p > 0 else 0
        note = getattr(_sink, "note_tokens", None)
        if note is not None and W.requires_grad:
            note(tok)  # the engine agrees on the embedding rows this window touches (range-limited all-reduce)
        ops.embed_fwd(tok, pos, slot, W, T if pos is not None else None, vis, out, L * d, B, L, d, V, drop_p, seed, 0)
        ctx.save_for_backward(tok, pos, slot)
        ctx.cfg = (B, L, d, V, drop_p, seed, T.shape[0] if T is not None else 0,
                   tuple(vis.shape) if vis is not None else None)
        ctx.params = (W, T if pos is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        tok, pos, slot = ctx.saved_tensors
        B, L, d, V, drop_p, seed, nT, vshape = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        # scatter-adds: straight into the gradient bucket when it already holds this window's gradient, else into zeros
        pW, pT = ctx.params
        gW = _Grad(pW, zero=True)
        side = wgrad_stream(dev)
        if side is not None and gW.direct:
            # a tied head's weight gradient went into the same buffer on the side stream at the start of backward
            torch.cuda.current_stream(dev).wait_stream(side)
        if gW.direct and not gW.acc:
            gW.buf.zero_()
        gT = None
        if pos is not None:
            gT = _Grad(pT, zero=True)
            if gT.direct and not gT.acc:
                gT.buf.zero_()
        dvis = torch.zeros(vshape, dtype=torch.float16, device=dev) if vshape is not None else None
        ops.embed_bwd(tok, pos, slot, dout, L * d, gW.buf, gT.buf if gT is not None else None, dvis, B, L, d, V, drop_p,
                      seed, 0)
        return None, None, gW.ret(), gT.ret() if gT is not None else None, dvis, None


def positional_rows(inv_freq, klen, d, clamp_len, drop_p, half_phase=False):
    """PositionalEmbedding + its dropout (transformer_xl.py:569-575) -> [klen, d] fp16, reference row order.
    half_phase: phase arithmetic in fp16 as the reference does after module.half()."""
    out = torch.empty(klen, d, dtype=torch.float16, device=inv_freq.device)
    seed = seeds.next() if drop_p > 0 else 0
    ops.posemb(out, inv_freq, klen, d, clamp_len, drop_p, seed, half_phase=half_phase)
    return out


def patch_embed(module, pixel_values, pos_sum):
    """PatchEmbeddings.forward (vision_embedding.py:65-86) on the sm_100a kernels."""
    from . import vision
    note = getattr(_sink, "note_vision", None)
    if note is not None and torch.is_grad_enabled():
        note()
    return vision.patch_embed(module, pixel_values, pos_sum)


def dropout_rows(x, p):
    """Embedding dropout (transformer_xl.py:545) on the image-patch rows of an image-caption / VQA sequence."""
    from . import vision
    return vision.DropoutFn.apply(x, p, seeds.next())
