"""Data-parallel training engine with the protocol DB1's loops expect from DeepSpeed
(src/train_utils/train.py:210-243, src/checkpointing.py:17-22, src/evaluation/evaluate_rl.py:508-512):

    engine(inputs) -> (logits, loss);  engine.backward(loss);  engine.step();
    engine.gradient_accumulation_steps();  engine.train()/eval();  engine.device;
    engine.save_checkpoint(dir, client_state=..., tag=...);  engine.load_checkpoint(dir, tag) -> (path, client_state)

What DeepSpeed 0.6.7 does after backward, serially (one fp16 gradient all-reduce over the data-parallel group at every
accumulation boundary), is done here overlapped with backward:
  * every parameter's .grad is a view into a flat per-bucket fp16 buffer (one bucket per decoder layer + one for the
    embeddings / vision encoder / shared biases), so autograd accumulates straight into communication buffers;
  * when the last gradient of a bucket has been accumulated (post-accumulate-grad hooks), an event is recorded on the
    compute stream and the bucket's NCCL all-reduce (average) is enqueued on a dedicated high-priority stream — layer
    l's reduction runs under layer l-1's backward kernels;
  * parameters that received no gradient this step (e.g. the vision encoder on a text-only rank) keep their zero-filled
    views, so every rank issues the same collectives in the same order;
  * only the boundary micro-step of an accumulation window communicates.
The same code runs over gloo on CPU tensors (tests) — the overlap machinery is skipped there.
"""
import os
from collections import OrderedDict

import torch
import torch.distributed as dist


def _default_bucket_of(name):
    """Bucket key for a parameter name: 'h.l' = decoder layer l; 'emb' = the (tied) word embedding; 'vision' = the image
    patch embedder; 'rest' = everything else (shared u / v, timestep embedding, an untied head).
    DB1_BUCKET_SPLIT_FF=1 splits a layer into 'h.l.ff' (the two feed-forward matrices, complete half a layer earlier in
    backward) and 'h.l.attn': measured on 8 GPUs (profiles/README.md) the 24 extra collectives cost more than the finer
    overlap gains (38.6 against 36.6 ms per step), so one bucket per layer is the default."""
    parts = name.split(".")
    if len(parts) > 2 and parts[0] == "h" and parts[1].isdigit():
        if os.environ.get("DB1_BUCKET_SPLIT_FF", "0") == "1":
            ff = parts[2] == "pos_ff" and "CoreNet" in parts and parts[-1] == "weight"
            return "h.%s.%s" % (parts[1], "ff" if ff else "attn")
        return "h.%s" % parts[1]
    if name == "word_embedding.weight":
        return "emb"
    if parts[0] == "vision_encoder":
        return "vision"
    return "rest"


# buckets whose parameters are used more than once per forward (or whose use is data dependent): they are never launched
# from inside backward by the completion count
_LATE = ("rest", "emb", "vision")


class _Bucket:
    def __init__(self, key, params, names):
        self.key = key
        self.params = params
        self.names = names
        self.flat = None
        self.pending = 0
        self.work = None
        self.event = None


class DB1Engine:
    def __init__(self, model, optimizer=None, lr_scheduler=None, mpu=None, gradient_accumulation_steps=1,
                 loss_scale=4096.0, clip_grad=0.0, bucket_of=_default_bucket_of, overlap_comm=True, direct_grads=True,
                 fused_adam=None, loss_scale_window=1000, min_loss_scale=1.0):
        self.module = model
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.mpu = mpu
        self._ga = int(gradient_accumulation_steps)
        self.loss_scale = float(loss_scale)
        self.clip_grad = float(clip_grad)
        self.micro_steps = 0
        self.global_steps = 0
        self.enable_backward_allreduce = True
        self._group = None
        self._world = 1
        if dist.is_available() and dist.is_initialized():
            if mpu is not None and not mpu.is_unitialized():
                self._group = mpu.get_data_parallel_group()
                self._world = mpu.get_data_parallel_world_size()
            else:
                self._group = dist.group.WORLD
                self._world = dist.get_world_size()
        p0 = next(model.parameters())
        self.device = p0.device
        self._cuda = p0.is_cuda
        self._overlap = bool(overlap_comm) and self._cuda and self._world > 1
        self._comm_stream = torch.cuda.Stream(device=self.device, priority=-1) if self._overlap else None
        # SMs left to the collective while it overlaps backward: the persistent kernels' grids are sized to the rest
        # (include/db1_sm100.h:db1_set_sm_budget). NCCL_MAX_CTAS (set before the communicator is created) bounds how many
        # CTAs NCCL takes; DB1_COMM_SMS overrides the reservation (0 disables it).
        self._comm_sms = int(os.environ.get("DB1_COMM_SMS", os.environ.get("NCCL_MAX_CTAS", "0") or 0)) if self._overlap else 0
        self._build_buckets(bucket_of)
        self._hooks = []
        if self._world > 1:
            self._install_hooks()
        # fused optimizer (db1_adam_step over the flat buckets): fp16 parameters become views of a flat buffer per bucket,
        # with fp32 master weights and Adam moments beside it
        self._fused = None
        self._good_steps = 0
        self._scale_window = int(loss_scale_window)
        self._min_scale = float(min_loss_scale)
        if fused_adam is not None:
            if not self._cuda:
                raise ValueError("fused_adam needs CUDA fp16 parameters")
            self._setup_fused_adam(dict(fused_adam))
        # gradient sink (db1_sm100.functions): weight-gradient kernels write into the bucket views directly
        self._written = set()    # ids of parameters whose bucket view already holds this window's gradient
        self._sink_seen = set()  # ids of parameters that have ever been written through the sink
        self._param_by_id = {id(p): p for b in self.buckets for p in b.params}
        self._sink_on = bool((self._cuda and direct_grads) or direct_grads == "force")
        # weight-gradient GEMMs may run on a side stream (functions.wgrad_stream): only meaningful with the sink, where
        # nothing but this engine reads the gradients before the end of backward
        self.wgrad_side_stream = bool(self._cuda and self._sink_on)
        if self._sink_on:
            from . import functions
            functions.set_grad_sink(self)
        self._setup_emb_split()

    # ------------------------------------------------------------------------------------------ buckets
    def _build_buckets(self, bucket_of):
        groups = OrderedDict()
        seen = set()
        for name, p in self.module.named_parameters():  # shared parameters appear once
            if not p.requires_grad or id(p) in seen:
                continue
            seen.add(id(p))
            groups.setdefault(bucket_of(name), ([], []))
            groups[bucket_of(name)][0].append(p)
            groups[bucket_of(name)][1].append(name)
        self.buckets = []
        self._bucket_of_param = {}
        for key, (params, names) in groups.items():
            b = _Bucket(key, params, names)
            n = sum((p.numel() + 7) // 8 * 8 for p in params)  # keep every view 16-byte aligned
            b.flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            off = 0
            for p in params:
                p.grad = b.flat[off:off + p.numel()].view_as(p)
                off += (p.numel() + 7) // 8 * 8
                self._bucket_of_param[id(p)] = b
            self.buckets.append(b)

    # ------------------------------------------------------------------------------------------ tied-embedding split
    # The tied embedding's gradient has a dense part (the head's weight gradient, complete right after backward starts)
    # and a sparse part (the embedding scatter, the very last kernel of backward). All-reduce is linear, so the dense part
    # is reduced at once - hidden under the whole backward - and the scatter goes to a zero-kept side buffer of which only
    # the row range [lo, hi) that any rank's tokens touched this window is reduced afterwards and added. The range is
    # agreed by a 3-element MAX all-reduce enqueued before the first bucket of the step (so every rank slices the same
    # rows); its third element says whether any rank fed the image patch embedder, else the (all-zero) 'vision' bucket is
    # not communicated at all. What is left exposed after backward is then the first layer's attention bucket plus a few MB.
    def _setup_emb_split(self):
        self._emb = None
        self._vision = None
        for b in self.buckets:
            if b.key == "emb" and len(b.params) == 1 and b.params[0].dim() == 2:
                self._emb = b
            if b.key == "vision":
                self._vision = b
        on = os.environ.get("DB1_EMB_SPLIT", "1") != "0"
        if not (self._sink_on and self._world > 1 and on):
            self._emb = None
        self._emb_phase = 0
        self._split_active = False
        self._rng = None
        if self._world > 1 and self._sink_on and on and (self._emb is not None or self._vision is not None):
            dev = self.device
            self._rng_init = torch.tensor([-(1 << 40), -1, 0], dtype=torch.int64, device=dev)  # [-min token, max token, vision]
            self._rng = self._rng_init.clone()
            self._rng_host = torch.zeros(3, dtype=torch.int64, pin_memory=self._cuda)
            self._rng_ev = torch.cuda.Event() if self._cuda else None
        if self._emb is not None:
            p = self._emb.params[0]
            self._scatter = torch.zeros_like(p)  # kept all-zero between uses

    def note_tokens(self, tok):
        """Called by the embedding forward with the token ids it looks up (device tensor; -1 = image slot)."""
        if self._rng is None or self._emb is None:
            return
        mn, mx = torch.aminmax(tok)
        torch.maximum(self._rng[:2], torch.stack([-mn, mx]).to(torch.int64), out=self._rng[:2])

    def note_vision(self):
        """Called by the image patch embedder's forward: its parameters receive gradients this window."""
        if self._rng is not None:
            self._rng[2:3].fill_(1)

    def _agree_begin(self):
        """Enqueue the MAX all-reduce of (token range, vision flag) ahead of this step's bucket collectives."""
        if self._overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                dist.all_reduce(self._rng, op=dist.ReduceOp.MAX, group=self._group)
                self._rng_host.copy_(self._rng, non_blocking=True)
                self._rng_ev.record(self._comm_stream)
        else:
            dist.all_reduce(self._rng, op=dist.ReduceOp.MAX, group=self._group)

    def _agree_end(self):
        """(lo, hi, vision_used) as host integers, identical on every rank; re-arms the device vector."""
        if self._overlap:
            self._rng_ev.synchronize()  # enqueued before backward's first kernel: long complete, no pipeline bubble
            nlo, hi, vis = [int(x) for x in self._rng_host.tolist()]
            with torch.cuda.stream(self._comm_stream):
                self._rng.copy_(self._rng_init)
        else:
            nlo, hi, vis = [int(x) for x in self._rng.tolist()]
            self._rng.copy_(self._rng_init)
        return max(0, -nlo), hi + 1, vis != 0

    def _finish_emb_split(self, lo, hi):
        """Reduce rows [lo, hi) of the scatter buffer, add them to the (already reduced) dense part, re-zero them."""
        p = self._emb.params[0]
        hi = min(hi, p.shape[0])
        if hi <= lo:
            return
        rows = self._scatter[lo:hi]
        if self._overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                dist.all_reduce(rows, op=dist.ReduceOp.AVG, group=self._group)
                self._emb.work.wait()  # the dense part (enqueued long before)
                p.grad[lo:hi].add_(rows)
                rows.zero_()
        else:
            if dist.get_backend(self._group) == "gloo":
                dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=self._group)
                rows.mul_(1.0 / self._world)
            else:
                dist.all_reduce(rows, op=dist.ReduceOp.AVG, group=self._group)
            self._emb.work.wait()
            if getattr(self._emb, "post_scale", None):
                self._emb.flat.mul_(self._emb.post_scale)
                self._emb.post_scale = None
            p.grad[lo:hi].add_(rows)
            rows.zero_()

    def _setup_fused_adam(self, cfg):
        betas = cfg.get("betas", (0.9, 0.999))
        self._fused = FusedAdamState(lr=float(cfg.get("lr", 1e-4)), beta1=float(betas[0]), beta2=float(betas[1]),
                                     eps=float(cfg.get("eps", 1e-8)), weight_decay=float(cfg.get("weight_decay", 0.0)),
                                     adamw=bool(cfg.get("adamw", True)))
        if self.optimizer is None:
            self.optimizer = self._fused  # exposes param_groups[0]["lr"] to LR schedulers (OptimizerParamScheduler)
        for b in self.buckets:
            b.pflat = torch.zeros_like(b.flat)
            off = 0
            with torch.no_grad():
                for p in b.params:
                    n = p.numel()
                    b.pflat[off:off + n].copy_(p.data.reshape(-1))
                    p.data = b.pflat[off:off + n].view_as(p)
                    off += (n + 7) // 8 * 8
            b.master = b.pflat.float()
            b.m = torch.zeros_like(b.master)
            b.v = torch.zeros_like(b.master)
        dev = self.device
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gcoef = torch.zeros(2, dtype=torch.float32, device=dev)
        self._oflag = torch.zeros(1, dtype=torch.int32, device=dev)

    def _install_hooks(self):
        for b in self.buckets:
            for p in b.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def _make_hook(self, bucket):
        def hook(param):
            if not self._is_boundary() or not self.enable_backward_allreduce:
                return
            if id(param) in self._written:
                # autograd runs AccumulateGrad (and this hook) even when the Function returned None because its kernel
                # already wrote the bucket view; those parameters are accounted for by done()
                return
            if bucket.key in _LATE:
                # launched after backward in bucket order: whether these complete inside backward depends on the rank's
                # data (an unused patch embedder never does), and every rank must issue its collectives in one order
                return
            bucket.pending -= 1
            if bucket.pending == 0:
                self._launch_allreduce(bucket)
        return hook

    # ------------------------------------------------------------------------------------------ gradient sink protocol
    def target(self, param):
        """(bucket view, accumulate) for a parameter this engine owns, else None. The first write of an accumulation
        window overwrites (so the buckets never need a zero-fill), later ones accumulate."""
        pid = id(param)
        if pid not in self._param_by_id:
            return None
        if self._split_active and self._emb_phase == 1 and param is self._emb.params[0]:
            return self._scatter, True  # the dense part is already on the wire: scatter-adds go to the side buffer
        acc = pid in self._written
        self._written.add(pid)
        return param.grad, acc

    def done(self, param):
        """A kernel finished writing one contribution to param's gradient. Parameters of a decoder-layer bucket are used
        once per forward, so this completes them; the shared 'rest' bucket (embeddings, u/v, vision) is launched after
        backward returns."""
        b = self._bucket_of_param.get(id(param))
        if b is None or self._world == 1:
            return
        if not self._is_boundary() or not self.enable_backward_allreduce:
            return
        if b is self._emb and self._split_active:
            if self._emb_phase == 0:  # first complete contribution (the head's dense gradient): reduce it now
                self._emb_phase = 1
                self._launch_allreduce(b)
            return
        if b.key in _LATE:
            return
        b.pending -= 1
        if b.pending == 0:
            self._launch_allreduce(b)

    def _side_stream(self):
        if not (self._sink_on and self._cuda):
            return None
        from . import functions
        return functions.wgrad_stream(self.device)

    def _launch_allreduce(self, bucket):
        if bucket.work is not None or self._world == 1:
            return
        if self._overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            side = self._side_stream()
            ev2 = None
            if side is not None:  # the bucket's weight gradients were written there
                ev2 = torch.cuda.Event()
                ev2.record(side)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                if ev2 is not None:
                    self._comm_stream.wait_event(ev2)
                bucket.work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG, group=self._group, async_op=True)
        else:
            side = self._side_stream()
            if side is not None:
                torch.cuda.current_stream(self.device).wait_stream(side)
            if dist.get_backend(self._group) == "gloo":
                bucket.work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=self._group, async_op=True)
                bucket.post_scale = 1.0 / self._world
            else:
                bucket.work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.AVG, group=self._group, async_op=True)

    # ------------------------------------------------------------------------------------------ protocol
    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    def gradient_accumulation_steps(self):
        return self._ga

    def _is_boundary(self):
        return (self.micro_steps + 1) % self._ga == 0

    def is_gradient_accumulation_boundary(self):
        return self._is_boundary()

    def zero_grad(self):
        """Start of an accumulation window. Gradients that arrive through the sink overwrite their views, so only
        parameters that come through autograd's own accumulation (never seen by the sink) are zero-filled."""
        self._written.clear()
        for b in self.buckets:
            unseen = [p for p in b.params if id(p) not in self._sink_seen]
            if len(unseen) == len(b.params):
                b.flat.zero_()
            else:
                for p in unseen:
                    p.grad.zero_()

    def _zero_unwritten(self):
        """Sink-fed parameters that received nothing in this window (e.g. unused this step) must read as zero."""
        for pid in self._sink_seen - self._written:
            self._param_by_id[pid].grad.zero_()
        self._sink_seen |= self._written

    def backward(self, loss):
        """Scale, back-propagate; on the boundary micro-step the bucket all-reduces start as their gradients complete."""
        if self.micro_steps % self._ga == 0:
            self.zero_grad()
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
            b.post_scale = None
        scaled = loss * (self.loss_scale / self._ga)
        comm_step = self._world > 1 and self._is_boundary() and self.enable_backward_allreduce
        self._split_active = bool(comm_step and self._emb is not None)
        self._emb_phase = 0
        agree = comm_step and self._rng is not None
        if agree:
            self._agree_begin()
        reserve = self._comm_sms if (self._overlap and self._is_boundary() and self.enable_backward_allreduce) else 0
        if reserve > 0:
            from . import _lib
            _lib.lib().db1_set_sm_budget(max(8, _lib.lib().db1_sm_count() - reserve))
        if self._sink_on and loss.is_cuda:
            from . import functions
            functions.begin_backward(loss.device)  # one zero-filled arena for the blocks' small fp32 accumulators
            try:
                scaled.backward()
            finally:
                functions.end_backward()
        else:
            scaled.backward()
        if reserve > 0:
            _lib.lib().db1_set_sm_budget(0)
        side = self._side_stream()
        if side is not None:  # weight gradients written on the side stream are part of this backward
            torch.cuda.current_stream(self.device).wait_stream(side)
        if self._is_boundary():
            self._zero_unwritten()
        else:
            self._sink_seen |= self._written
        if self._world > 1 and self._is_boundary() and self.enable_backward_allreduce:
            lo, hi, vision_used = self._agree_end() if agree else (0, 0, True)
            if self._split_active and self._emb_phase == 1:
                self._finish_emb_split(lo, hi)
            # buckets whose parameters got no gradient at all this step still take part (zero contribution) - except
            # the patch embedder's when no rank used it in this window (all zeros everywhere)
            for b in self.buckets:
                if b.work is None and not (b is self._vision and agree and not vision_used):
                    self._launch_allreduce(b)
            self._finish_allreduce()
            self._split_active = False
        self.micro_steps += 1
        return loss

    def _finish_allreduce(self):
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()  # NCCL: makes the current stream wait for the collective; gloo: blocks the host
                if getattr(b, "post_scale", None):
                    b.flat.mul_(b.post_scale)
        if self._overlap:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def step(self):
        """Optimizer step on accumulation boundaries: unscale, optional global-norm clip, skip on overflow."""
        if self.micro_steps % self._ga != 0:
            return
        self.global_steps += 1
        if self._fused is not None:
            return self._fused_step()
        if self.optimizer is None:
            return
        inv = 1.0 / self.loss_scale
        flats = [b.flat for b in self.buckets]
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        for f in flats:
            total += f.float().pow(2).sum()
        norm = total.sqrt() * inv
        if not torch.isfinite(norm):
            self.loss_scale = max(self.loss_scale / 2, self._min_scale)
            self._good_steps = 0
            self.skipped_steps = getattr(self, "skipped_steps", 0) + 1
            return
        coef = inv
        if self.clip_grad > 0:
            coef = inv * min(1.0, self.clip_grad / (norm.item() + 1e-6))
        for f in flats:
            f.mul_(coef)
        self.optimizer.step()
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        self._good_steps += 1  # same growth window as the fused path (DeepSpeed's dynamic loss scaler)
        if self._scale_window > 0 and self._good_steps % self._scale_window == 0:
            self.loss_scale *= 2.0

    def _fused_step(self):
        """Unscale, global-norm clip, overflow check and Adam(W) on fp32 masters: two passes over the flat buckets
        (db1_grad_sumsq, db1_adam_step), one 4-byte host read for the overflow flag (as DeepSpeed's has_overflow)."""
        from . import ops
        st = self._fused
        self._sumsq.zero_()
        for b in self.buckets:
            ops.grad_sumsq(b.flat, self._sumsq)
        ops.clip_coef(self._sumsq, 1.0 / self.loss_scale, self.clip_grad, self._gcoef, self._oflag)
        if int(self._oflag.item()) != 0:  # inf / nan in the gradients: skip the step, halve the loss scale
            self.loss_scale = max(self.loss_scale / 2.0, self._min_scale)
            self._good_steps = 0
            self.skipped_steps = getattr(self, "skipped_steps", 0) + 1
            return False
        st.step += 1
        lr = float(st.param_groups[0]["lr"])
        wd = float(st.param_groups[0].get("weight_decay", st.weight_decay))
        for b in self.buckets:
            ops.adam_step(b.flat, b.pflat, b.master, b.m, b.v, self._gcoef, lr, st.beta1, st.beta2, st.eps, wd, st.step,
                          st.adamw)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        self._good_steps += 1
        if self._scale_window > 0 and self._good_steps % self._scale_window == 0:
            self.loss_scale *= 2.0
        return True

    def grad_norm(self):
        """Unscaled global gradient norm of the last fused step (device scalar)."""
        return self._gcoef[1]

    # ------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, save_dir, tag=None, client_state=None):
        tag = tag if tag is not None else "global_step%d" % self.global_steps
        path = os.path.join(save_dir, str(tag))
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if rank == 0:
            os.makedirs(path, exist_ok=True)
            # key set of DeepSpeed 0.6.7's mp_rank_00_model_states.pt (engine.py:_save_checkpoint) + this engine's extras
            state = {"module": self.module.state_dict(), "buffer_names": [n for n, _ in self.module.named_buffers()],
                     "optimizer": None, "lr_scheduler": None, "sparse_tensor_module_names": [],
                     "skipped_steps": getattr(self, "skipped_steps", 0), "global_steps": self.global_steps,
                     "global_samples": self.global_steps * self._ga * self._world, "dp_world_size": self._world,
                     "mp_world_size": 1, "ds_config": None, "ds_version": "db1_sm100",
                     "micro_steps": self.micro_steps, "loss_scale": self.loss_scale}
            if self.lr_scheduler is not None and hasattr(self.lr_scheduler, "state_dict"):
                state["lr_scheduler"] = self.lr_scheduler.state_dict()
            if self._fused is not None:
                state["optimizer"] = {"db1_engine": True, "fused_adam": self._fused.state_dict(),
                                      "buckets": [{"key": b.key, "master": b.master.cpu(), "m": b.m.cpu(), "v": b.v.cpu()}
                                                  for b in self.buckets]}
            elif self.optimizer is not None:
                state["optimizer"] = {"db1_engine": True, "torch_optimizer": self.optimizer.state_dict()}
            state.update(client_state or {})
            torch.save(state, os.path.join(path, "mp_rank_00_model_states.pt"))
            with open(os.path.join(save_dir, "latest"), "w") as f:
                f.write(str(tag))
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=self._group)
        return True

    def load_checkpoint(self, load_dir, tag=None, load_optimizer_states=True):
        """Reads `<dir>/<tag>/mp_rank_00_model_states.pt` (tag from `<dir>/latest` when omitted) - the file DeepSpeed 0.6.7
        writes (src/checkpointing.py:17-22, evaluate_rl.py:509-511) as well as this engine's own. Module weights always
        load; optimizer state is restored only when it was written by this engine (marker `db1_engine`): DeepSpeed stores
        None or an FP16_Optimizer dict there, in which case the fp32 masters are rebuilt from the loaded weights and the
        moments start from zero (with a warning). Everything that is not engine bookkeeping comes back as client state."""
        import warnings
        if tag is None:
            with open(os.path.join(load_dir, "latest")) as f:
                tag = f.read().strip()
        path = os.path.join(load_dir, str(tag), "mp_rank_00_model_states.pt")
        state = torch.load(path, map_location="cpu", weights_only=False)
        self.module.load_state_dict(state["module"], strict=True)
        self.global_steps = int(state.get("global_steps", 0) or 0)
        self.micro_steps = int(state.get("micro_steps", self.global_steps * self._ga) or 0)
        opt = state.get("optimizer")
        ours = isinstance(opt, dict) and opt.get("db1_engine") is True
        if "loss_scale" in state:
            self.loss_scale = float(state["loss_scale"])
        elif isinstance(opt, dict) and "cur_scale" in opt:  # DeepSpeed FP16_Optimizer
            self.loss_scale = float(opt["cur_scale"])
        if self._fused is not None:
            # parameters are views of the flat buffers: load_state_dict above copied into them in place
            restored = False
            if load_optimizer_states and ours and "buckets" in opt and len(opt["buckets"]) == len(self.buckets):
                self._fused.load_state_dict(opt["fused_adam"])
                for b, sb in zip(self.buckets, opt["buckets"]):
                    b.master.copy_(sb["master"]); b.m.copy_(sb["m"]); b.v.copy_(sb["v"])
                restored = True
            if not restored:
                if load_optimizer_states and opt is not None:
                    warnings.warn("checkpoint optimizer state was not written by DB1Engine (DeepSpeed layout?): fp32 master "
                                  "weights rebuilt from the loaded fp16 weights, Adam moments reset")
                for b in self.buckets:
                    b.master.copy_(b.pflat.float())
                    b.m.zero_(); b.v.zero_()
        elif load_optimizer_states and self.optimizer is not None and opt is not None:
            if ours and "torch_optimizer" in opt:
                self.optimizer.load_state_dict(opt["torch_optimizer"])
            elif isinstance(opt, dict) and "state" in opt and "param_groups" in opt:
                self.optimizer.load_state_dict(opt)
            else:
                warnings.warn("checkpoint optimizer state is not a torch optimizer state_dict (DeepSpeed FP16_Optimizer "
                              "layout?): optimizer state not restored")
        known = {"module", "global_steps", "micro_steps", "loss_scale", "optimizer", "lr_scheduler", "buffer_names",
                 "sparse_tensor_module_names", "skipped_steps", "global_samples", "dp_world_size", "mp_world_size",
                 "ds_config", "ds_version", "csr_tensor_module_names", "param_shapes"}
        if self.lr_scheduler is not None and state.get("lr_scheduler") is not None and hasattr(self.lr_scheduler, "load_state_dict"):
            self.lr_scheduler.load_state_dict(state["lr_scheduler"])
        self.skipped_steps = int(state.get("skipped_steps", getattr(self, "skipped_steps", 0)) or 0)
        return path, {k: v for k, v in state.items() if k not in known}


class FusedAdamState:
    """Hyper-parameters + step counter of the fused optimizer; looks enough like a torch optimizer (param_groups,
    state_dict) for the reference's OptimizerParamScheduler (optimizer_param_scheduler.py:100-142) to drive lr / wd."""

    def __init__(self, lr, beta1, beta2, eps, weight_decay, adamw):
        self.beta1, self.beta2, self.eps, self.weight_decay, self.adamw = beta1, beta2, eps, weight_decay, adamw
        self.step = 0
        self.param_groups = [{"lr": lr, "weight_decay": weight_decay}]

    def state_dict(self):
        return {"step": self.step, "param_groups": [dict(g) for g in self.param_groups], "beta1": self.beta1,
                "beta2": self.beta2, "eps": self.eps, "adamw": self.adamw}

    def load_state_dict(self, sd):
        self.step = sd["step"]
        self.param_groups = [dict(g) for g in sd["param_groups"]]
        self.beta1, self.beta2, self.eps, self.adamw = sd["beta1"], sd["beta2"], sd["eps"], sd["adamw"]


def initialize(args=None, model=None, optimizer=None, lr_scheduler=None, mpu=None, **kw):
    """deepspeed.initialize-shaped constructor: returns (engine, optimizer, None, lr_scheduler)."""
    ga = kw.pop("gradient_accumulation_steps", getattr(args, "gradient_accumulation_steps", 1) if args else 1)
    eng = DB1Engine(model, optimizer=optimizer, lr_scheduler=lr_scheduler, mpu=mpu, gradient_accumulation_steps=ga, **kw)
    return eng, optimizer, None, lr_scheduler
