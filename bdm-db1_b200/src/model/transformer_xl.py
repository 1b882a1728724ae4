"""DB1's Gato-style Transformer-XL with the reference's module surface, computed by hand-written sm_100a kernels.

Same class names, constructor arguments, attribute paths, state_dict keys, forward signature and return tuple as
src/model/transformer_xl.py of Shanghai-Digital-Brain-Laboratory/BDM-DB1, so a DeepSpeed-style engine or the
reference's train/eval loops can hold this module instead of theirs. Every numeric step is a launch from
libdb1_sm100.so (db1_sm100.functions); there is no PyTorch-op or CPU fallback: running it needs CUDA fp16 parameters
(`module.half().cuda()`, which is what DeepSpeed's fp16 mode does to the reference).

Kernel path restrictions (the released configuration, scripts/evaluate/evaluate_rl_1.2B.sh:70-83): post-LN
(`pre_lnorm=False`), `activation_fn="geglu"`, `dropattn=0`, no DeepNorm. Anything else raises NotImplementedError.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from db1_sm100 import functions as F_
from db1_sm100._lib import Db1Error
from src.data.input_specs import GatoInputBase, ICTaskInput, NLPTaskInput, RLTaskInput, VQATaskInput
from src.mpu import print_with_rank
from src.tokenizer.vision_embedding import VisionEmbedding

from .activations import GEGLU

ACT2FN = {"relu": nn.ReLU, "gelu": nn.GELU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "geglu": GEGLU}


def _require_kernel_tensor(t, what):
    if not t.is_cuda or t.dtype != torch.float16:
        raise Db1Error(
            "%s is %s on %s: the DB1 sm_100a path runs on CUDA fp16 parameters only (call .half().cuda()); "
            "there is no CPU / PyTorch fallback" % (what, t.dtype, t.device))


class PositionalEmbedding(nn.Module):
    """Sinusoid table of relative distances (reference :34-50), evaluated by db1_posemb.

    Phase precision follows the reference: after module.half() the reference builds pos_seq in fp16 (:569-571) and
    multiplies by the fp16-cast `inv_freq` buffer (:44), i.e. position, frequency and product are rounded to fp16
    before sin / cos - a property of everything trained / released under DeepSpeed fp16. That mode is the default here
    whenever the buffer has been cast to fp16; `phase_dtype = torch.float32` opts into exact fp32 phases (what the
    fp32 oracle and the fp32 reference compute)."""

    def __init__(self, demb):
        super().__init__()
        self.demb = demb
        inv_freq = 1 / (10000 ** (torch.arange(0.0, demb, 2.0) / demb))
        self.register_buffer("inv_freq", inv_freq)
        self.phase_dtype = None  # None: follow the buffer's dtype (reference behaviour); torch.float32: exact phases

    def rows(self, klen, clamp_len, drop_p):
        """[klen, demb] fp16: row c holds distance min(klen-1-c, clamp_len) (reference :569-575 incl. the dropout)."""
        inv = self.inv_freq
        half_phase = (inv.dtype == torch.float16) if self.phase_dtype is None else (self.phase_dtype == torch.float16)
        if inv.dtype != torch.float32:
            # module.half() casts buffers too; the kernel takes the constructor's fp32 frequencies and does the fp16
            # rounding itself (fp16(fp32 value) == the cast buffer)
            # (built once per device: a pageable host-to-device copy in every forward would synchronise the stream and
            # cannot be captured in a CUDA graph)
            cached = getattr(self, "_inv32", None)
            if cached is None or cached.device != inv.device:
                cached = (1 / (10000 ** (torch.arange(0.0, self.demb, 2.0) / self.demb))).to(inv.device)
                self._inv32 = cached
            inv = cached
        return F_.positional_rows(inv.contiguous(), klen, self.demb, clamp_len, drop_p, half_phase=half_phase)

    def forward(self, pos_seq, bsz=None):
        raise RuntimeError("PositionalEmbedding is evaluated by db1_posemb; use .rows(klen, clamp_len, drop_p)")


class RelPartialLearnableMultiHeadAttn(nn.Module):
    def __init__(self, n_head, d_model, d_head, dropout, dropatt=0, pre_lnorm=False, r_r_bias=None, r_w_bias=None,
                 layer_norm_epsilon=1e-5):
        super().__init__()
        self.n_head = n_head
        self.d_model = d_model
        self.d_head = d_head
        assert self.d_head * self.n_head == self.d_model, (self.d_head, self.n_head, self.d_model)
        self.qkv_net = nn.Linear(d_model, 3 * n_head * d_head, bias=False)
        self.drop = nn.Dropout(dropout)
        self.dropatt = nn.Dropout(dropatt)
        self.o_net = nn.Linear(n_head * d_head, d_model, bias=False)
        self.scale = 1 / (d_head ** 0.5)
        if r_r_bias is None or r_w_bias is None:
            self.r_r_bias = nn.Parameter(torch.empty(n_head, d_head))
            self.r_w_bias = nn.Parameter(torch.empty(n_head, d_head))
        else:
            self.r_r_bias = r_r_bias
            self.r_w_bias = r_w_bias
        self.r_net = nn.Linear(d_model, n_head * d_head, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=layer_norm_epsilon)
        self.pre_lnorm = pre_lnorm

    def forward(self, w, r, mem=None, attention_mask=None, head_mask=None, output_attentions=False,
                deepnorm_alpha: Optional[float] = None, window: Optional[int] = None):
        """w [B,L,d]; r [1,L,d] or [L,d] positional rows. `attention_mask` is accepted for signature compatibility; the
        kernel derives the causal / same-length mask from indices (`window` = number of keys each query may see)."""
        if self.pre_lnorm or (deepnorm_alpha is not None and deepnorm_alpha != 1.0):
            raise NotImplementedError("sm_100a path implements the released post-LN, non-DeepNorm configuration")
        if head_mask is not None or output_attentions:
            raise NotImplementedError("head_mask / output_attentions are never used by DB1's callers")
        if self.training and self.dropatt.p > 0:
            raise NotImplementedError("attention-probability dropout (dropattn) is 0 in DB1")
        _require_kernel_tensor(w, "hidden states")
        _require_kernel_tensor(self.qkv_net.weight, "qkv_net.weight")
        L = w.size(1)
        r2 = r.reshape(-1, r.size(-1))
        if attention_mask is not None and window is None:
            # reference semantics (:177): a mask that is all zeros is an error
            pass
        win = int(window) if window is not None else (1 << 30)
        if isinstance(mem, tuple) and isinstance(mem[1], F_.KVMemory):
            # decode step on cached keys / values (F_.attn_block_cached): mem = (layer index, KVMemory)
            with torch.no_grad():
                out = F_.attn_block_cached(w, mem[0], mem[1], r2, self.qkv_net.weight, self.r_net.weight, self.o_net.weight,
                                           self.r_w_bias, self.r_r_bias, self.layer_norm.weight, self.layer_norm.bias,
                                           self.n_head, self.layer_norm.eps, min(win, 1 << 30))
            return (out,)
        if mem is not None and mem.size(1) > 0:
            # memory-augmented inference (reference :124-133): no autograd, no dropout (callers run under eval())
            if torch.is_grad_enabled() and (w.requires_grad or self.qkv_net.weight.requires_grad) and self.training:
                raise NotImplementedError("training with mems is asserted off by the reference (:515-517)")
            with torch.no_grad():
                out = F_.attn_block_with_memory(w, mem.to(w.dtype), r2, self.qkv_net.weight, self.r_net.weight,
                                                self.o_net.weight, self.r_w_bias, self.r_r_bias,
                                                self.layer_norm.weight, self.layer_norm.bias, self.n_head,
                                                self.layer_norm.eps, min(win, 1 << 30))
            return (out,)
        p = self.drop.p if self.training else 0.0
        out = F_.AttnBlockFn.apply(w, r2, self.qkv_net.weight, self.r_net.weight, self.o_net.weight, self.r_w_bias,
                                   self.r_r_bias, self.layer_norm.weight, self.layer_norm.bias, self.n_head,
                                   self.layer_norm.eps, p, min(win, max(L, 1) + (1 << 20)))
        return (out,)


class PositionwiseFF(nn.Module):
    def __init__(self, d_model, d_inner, dropout, activation, pre_lnorm=False, layer_norm_epsilon=1e-5):
        super().__init__()
        self.d_model = d_model
        self.d_inner = d_inner
        self.dropout = dropout
        self.activation = activation
        if activation == "geglu":
            assert d_inner % 2 == 0
        self.CoreNet = nn.Sequential(
            nn.Linear(d_model, d_inner),
            ACT2FN[activation](),
            nn.Linear(d_inner if activation != "geglu" else d_inner // 2, d_model),
            nn.Dropout(dropout),
        )
        self.layer_norm = nn.LayerNorm(d_model, eps=layer_norm_epsilon)
        self.pre_lnorm = pre_lnorm

    def forward(self, inp, deepnorm_alpha: Optional[float] = None):
        if self.pre_lnorm or (deepnorm_alpha is not None and deepnorm_alpha != 1.0):
            raise NotImplementedError("sm_100a path implements the released post-LN, non-DeepNorm configuration")
        if self.activation != "geglu":
            raise NotImplementedError("sm_100a path implements activation_fn='geglu' (the released configuration)")
        _require_kernel_tensor(inp, "hidden states")
        _require_kernel_tensor(self.CoreNet[0].weight, "CoreNet.0.weight")
        p = self.CoreNet[3].p if self.training else 0.0
        return F_.FFBlockFn.apply(inp, self.CoreNet[0].weight, self.CoreNet[0].bias, self.CoreNet[2].weight,
                                  self.CoreNet[2].bias, self.layer_norm.weight, self.layer_norm.bias,
                                  self.layer_norm.eps, p)


class RelPartialLearnableDecoderLayer(nn.Module):
    def __init__(self, n_head, d_model, d_head, d_inner, dropout, activation, layer_norm_epsilon=1e-5, **kwargs):
        super().__init__()
        self.dec_attn = RelPartialLearnableMultiHeadAttn(n_head, d_model, d_head, dropout,
                                                         layer_norm_epsilon=layer_norm_epsilon, **kwargs)
        self.pos_ff = PositionwiseFF(d_model, d_inner, dropout, activation=activation,
                                     pre_lnorm=kwargs.get("pre_lnorm"), layer_norm_epsilon=layer_norm_epsilon)

    def forward(self, dec_inp, r, attention_mask=None, mems=None, head_mask=None, output_attentions=False,
                deepnorm_alpha: Optional[float] = None, window: Optional[int] = None):
        attn_outputs = self.dec_attn(dec_inp, r, attention_mask=attention_mask, mem=mems, head_mask=head_mask,
                                     deepnorm_alpha=deepnorm_alpha, output_attentions=output_attentions, window=window)
        ff_output = self.pos_ff(attn_outputs[0], deepnorm_alpha=deepnorm_alpha)
        return (ff_output,) + attn_outputs[1:]


class TransformerXL(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.n_embed = config.n_embed
        self.n_position = config.n_position
        self.n_layer = config.n_layer
        self.n_head = config.n_head
        self.d_head = self.n_embed // self.n_head
        self.d_model = self.n_embed
        self.d_inner = 4 * self.d_model if config.n_inner is None else config.n_inner
        self.pre_lnorm = config.pre_lnorm
        self.mem_len = config.mem_len if config.mem_len is not None else 0
        self.same_length = config.same_length
        self.clamp_len = self.n_position
        self.untie_r = config.untie_r

        self.text_vocab_size = config.text_vocab_size
        self.discrete_vocab_size = config.num_discrete_values
        self.continuous_vocab_size = config.num_continuous_bin
        self.discrete_overlap_with_text = config.overlap_with_text
        total = self.text_vocab_size + self.continuous_vocab_size + \
            (0 if self.discrete_overlap_with_text else self.discrete_vocab_size)
        self.total_vocab_size = total + 1  # + the RL separator token '|'
        self.rl_separator_token_id = total

        self.word_embedding = nn.Embedding(self.total_vocab_size, self.n_embed)
        self.pos_emb = PositionalEmbedding(self.n_embed)
        if not self.untie_r:
            self.r_w_bias = nn.Parameter(torch.empty(self.n_head, self.d_head))
            self.r_r_bias = nn.Parameter(torch.empty(self.n_head, self.d_head))
        self.vision_encoder = VisionEmbedding(config)
        self.ic_encoder = self.vision_encoder
        self.rl_local_timestep_embedding = nn.Embedding(512 + 1, self.n_embed)
        self.drop = nn.Dropout(config.embd_pdrop)
        self.h = nn.ModuleList([
            RelPartialLearnableDecoderLayer(
                self.n_head, self.d_model, self.d_head, self.d_inner, config.drop, dropatt=config.dropattn,
                activation=config.activation_fn, pre_lnorm=self.pre_lnorm,
                r_w_bias=None if self.untie_r else self.r_w_bias, r_r_bias=None if self.untie_r else self.r_r_bias,
                layer_norm_epsilon=config.layer_norm_epsilon)
            for _ in range(config.n_layer)
        ])
        self.share_input_output_embedding = config.share_input_output_embedding
        self.lm_head = None if config.share_input_output_embedding else \
            nn.Linear(config.n_embed, self.total_vocab_size, bias=False)
        self.apply(self._init_weights)
        self.use_deepnorm = config.use_deepnorm
        self.deepnorm_alpha = (2 * self.n_layer) ** 0.25 if self.use_deepnorm else None
        self.deepnorm_beta = (8 * self.n_layer) ** -0.25 if self.use_deepnorm else None
        if self.use_deepnorm:
            raise NotImplementedError("DeepNorm is off in the released configuration and not built on the sm_100a path")

    def _init_weights(self, module):
        """N(0, 0.02) for Linear / Embedding / the shared biases u, v; LayerNorm = (1, 0) (reference :456-468)."""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        else:
            if hasattr(module, "r_r_bias"):
                module.r_r_bias.data.normal_(mean=0.0, std=0.02)
            if hasattr(module, "r_w_bias"):
                module.r_w_bias.data.normal_(mean=0.0, std=0.02)

    def init_mem(self, batch_size, kv_cache=False):
        """Reference :470-485. kv_cache=True returns a db1_sm100.functions.KVMemory instead of the list of hidden-state
        tensors: forward(..., mems=<KVMemory>) then runs the decode step on cached keys / values (same logits, no
        re-projection of the memory) and returns the same object, updated in place, as `new_mems`."""
        if self.mem_len > 0 and kv_cache:
            param = next(self.parameters())
            return F_.KVMemory(self.n_layer, batch_size, self.mem_len, self.n_embed, param.device)
        if self.mem_len > 0:
            param = next(self.parameters())
            return [torch.zeros(batch_size, self.mem_len, self.n_embed, dtype=param.dtype, device=param.device)
                    for _ in range(self.n_layer)]
        return None

    def _update_mem(self, hiddens, mems, mlen, qlen):
        if mems is None:
            return None
        assert len(hiddens) == len(mems), "len(hids) != len(mems)"
        with torch.no_grad():
            end_idx = mlen + max(0, qlen)
            beg_idx = max(0, end_idx - self.mem_len)
            return [torch.cat([mems[i], hiddens[i]], dim=1)[:, beg_idx:end_idx] for i in range(len(hiddens))]

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, tasks_input: List[GatoInputBase], compute_loss: bool = True, mems=None):
        assert not (compute_loss and mems is not None), "During training, Gato does not use memory mechanism."
        embs, masks, labels = [], [], []
        p_emb = self.drop.p if self.training else 0.0
        for task in tasks_input:
            if isinstance(task, RLTaskInput):
                fn = self._forward_rl
            elif isinstance(task, NLPTaskInput):
                fn = self._forward_nlp
            elif isinstance(task, ICTaskInput):
                fn = self._forward_ic
            elif isinstance(task, VQATaskInput):
                fn = self._forward_vqa
            else:
                raise TypeError("unknown task input type %r" % type(task))
            e, m, _a, l = fn(task, p_emb)
            embs.append(e)
            masks.append(m)
            labels.append(l)
        loss_masks = torch.cat(masks, dim=0) if compute_loss else None
        labels_all = torch.cat(labels, dim=0).long() if compute_loss else None
        hidden = embs[0] if len(embs) == 1 else torch.cat(embs, dim=0)

        qlen = hidden.size(1)
        cached = isinstance(mems, F_.KVMemory)
        if cached and (self.training or not self.same_length or mems.batch_size != hidden.size(0) or qlen > mems.cap):
            raise ValueError("KVMemory decode needs eval mode, same_length, qlen <= mem_len and the memory's batch size")
        mlen = (mems.cap if cached else mems[0].size(1)) if mems is not None else 0
        klen = mlen + qlen
        # same_length: every query sees exactly mem_len keys once klen exceeds mem_len (reference :551-562);
        # otherwise plain causal. mem_len == 0 with same_length masks everything -> the reference raises ValueError.
        if self.same_length and klen > self.mem_len:
            if self.mem_len <= 0:
                raise ValueError("attention mask removes every key (same_length with mem_len == 0)")
            window = self.mem_len
        else:
            window = 1 << 30
        pos_rows = self.pos_emb.rows(klen, self.clamp_len, self.drop.p if self.training else 0.0)

        if cached and F_.decode_step_fused_applies(self, hidden.size(0) * qlen, hidden.size(2)):
            # few new rows: all layers + the head with every LayerNorm applied inside the consuming few-row GEMM
            with torch.no_grad():
                lm_logits = F_.decode_step_fused(self, hidden, mems, pos_rows, window)
            mems.advance(qlen)
            return (lm_logits, None, mems)

        hids = []
        for li, block in enumerate(self.h):
            if mems is not None and not cached:
                hids.append(hidden)
            layer_mem = None if mems is None else ((li, mems) if cached else mems[li])
            hidden = block(hidden, pos_rows, attention_mask=None, mems=layer_mem,
                           head_mask=None, output_attentions=False, deepnorm_alpha=self.deepnorm_alpha, window=window)[0]
        if cached:
            mems.advance(qlen)  # every layer appended its qlen new rows over the oldest slots
            new_mems = mems
        else:
            new_mems = self._update_mem(hids, mems, mlen, qlen) if mems is not None else None

        W = self.word_embedding.weight if self.share_input_output_embedding else self.lm_head.weight
        if compute_loss:
            lm_logits, loss = F_.HeadLossFn.apply(hidden, W, labels_all, loss_masks)
        else:
            with torch.no_grad():
                lm_logits = F_.head_logits(hidden, W)
            loss = None
        if new_mems is not None:
            return (lm_logits, loss, new_mems)  # reference :615-619
        return (lm_logits, loss)

    # ------------------------------------------------------------------------------------------------ task embeddings
    def _forward_rl(self, rl_input: RLTaskInput, p_emb=0.0):
        tok = rl_input.tensor_seq
        label = rl_input.label
        vis = None
        if rl_input.vision_seq is not None:
            img = rl_input.vision_seq
            bsz = tok.size(0)
            vis = self.vision_encoder(img.view(-1, *img.shape[-3:])).reshape(bsz, -1, self.n_embed)
            if label is not None:
                label[label == -1] = 0  # in-place, as the reference does (:645)
        emb = F_.EmbedFn.apply(tok, rl_input.position_id, self.word_embedding.weight,
                               self.rl_local_timestep_embedding.weight, vis, p_emb)
        return emb, rl_input.loss_mask, rl_input.attention_mask, label

    def _forward_nlp(self, nlp_input: NLPTaskInput, p_emb=0.0):
        emb = F_.EmbedFn.apply(nlp_input.text_seq, None, self.word_embedding.weight, None, None, p_emb)
        return emb, nlp_input.loss_mask, nlp_input.attention_mask, nlp_input.label

    def _cat_prompt_image_text(self, prompt_seq, img_seq, text_seq, p_emb):
        W = self.word_embedding.weight
        prompt = F_.EmbedFn.apply(prompt_seq, None, W, None, None, p_emb)
        vis = self.ic_encoder(img_seq)
        if p_emb > 0:
            vis = F_.dropout_rows(vis, p_emb)
        text = F_.EmbedFn.apply(text_seq, None, W, None, None, p_emb)
        return torch.cat([prompt, vis, text], dim=1)

    def _forward_ic(self, ic_input: ICTaskInput, p_emb=0.0):
        enc = self._cat_prompt_image_text(ic_input.prompt_seq, ic_input.img_seq, ic_input.text_seq, p_emb)
        return enc, ic_input.loss_mask, ic_input.attention_mask, ic_input.label

    def _forward_vqa(self, vqa_input: VQATaskInput, p_emb=0.0):
        enc = self._cat_prompt_image_text(vqa_input.prompt_seq, vqa_input.img_seq, vqa_input.text_seq, p_emb)
        return enc, vqa_input.loss_mask, vqa_input.attention_mask, vqa_input.label
