/*
 * db1_sm100.h — C ABI of libdb1_sm100.so: the B200 (sm_100a) kernels behind DB1's Transformer-XL forward/backward.
 *
 * The reference (Shanghai-Digital-Brain-Laboratory/BDM-DB1) has no FFI of its own for this path: every numeric op is
 * a PyTorch library call made from src/model/transformer_xl.py and src/tokenizer/vision_embedding.py. Each entry
 * point below names the reference call site(s) it replaces. The Python side (bdm-db1_b200/db1_sm100/_lib.py) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the name says host; no torch types cross this boundary
 *   - all 16-bit tensors are IEEE fp16 ("half"); reductions / statistics are fp32
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises, allocates or frees device memory
 *   - return value: 0 = ok, < 0 = argument / environment error, > 0 = cudaError_t; db1_last_error() (thread-local)
 *     describes the last non-zero return
 */
#ifndef DB1_SM100_H_
#define DB1_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* db1_last_error(void);
/* ABI version of this header; bumped whenever a struct layout changes (3: db1_gemm_desc.b_static and ln_* appended). */
int db1_abi_version(void);
/* Number of SMs the persistent kernels may cover (0 = all). Lower it while a long-running communication kernel occupies
 * SMs (DB1Engine does during backward at world size > 1: physical SMs minus NCCL's CTAs), so that a grid never needs a
 * second wave for CTAs that found no free SM. Process-wide; takes effect at the next launch. */
int db1_set_sm_budget(int n_sms);
int db1_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma, TMEM accumulators, TMA-fed smem ring; persistent, one CTA per SM).
 *
 *   for every batch index (z1, z2):   C[M,N] (+)= epilogue( alpha * A[M,K] . B[N,K]^T )
 *
 * Replaces: nn.Linear / F.linear / torch.einsum and their autograd backward at
 *   transformer_xl.py:138 (qkv_net), :139 (r_net), :228 (o_net), :265-268 (CoreNet), :595 (tied head),
 *   vision_embedding.py:53-55 (16x16/stride-16 projection conv == GEMM),
 *   and, in the attention backward, the einsum adjoints of :163-170 and :220.
 * ------------------------------------------------------------------------------------------------------------------ */
enum {
  DB1_EPI_PLAIN = 0,  /* C = [resid +] dropout(alpha*acc [+ bias]) [+ C]                                              */
  DB1_EPI_QKV = 1,    /* N = 3*d_model: columns < d_model are written twice (+u, +v): C row = [q+u | q+v | k | v]       */
  DB1_EPI_GEGLU = 2,  /* N = 2F, tile pairs column n with n+F: H = [a|g] + bias (pre-activation), C = a * gelu_erf(g)   */
  DB1_EPI_DGEGLU = 3, /* N = F, acc = dY: reads H = [a|g]; C = [dY*gelu(g) | dY*a*gelu'(g)] (width 2F)                 */
  DB1_EPI_DS = 4      /* attention backward: acc = dP; C = dS = P*(dP - Drow)*alpha on j<=i (and i-j<window), else 0    */
};
enum {
  DB1_K_FULL = 0,        /* k in [0, K)                                                                                */
  DB1_K_END_BY_ROW = 1,  /* k in [0, min(K, (mt+1)*128))          (causal: contraction over keys j <= i)               */
  DB1_K_BEGIN_BY_ROW = 2,/* k in [mt*128, K)                      (causal: contraction over queries i >= j)            */
  DB1_K_BEGIN_REV = 3    /* k in [max(0, K - (mt+1)*128), K)      (relative-position order: c >= K-1-i)                */
};

typedef struct db1_gemm_desc {
  int32_t epilogue; /* DB1_EPI_* */
  int32_t M, N, K;
  int32_t a_mn;     /* 0: A stored [M][K] (K contiguous); 1: stored [K][M] (M contiguous) — no transposed copy needed */
  int32_t b_mn;     /* 0: B stored [N][K];                1: stored [K][N]                                              */
  const void* A;
  const void* B;
  void* C;
  int64_t lda, ldb, ldc; /* row strides of the STORED 2-D matrices, in elements (multiples of 8) */
  /* batching: z1 (inner, e.g. head) and z2 (outer, e.g. sequence in the batch); strides in elements, 0 = broadcast */
  int32_t Z1, Z2;
  int64_t a_z1, a_z2, b_z1, b_z2, c_z1, c_z2;
  int32_t reduce_z2; /* 1: the contraction also runs over z2 (C has no z2 dimension) */
  int32_t k_mode;    /* DB1_K_* */
  int32_t skip_upper;/* 1: output tiles entirely above the diagonal (first col > last row) are skipped */
  float alpha;
  int32_t accumulate; /* C += */
  const void* bias;   /* [N] fp16 or NULL */
  const void* resid;  /* [M, ldr] fp16 or NULL (un-batched calls only) */
  int64_t ldr;
  float drop_p;       /* dropout on (alpha*acc + bias) before the residual add; mask = f(seed, row*N + col) */
  uint64_t seed;
  const void* u;      /* QKV: r_w_bias flattened [d_model] */
  const void* v;      /* QKV: r_r_bias flattened [d_model] */
  int32_t d_model;
  void* H;            /* GEGLU: out [M, ldh]; DGEGLU: in */
  int64_t ldh;
  int32_t F;
  /* DS epilogue */
  const void* P;      /* fp16, same layout/strides as C */
  void* C2;           /* reserved (unused) */
  const float* Drow;  /* fp32 [Z2][Z1][M] contiguous */
  int32_t window;     /* attend iff 0 <= i-j < window */
  int32_t bn_hint;    /* 0 = auto, 128 or 256 */
  /* Row-dot side output (plain, un-batched GEMMs with N == dot_H * 128 at CTA-pair size; ABI version 2):
   * dot_out[(b * dot_H + h) * dot_L + i] = sum_{c in head h} acc[b * dot_L + i, c] * dot_with[b * dot_L + i, c].
   * Attention backward: C = dO = dZ . W_o and D = rowsum(dO * O) (softmax-backward row term) from the same epilogue. */
  const void* dot_with; /* fp16 [M, ld_dot] or NULL */
  int64_t ld_dot;
  float* dot_out;       /* fp32 [M / dot_L, dot_H, dot_L] */
  int32_t dot_L, dot_H;
  /* ABI version 3. 1: B holds parameters that no kernel enqueued just before this call writes (model weights during
   * inference): the few-row path (M <= 8) may then request them before the stream's previous kernel has completed
   * (programmatic dependent launch) - its weight stream overlaps that kernel's tail. 0 (default): B is read only after it. */
  int32_t b_static;
  /* ABI version 3, few-row path only (M <= 8, K-major A; the tensor-core path rejects it): the rows of A are LayerNorm-ed
   * over K before the product - A' = (A - mean) * rstd * ln_gamma + ln_beta, rounded to fp16 exactly as db1_layernorm_fwd
   * stores it (transformer_xl.py:238, :290 feeding :138 / :265 / :595) - so that the decode step needs no LayerNorm
   * launch between a block's last GEMM and the next block's first. ln_out (optional, [M, K] fp16, row stride K) receives
   * the normalised rows (the next residual / memory row). */
  const void* ln_gamma; /* [K] fp16 or NULL */
  const void* ln_beta;  /* [K] fp16 */
  float ln_eps;
  void* ln_out;
} db1_gemm_desc;

int db1_gemm_f16(const db1_gemm_desc* d, void* stream);

/* Optional scratch for the GEMM's stream-K tail (the library never allocates): when the tile count of a large GEMM is
 * not a multiple of the CTA pairs, the last partial wave is cut along K across all pairs; partial fp32 accumulators
 * and their arrival counters live here. `ws` = db1_gemm_workspace_bytes() bytes of device memory, ZERO-FILLED once by
 * the caller, 16-byte aligned, owned by the caller for as long as GEMMs are launched; GEMM launches that share it must
 * be stream-ordered (one compute stream). Without a workspace (or ws == NULL) every GEMM uses the plain data-parallel
 * tile schedule - same results up to fp32 summation order. Registration is per process and bound to the current device. */
int db1_gemm_set_workspace(void* ws, long long bytes);
long long db1_gemm_workspace_bytes(void);
/* Host-side view of that schedule (tests, tooling; no device work). db1_gemm_sk_choose: returns 1 and the number of
 * trailing pair-tiles / participating pairs if a GEMM with `tiles` 256 x 256 pair-tiles of KB 64-wide k-blocks would
 * use the stream-K tail on `pairs` CTA pairs, else 0. db1_gemm_sk_plan: work list of CTA pair `pair`:
 * out[14] = { nseg, T_dp, tile[2], kb0[2], kb1[2], role[2] (0 full, 1 contributor, 2 owner), peer[2], npeer[2] };
 * the pair then walks the data-parallel tiles pair, pair + pairs, ... < T_dp. */
int db1_gemm_sk_choose(long long tiles, int pairs, int KB, int* sk_tiles, int* sk_pairs);
int db1_gemm_sk_plan(int pair, long long tiles, int KB, int sk_tiles, int sk_pairs, int* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused relative-position causal attention, forward (tcgen05 QK^T / QR^T / PV on TMEM, online softmax, in-kernel
 * _rel_shift, causal + sliding-window predicate from indices).
 * Replaces transformer_xl.py:161-225 (AC/BD einsums, _rel_shift :98-110, scale, mask :177-204, softmax :209, PV :220)
 * for qlen == klen == L (training; mems == None).
 *   qu, qv, k, v : fp16 views [B, L, H, dh] with row stride ld_qkv (columns of the fused QKV buffer); qu = q + r_w_bias,
 *                  qv = q + r_r_bias (written by db1_gemm_f16's QKV epilogue)
 *   r            : fp16 [L, H*dh] = r_net(pos_emb) in the reference's row order (row c <-> distance L-1-c)
 *   mode 0       : out [B*L, ld_out] fp16 (head h at column h*dh), lse2 [B,H,L] fp32 (log2-domain log-sum-exp) written
 *   mode 1       : lse2 read; probs [B,H,L,L] fp16 written on the visited (causal) tiles — backward recompute
 *   window       : attend iff 0 <= i-j < window   (same_length/mem_len semantics, transformer_xl.py:551-562)
 * ------------------------------------------------------------------------------------------------------------------ */
int db1_relattn_fwd(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv, const void* r,
                    long long ld_r, void* out, long long ld_out, float* lse2, void* probs, int B, int L, int H, int dh,
                    int window, float scale, int mode, void* stream);

/* Memory-augmented inference (transformer_xl.py:124-133 with mems, evaluate_rl.py:157-266): the same kernel over the
 * K = mlen + qlen rows of cat(mem, w); only query rows >= mlen are computed and written (out / lse2 are addressed
 * with the row index in the concatenated sequence). r holds K rows. window as above (keys with 0 <= delta < window). */
int db1_relattn_mem_fwd(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv, const void* r,
                        long long ld_r, void* out, long long ld_out, float* lse2, int B, int K, int H, int dh,
                        int window, float scale, int mlen, void* stream);

/* Attention backward, first half (same tiling as the forward): recomputes P = softmax(S) from the saved LSE, forms
 * dP = dO . V^T on the tensor cores and writes both P and dS = P * (dP - D) * scale ([B,H,L,L] fp16, visited causal
 * tiles only; D = rowsum(dO * O) from db1_rowdot). Adjoint of transformer_xl.py:173-225; the contractions that consume
 * P / dS (dV, dK, dQ, dR) are db1_gemm_f16 calls. */
int db1_relattn_bwd_ds(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv, const void* r,
                       long long ld_r, const void* dout, long long ld_do, const float* lse2, const float* drow,
                       void* probs, void* ds, int B, int L, int H, int dh, int window, float scale, void* stream);
/* Same, with D = rowsum(dO * O) formed inside the kernel from the forward output `o` ([B*L, ld_o] fp16, head h at column
 * h*dh) instead of a db1_rowdot result. */
int db1_relattn_bwd_ds_o(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv, const void* r,
                         long long ld_r, const void* dout, long long ld_do, const void* o, long long ld_o,
                         const float* lse2, void* probs, void* ds, int B, int L, int H, int dh, int window, float scale,
                         void* stream);

/* As db1_relattn_bwd_ds (drow given) / db1_relattn_bwd_ds_o (o given, drow NULL), but P and dS are written in the TILED
 * layout [B][H][nq][nq][2][128][64] fp16, nq = ceil(L/128): tile (I, J) of a head = two contiguous slabs of 128 rows x 64
 * columns. Visited tiles are written completely (masked entries, rows / columns beyond L: zeros); any L. This is the
 * producer for db1_relattn_bwd_dkdv / _dq / _dr, whose TMA boxes are then single contiguous 16 KB reads. */
int db1_relattn_bwd_ds_tiled(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv, const void* r,
                             long long ld_r, const void* dout, long long ld_do, const void* o, long long ld_o,
                             const float* lse2, const float* drow, void* probs, void* ds, int B, int L, int H, int dh,
                             int window, float scale, void* stream);

/* Attention backward, second half (csrc/relattn_bwd.cu): everything that consumes dS, on tcgen05 with the accumulators
 * resident in TMEM; the adjoint of _rel_shift (transformer_xl.py:98-110) is done on registers per tile row, so no
 * re-laid-out copy of dS (dSr) exists and P / dS are not re-read by generic GEMMs.
 *   ds : fp16, TILED layout [B,H,nq,nq,2,128,64] from db1_relattn_bwd_ds_tiled (nq = ceil(L/128); visited causal /
 *        in-window tiles complete, masked entries exactly 0; tile (I,J) = two contiguous 16 KB slabs = one TMA box each)
 * db1_relattn_bwd_dq (query-outer): dq[b*L+i, h*dh..] = sum_j ds[i,j] k_j + sum_j ds[i,j] r[j+L-1-i]   (fp16, written)
 *   du[h*dh..] += sum_{b,i} (ds . K)_i ;  dv[h*dh..] += sum_{b,i} (unshift(ds) . R)_i     (fp32, accumulated: gradients of
 *   r_w_bias / r_r_bias, transformer_xl.py:161, :167).   k: fp16 view [B,L,H,dh] with row stride ld_qkv; r: [L, H*dh].
 * db1_relattn_bwd_dr (diagonal-outer): dr[c, h*dh..] += sum_{b,(i,j): j+L-1-i=c} ds[i,j] * qv[b,i,h,:]   (fp32 [L, ld_dr],
 *   accumulated: the caller zero-fills it; gradient of r_net's output, adjoint of transformer_xl.py:167-171). */
/* db1_relattn_bwd_dkdv (key-outer): dv[b*L+j, h*dh..] = sum_i probs[i,j] dout[b*L+i, h*dh..]  (adjoint of transformer_xl.py:220),
 *   dk[b*L+j, h*dh..] = sum_i ds[i,j] qu[b,i,h,:]  (adjoint of :161-165); fp16, written; dk / dv share the row stride ld_dkv.
 *   probs : fp16, tiled like ds (db1_relattn_bwd_ds_tiled). */
int db1_relattn_bwd_dkdv(const void* probs, const void* ds, const void* dout, long long ld_do, const void* qu,
                         long long ld_qkv, void* dk, void* dv, long long ld_dkv, int B, int L, int H, int dh, int window,
                         void* stream);
int db1_relattn_bwd_dq(const void* ds, const void* k, long long ld_qkv, const void* r, long long ld_r, void* dq,
                       long long ld_dq, float* du, float* dv, int B, int L, int H, int dh, int window, void* stream);
int db1_relattn_bwd_dr(const void* ds, const void* qv, long long ld_qkv, float* dr, long long ld_dr, int B, int L, int H,
                       int dh, int window, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Memory-augmented decode step with cached keys / values (csrc/decode.cu). The reference recomputes qkv_net over
 * cat(mem, w) for every call of forward(..., mems=...) (transformer_xl.py:124-141); here the k / v rows of the memory
 * are kept per layer in ring buffers [B, cap, H*dh] (cap = mem_len, always full: init_mem starts from zeros, :470-485)
 * and only the new rows go through the projection GEMMs. Logical memory row j lives in slot (head + j) % cap.
 * ------------------------------------------------------------------------------------------------------------------ */
/* number of key splits db1_relattn_decode uses for this problem (sizes its workspace) */
int db1_decode_splits(int B, int Q, int H);
/* out[b*Q+i, h*dh..] = softmax_j( scale * ((q_i+u).k_j + (q_i+v).r[Q-1-i+j]) ) v_j over the keys query i may see
 * (j <= i + cap and cap + i - j < window; j < cap: cache rows in logical order, j >= cap: the new rows), fp32 softmax.
 * qu / qv / knew / vnew: columns of the NEW rows' fused qkv buffer [B*Q, ld_qkv]; r: [cap + Q, ld_r] = r_net(pos_emb) in the
 * reference's row order; ws: fp32 workspace of B*Q*H * db1_decode_splits(B,Q,H) * (dh + 2) floats. */
int db1_relattn_decode(const void* qu, const void* qv, const void* knew, const void* vnew, long long ld_qkv,
                       const void* kcache, const void* vcache, int cap, int head, const int* head_dev, const void* r,
                       long long ld_r, void* out, long long ld_out, float* ws, long long ws_floats, int B, int Q, int H,
                       int dh, int window, float scale, void* stream);
/* rings[z][b][(head + t) % cap][0..n) = srcs[z][b*Q + t][0..n) for nring <= 3 (source, ring) pairs in ONE launch (layer
 * input rows, new k rows, new v rows): appends the Q new rows of every sequence over the oldest slots - the in-place form
 * of _update_mem's cat(mem, h)[:, -mem_len:] (transformer_xl.py:487-504); the caller advances the head by Q afterwards.
 * srcs / ld_srcs / rings are HOST arrays of device pointers / row strides.
 * head_dev (both calls): if not NULL the head is read from this device int instead of `head`, so that a decode step
 * captured in a CUDA graph replays correctly while the head moves. */
int db1_ring_append(const void* const* srcs, const long long* ld_srcs, void* const* rings, int nring, int cap, int head,
                    const int* head_dev, int B, int Q, int n, void* stream);
/* out[row] = argmax over columns [lo, hi) of logits[row] (- add_mask[c - lo] if given); first maximum wins.
 * = masked_logits_for_action (evaluate_rl.py:96-124: everything outside the action-token range gets -1e10, optional
 * environment action mask) followed by argmax (:196-199). */
int db1_masked_argmax(const void* logits, long long ld, int rows, int lo, int hi, const float* add_mask, long long* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * HBM-bound kernels (single pass over the large operand, 16-byte vector accesses, fp32 math).
 * ------------------------------------------------------------------------------------------------------------------ */

/* out = LayerNorm(y) * gamma + beta; stats[row] = (mean, rstd) fp32 kept for the backward.
 * Replaces nn.LayerNorm at transformer_xl.py:238 and :290 (the residual sum y is produced by the GEMM epilogue). */
int db1_layernorm_fwd(const void* y, const void* gamma, const void* beta, void* out, float* stats, int rows, int d,
                      float eps, void* stream);
/* dy = dLN(dout); dz = dy * dropout_mask(seed)/(1-p) (written only when drop_p > 0; the mask is the one the producing
 * GEMM epilogue applied); fp32 accumulators: dgamma += dout*xhat, dbeta += dout, dbias += dz (dbias may be NULL). */
int db1_layernorm_bwd(const void* dout, const void* y, const void* gamma, const float* stats, void* dy, void* dz,
                      float* dgamma, float* dbeta, float* dbias, int rows, int d, float drop_p, uint64_t seed,
                      void* stream);

/* Masked cross-entropy (transformer_xl.py:602-609): row_loss = (lse - z[label]) * mask, loss2 = {sum(row_loss)/sum(mask),
 * sum(mask)} (device scalars; no host sync). logits fp16 [rows, ld], V valid columns. */
int db1_ce_fwd(const void* logits, long long ld, const long long* labels, const float* mask, float* row_loss,
               float* row_lse, float* loss2, int rows, int V, void* stream);
/* dlogits = (softmax - onehot) * mask / sum(mask) * gscale[0]  (gscale: device fp32 scalar, the upstream gradient) */
int db1_ce_bwd(const void* logits, long long ld, const long long* labels, const float* mask, const float* row_lse,
               const float* loss2, const float* gscale, void* dlogits, long long ldd, int rows, int V, void* stream);

/* Embedding assembly (transformer_xl.py:627-649, :665, :677-695): out[b,l,:] = dropout((tok>=0 ? W[tok] : vis[b,slot])
 * + (pos ? T[pos] : 0)); slot[b,l] = number of -1 tokens before l (written here, reused by the backward).
 * out/vis are addressed as base + b*batch_stride + index*d so segments of a longer sequence can be written in place. */
int db1_embed_fwd(const long long* tok, const long long* pos, int* slot, const void* W, const void* T, const void* vis,
                  long long vis_bs, int nvis, void* out, long long out_bs, int B, int L, int d, int V, float drop_p,
                  uint64_t seed, long long seed_row0, void* stream);
/* Adjoint: dW[tok] += g, dT[pos] += g (fp16 atomics), dvis[b,slot] = g with g = dout * dropout mask. */
int db1_embed_bwd(const long long* tok, const long long* pos, const int* slot, const void* dout, long long dout_bs,
                  void* dW, void* dT, void* dvis, long long vis_bs, int nvis, int B, int L, int d, int V, float drop_p,
                  uint64_t seed, long long seed_row0, void* stream);

/* out[c] += sum_r in[r,c] (fp16 -> fp32): bias gradients. */
int db1_colsum(const void* in, long long ld, float* out, int rows, int n, void* stream);
/* dq = dqu + dqv; du += colsum(dqu); dv += colsum(dqv): gradients of q, r_w_bias, r_r_bias (transformer_xl.py:161,167) */
int db1_dq_finalize(const void* dqu, const void* dqv, long long ld_in, void* dq, long long ld_out, float* du, float* dv,
                    int rows, int n, void* stream);
/* out[b,h,i] = sum_d a[b,i,h,d] * b[b,i,h,d]  (softmax-backward row term) */
int db1_rowdot(const void* a, const void* b, long long ld, float* out, int B, int L, int H, int dh, void* stream);
/* Sinusoid rows of PositionalEmbedding (transformer_xl.py:43-50, 569-575): row c = [sin|cos](min(klen-1-c, clamp)*inv_freq) */
int db1_posemb(void* out, const float* inv_freq, int klen, int d, int clamp_len, float drop_p, uint64_t seed,
               void* stream);
/* Same rows with the phase arithmetic of the reference under module.half() (what DeepSpeed fp16 training and the released
 * checkpoint use): position, inv_freq and their product each rounded to fp16 before sin / cos (:44 with the fp16-cast
 * buffer, :569-571 with fp16 hidden states). inv_freq is still passed as fp32 (the constructor's values). */
int db1_posemb_half_phase(void* out, const float* inv_freq, int klen, int d, int clamp_len, float drop_p, uint64_t seed,
                          void* stream);
/* dsr[z][i][c] = ds[z][i][c-(L-1-i)] for c >= L-1-i, else 0: adjoint of _rel_shift (transformer_xl.py:98-110) */
int db1_rel_unshift(const void* ds, void* dsr, int Z, int L, void* stream);
int db1_f32_to_f16(const float* src, void* dst, long long n, int accumulate, void* stream);
/* nseg <= 8 conversions in one launch: dst[k][0..n[k]) (+)= fp16(src[off[k] .. off[k]+n[k])). dst / off / n / accumulate
 * are HOST arrays. Used to drop a block's small fp32 gradient accumulators straight into the gradient buckets. */
int db1_f32_to_f16_multi(const float* src, int nseg, void* const* dst, const long long* off, const long long* n,
                         const int* accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Image patch embedder (src/tokenizer/vision_embedding.py:36-86). Patches are the batch dimension; activations are
 * [P, 256 pixels, 64 channels] fp16 (channels-last). The two conv3x3(64->64) layers and the 16x16/stride-16 projection
 * run as db1_gemm_f16 on the matrices these kernels produce.
 * ------------------------------------------------------------------------------------------------------------------ */
/* out[b][c][r] = in[b][r][c]: weight re-layouts (conv [co][ci][9] <-> [co][9][ci]; projection [d][64][256] <-> [d][256][64]) */
int db1_transpose_f16(const void* in, void* out, int batch, int rows, int cols, void* stream);
/* einops patch split + per-(patch, channel) standardisation (x-mean)/(1e-6+std_unbiased)/4 (:67-79) + conv1 3x3 (C->64)
 * (:80). pixels [N,C,H,W] fp16; xs [P,C,256] (standardised pixels, kept for the weight gradient); y1 [P,256,64]. */
int db1_patch_conv1_fwd(const void* pixels, const void* W1, const void* b1, void* xs, void* y1, int N, int C, int Himg,
                        int Wimg, void* stream);
/* Same with fp32 pixels: mean / std are formed from the unrounded values, as the reference does (it standardises in the
 * input dtype, :73-77, and casts to the module dtype afterwards, :79). */
int db1_patch_conv1_fwd_f32(const float* pixels, const void* W1, const void* b1, void* xs, void* y1, int N, int C,
                            int Himg, int Wimg, void* stream);
/* dW1 [64, C*9] fp32 += sum over patches and pixels of dy1 (x) shifted xs (weight gradient of conv1). */
int db1_patch_conv1_bwd(const void* xs, const void* dy1, float* dW1, int P, int C, void* stream);
/* GroupNorm(32 groups of 2 channels, eps) -> exact GELU -> im2col: col [P*256, 576], col[q][tap*64+ci] = a[q+tap][ci]
 * with per-patch zero padding (:57-62 residual_path.{0,1} / {3,4} and the unfold of the following conv3x3).
 * stats [P,32,2] = (mean, rstd) per group for the backward. */
int db1_gn_gelu_im2col(const void* x, const void* gamma, const void* beta, void* col, float* stats, int P, float eps,
                       void* stream);
/* Adjoint of the above: col2im of dcol [P*256,576], GELU', GroupNorm backward; dx [P,256,64] (+ dres if not NULL),
 * dgamma / dbeta [64] fp32 accumulated. */
int db1_col2im_gn_gelu_bwd(const void* dcol, const void* x, const float* stats, const void* gamma, const void* beta,
                           const void* dres, void* dx, float* dgamma, float* dbeta, int P, void* stream);
/* out = x * keep(seed, element index) / (1 - p): embedding dropout on image-patch rows (transformer_xl.py:545);
 * applying it to a gradient with the same seed is its adjoint. n % 8 == 0. */
int db1_dropout_f16(const void* x, void* out, long long n, float drop_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused optimizer step over flat buckets (what DeepSpeed's fp16 optimizer + FusedAdam do in the reference's loop,
 * src/train_utils/train.py:231-232; flags train_config.py:212-235). HBM-bound: 28 bytes per parameter.
 * ------------------------------------------------------------------------------------------------------------------ */
/* out[0] += sum(g^2) over an fp16 bucket (n % 8 == 0); inf / nan propagate (overflow detection). */
int db1_grad_sumsq(const void* g16, long long n, float* out, void* stream);
/* gcoef2[0] = inv_scale * min(1, clip / (global norm + 1e-6)) (0 on overflow), gcoef2[1] = unscaled global gradient norm,
 * overflow_flag[0] = 1 iff sumsq is inf/nan. All device scalars: no host sync. clip <= 0 disables clipping. */
int db1_clip_coef(const float* sumsq, float inv_scale, float clip, float* gcoef2, int* overflow_flag, void* stream);
/* Adam / AdamW on fp32 master weights from an fp16 gradient bucket scaled by gcoef[0] (device scalar); refreshes the
 * fp16 parameters. step >= 1 (bias correction). */
int db1_adam_step(const void* g16, void* p16, float* master, float* m, float* v, long long n, const float* gcoef,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, int adamw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DB1_SM100_H_ */
