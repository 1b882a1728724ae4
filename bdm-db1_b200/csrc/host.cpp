// Host-side integer paths of the DB1 batch contract (libdb1_host.so, plain C++ — no CUDA).
//   db1_discretize / db1_decode : src/tokenizer/scalar_tokenizer.py:28-63
//   db1_rl_layout               : src/data/rl_dataset.py:44-71 (action flag / local position ids),
//                                 :683-697 (join), :711-716 + :865-872 (pad/truncate to L+1), :738-746 (shift)
//   db1_build_rl_sample_idx     : src/data/helpers.cpp:82-115 semantics (trajectory window index)
//   db1_build_sample_idx        : src/data/helpers.cpp:117-203 semantics (GPT packed-sample index)
// Compiled with -ffp-contract=off: every float32 operation below is rounded exactly once, in the order the
// reference's torch expression evaluates it.
#include "../../include/db1_host.h"

#include <math.h>
#include <stdint.h>
#include <string.h>

extern "C" {

int db1_discretize(const float* x, int32_t* out, long long n, int is_action, int num_bins, float mu, float M) {
  if (!x || !out || n < 0 || num_bins <= 0) return -1;
  const float denom = logf(mu * M + 1.0f);
  for (long long i = 0; i < n; ++i) {
    float v = x[i];
    if (!is_action) {
      const float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
      float t = fabsf(v) * mu;
      t = t + 1.0f;
      t = logf(t);
      t = sgn * t;
      t = t / denom;
      v = t < -1.f ? -1.f : (t > 1.f ? 1.f : t);
    }
    float s = v + 1.0f;
    s = s / 2.0f;
    s = s * (float)num_bins;
    // torch .int() truncates toward zero; NaN/inf behaviour is undefined there and not part of the contract
    long long b = (long long)s;
    if (b < 0) b = 0;
    if (b > num_bins - 1) b = num_bins - 1;
    out[i] = (int32_t)b;
  }
  return 0;
}

int db1_decode(const int32_t* tok, float* out, long long n, int is_action, int num_bins, float mu, float M) {
  if (!tok || !out || n < 0 || num_bins <= 0) return -1;
  for (long long i = 0; i < n; ++i) {
    int32_t t = tok[i];
    if (t < 0) t = 0;
    if (t > num_bins - 1) t = num_bins - 1;
    float v = (float)t / (float)num_bins;
    v = v * 2.0f;
    v = v - 1.0f;
    if (!is_action) {
      const float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
      float e = powf(1.0f + M * mu, fabsf(v));
      e = e - 1.0f;
      e = sgn * e;
      v = e / mu;
    }
    out[i] = v;
  }
  return 0;
}

// One RL sample. obs [T, obs_len], act [T, act_len] are vocabulary ids (-1 = image patch slot).
// Outputs (each seq_len long): tensor_seq, label (int64), loss_mask (float), position_id (int64).
int db1_rl_layout(const long long* obs, const long long* act, int T, int obs_len, int act_len, long long sep_id,
                  int seq_len, long long pad_id, int prepend_trans_num, long long* tensor_seq, long long* label,
                  float* loss_mask, long long* position_id) {
  if (!obs || !act || !tensor_seq || !label || !loss_mask || !position_id) return -1;
  if (T <= 0 || obs_len < 0 || act_len < 0 || seq_len <= 0) return -1;
  const int step = obs_len + act_len + 1;
  const long long total = (long long)T * step;
  const int n = seq_len + 1;
  for (int p = 0; p < n; ++p) {
    long long tokv = pad_id;
    long long flag = 0, pos = 0;
    if (p < total) {
      const int t = p / step, o = p % step;
      if (o < obs_len) tokv = obs[(long long)t * obs_len + o];
      else if (o == obs_len) tokv = sep_id;
      else tokv = act[(long long)t * act_len + (o - obs_len - 1)];
      pos = (o <= obs_len) ? (o + 1) : 0;
      flag = (o > obs_len && t >= prepend_trans_num) ? 1 : 0;
    }
    if (p < seq_len) {
      tensor_seq[p] = tokv;
      position_id[p] = pos;
    }
    if (p >= 1) {
      label[p - 1] = tokv;
      loss_mask[p - 1] = (float)flag;
    }
  }
  return 0;
}

// rows (i, j, min(j + transition_num, len_i)) for every trajectory i and start j in [0, len_i - 1)
long long db1_build_rl_sample_idx(const int32_t* path_lengths, long long n_paths, int transition_num, int32_t* out,
                                  long long out_rows) {
  if (!path_lengths || n_paths < 0) return -1;
  long long r = 0;
  for (long long i = 0; i < n_paths; ++i) {
    const int len = path_lengths[i];
    for (int j = 0; j + 1 < len; ++j) {
      if (out) {
        if (r >= out_rows) return -2;
        out[3 * r + 0] = (int32_t)i;
        out[3 * r + 1] = j;
        const int e = j + transition_num;
        out[3 * r + 2] = e < len ? e : len;
      }
      ++r;
    }
  }
  return r;
}

// GPT sample index (src/data/helpers.cpp:117-203 semantics, Megatron packing): documents doc_idx[0..n_doc_idx) are laid
// end to end; sample k covers seq_length + 1 tokens starting at flattened token k * seq_length (one token of overlap),
// so its start is the (document slot, offset) that holds that token. out: [num_samples + 1, 2] int32 with
// num_samples = (num_epochs * tokens_per_epoch - 1) / seq_length. out == NULL: returns the row count only.
// Returns the number of rows, -1 on bad arguments, -2 if out_rows is too small, -3 if the documents run out.
long long db1_build_sample_idx(const int32_t* sizes, const int32_t* doc_idx, long long n_doc_idx, int seq_length,
                               int num_epochs, long long tokens_per_epoch, int32_t* out, long long out_rows) {
  if (!sizes || !doc_idx || n_doc_idx <= 0 || seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) return -1;
  const long long num_samples = ((long long)num_epochs * tokens_per_epoch - 1) / seq_length;
  const long long rows = num_samples + 1;
  if (!out) return rows;
  if (out_rows < rows) return -2;
  long long slot = 0;       // index into doc_idx of the document holding the current start token
  long long slot_begin = 0; // flattened index of that document's first token
  for (long long k = 0; k < rows; ++k) {
    const long long start = k * (long long)seq_length;
    while (start >= slot_begin + sizes[doc_idx[slot]]) {
      slot_begin += sizes[doc_idx[slot]];
      if (++slot >= n_doc_idx) return -3;
    }
    out[2 * k] = (int32_t)slot;
    out[2 * k + 1] = (int32_t)(start - slot_begin);
  }
  return rows;
}

}  // extern "C"
