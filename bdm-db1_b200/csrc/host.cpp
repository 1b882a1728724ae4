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
#include <vector>
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


// One RL sample from raw arrays: RLFullDataset.get (src/data/rl_dataset.py:614-752) with postprocess_obs_and_act
// (:393-473) for the observation kinds {image frames, float vector} and float / discrete actions.
//   per transition: [n_img_slots x -1 | n_float continuous-bin tokens | SEP | act_len action tokens]     (:651-697)
//   float -> token: discretize + text_vocab (+ n_disc unless overlap_with_text)                           (:429-435, :459-464)
//   discrete action -> token: value (+ text_vocab unless overlap_with_text)                               (:465-471)
//   SEP = n_cont + text_vocab (+ n_disc unless overlap_with_text)                                          (:683-685)
//   pad / truncate to seq_len + 1 (:711-716), then the observation slots of the padded frames T..n_frames-1 are set to -1
//   (:718-726: the reference overwrites the WHOLE observation span of those transitions), then the shift (:738-746).
int db1_rl_assemble(const float* obs_float, int n_float, int n_img_slots, const float* act_float,
                    const long long* act_disc, int act_len, int T, int n_frames, int text_vocab, int n_disc, int n_cont,
                    int overlap_with_text, int seq_len, int prepend_trans_num, long long* tensor_seq, long long* label,
                    float* loss_mask, long long* position_id) {
  if (!tensor_seq || !label || !loss_mask || !position_id) return -1;
  if (T <= 0 || n_float < 0 || n_img_slots < 0 || act_len <= 0 || seq_len <= 0) return -1;
  if ((n_float > 0 && !obs_float) || ((act_float == nullptr) == (act_disc == nullptr))) return -1;
  const int obs_len = n_img_slots + n_float;
  const long long cont0 = (long long)text_vocab + (overlap_with_text ? 0 : n_disc);
  const long long disc0 = overlap_with_text ? 0 : text_vocab;
  const long long sep = (long long)n_cont + cont0;
  std::vector<long long> obs((size_t)T * (obs_len > 0 ? obs_len : 1)), act((size_t)T * act_len);
  std::vector<int32_t> tmp((size_t)T * (n_float > act_len ? n_float : act_len));
  if (n_float > 0 && db1_discretize(obs_float, tmp.data(), (long long)T * n_float, 0, n_cont, 100.0f, 256.0f)) return -1;
  for (int t = 0; t < T; ++t) {
    for (int o = 0; o < n_img_slots; ++o) obs[(size_t)t * obs_len + o] = -1;
    for (int o = 0; o < n_float; ++o) obs[(size_t)t * obs_len + n_img_slots + o] = tmp[(size_t)t * n_float + o] + cont0;
  }
  if (act_float) {
    if (db1_discretize(act_float, tmp.data(), (long long)T * act_len, 1, n_cont, 100.0f, 256.0f)) return -1;
    for (size_t i = 0; i < (size_t)T * act_len; ++i) act[i] = tmp[i] + cont0;
  } else {
    for (size_t i = 0; i < (size_t)T * act_len; ++i) {
      if (act_disc[i] < 0 || act_disc[i] >= n_disc) return -3;  // the reference asserts the range (:466)
      act[i] = act_disc[i] + disc0;
    }
  }
  // layout into a (seq_len + 1)-token window first: the -1 fill of padded frames applies to the un-shifted sequence
  const int n = seq_len + 1;
  std::vector<long long> ts(n + 1), lb(n + 1), ps(n + 1);
  std::vector<float> lm(n + 1);
  // db1_rl_layout on n tokens gives tensor_seq = joined[0..n), label = joined[1..n+1): run it with seq_len = n and keep
  // tensor_seq / position_id (the un-shifted window) and the flags through its label-aligned loss_mask
  if (db1_rl_layout(obs.data(), act.data(), T, obs_len, act_len, sep, n, 0, prepend_trans_num, ts.data(), lb.data(), lm.data(),
                    ps.data()))
    return -1;
  const int step = obs_len + act_len + 1;
  for (int i = T; i < n_frames; ++i) {
    const long long b = (long long)i * step;
    long long e = b + obs_len;
    if (e > n) e = n;
    for (long long q = b; q < e; ++q) ts[q] = -1;
  }
  for (int q = 0; q < seq_len; ++q) {
    tensor_seq[q] = ts[q];
    position_id[q] = ps[q];
    label[q] = ts[q + 1];
    loss_mask[q] = lm[q];  // flag of token q+1 (db1_rl_layout already shifts the flags by one)
  }
  return 0;
}

// my_collate_fn (src/data/data_samplers.py:28-42), the grouping: samples are grouped by task type in order of first
// appearance (the reference's defaultdict insertion order), inside a group in list order. perm[n]: sample indices
// group by group; group_type / group_count [<= n]: one entry per group. Returns the number of groups.
int db1_collate_plan(const int* type_ids, int n, int* perm, int* group_type, int* group_count) {
  if (!type_ids || !perm || !group_type || !group_count || n <= 0) return -1;
  int ng = 0;
  for (int i = 0; i < n; ++i) {
    int g = 0;
    while (g < ng && group_type[g] != type_ids[i]) ++g;
    if (g == ng) {
      group_type[ng] = type_ids[i];
      group_count[ng] = 0;
      ++ng;
    }
    ++group_count[g];
  }
  int w = 0;
  for (int g = 0; g < ng; ++g)
    for (int i = 0; i < n; ++i)
      if (type_ids[i] == group_type[g]) perm[w++] = i;
  return ng;
}

// ... and the concatenation on dim 0 (`x.apply(torch.cat, dim=0)`): n contiguous blocks copied back to back into dst
// (typically a pinned staging buffer, so the batch is ready for one asynchronous host-to-device copy).
int db1_concat_rows(void* dst, const void* const* srcs, const long long* nbytes, int n) {
  if (!dst || !srcs || !nbytes || n <= 0) return -1;
  char* w = static_cast<char*>(dst);
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] < 0 || (nbytes[i] > 0 && !srcs[i])) return -1;
    memcpy(w, srcs[i], (size_t)nbytes[i]);
    w += nbytes[i];
  }
  return 0;
}

}  // extern "C"
