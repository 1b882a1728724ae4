/*
 * db1_host.h — C ABI of libdb1_host.so: the host-side INTEGER paths of DB1's batch contract (plain C++, no CUDA).
 * All pointers are HOST pointers. Results are bit-exact with the reference (tests/test_oracle_golden.py).
 * Return 0 = ok, negative = argument error (db1_build_rl_sample_idx returns the row count).
 */
#ifndef DB1_HOST_H_
#define DB1_HOST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ContinuousScalarTokenizer.discretize (src/tokenizer/scalar_tokenizer.py:28-45): mu-law (observations only), clamp to
 * [-1,1], uniform bins; float32 operations rounded once each in the reference's order (-ffp-contract=off). */
int db1_discretize(const float* x, int32_t* out, long long n, int is_action, int num_bins, float mu, float M);
/* ContinuousScalarTokenizer.decode (src/tokenizer/scalar_tokenizer.py:47-63). */
int db1_decode(const int32_t* tok, float* out, long long n, int is_action, int num_bins, float mu, float M);
/* One RL sample -> (tensor_seq, label, loss_mask, position_id), each seq_len long: action flags / local position ids
 * (src/data/rl_dataset.py:44-71), join :683-697, pad/truncate to L+1 :711-716 + :865-872, shift :738-746.
 * obs [T, obs_len] and act [T, act_len] hold vocabulary ids (-1 = image patch slot). */
int db1_rl_layout(const long long* obs, const long long* act, int T, int obs_len, int act_len, long long sep_id,
                  int seq_len, long long pad_id, int prepend_trans_num, long long* tensor_seq, long long* label,
                  float* loss_mask, long long* position_id);
/* build_rl_sample_idx (src/data/helpers.cpp:82-115): rows (i, j, min(j + transition_num, len_i)) for every trajectory
 * i and start j in [0, len_i - 1). out == NULL counts only. Returns the number of rows, -2 if out_rows is too small. */
long long db1_build_rl_sample_idx(const int32_t* path_lengths, long long n_paths, int transition_num, int32_t* out,
                                  long long out_rows);

/* build_sample_idx (src/data/helpers.cpp:117-203, used by gpt_dataset.py:283-289): packed GPT samples of
 * seq_length + 1 tokens over the documents doc_idx[0..n_doc_idx) laid end to end; row k = (index into doc_idx, offset in
 * that document) of flattened token k * seq_length. Rows = (num_epochs * tokens_per_epoch - 1) / seq_length + 1.
 * out == NULL counts only. Returns rows, -2 if out_rows is too small, -3 if the documents run out. */
long long db1_build_sample_idx(const int32_t* sizes, const int32_t* doc_idx, long long n_doc_idx, int seq_length,
                               int num_epochs, long long tokens_per_epoch, int32_t* out, long long out_rows);

/* One RL sample from RAW arrays - RLFullDataset.get (src/data/rl_dataset.py:614-752) with postprocess_obs_and_act
 * (:393-473): per transition [n_img_slots x -1 | n_float mu-law tokens | SEP | act_len action tokens]; exactly one of
 * act_float (continuous, no mu-law) / act_disc (values in [0, n_disc)) is non-NULL; n_frames >= T is the length of the
 * zero-padded frame sequence (the observation slots of transitions T..n_frames-1 become -1, :718-726), 0 without images.
 * Outputs as db1_rl_layout. Returns -3 if a discrete action is out of range (the reference asserts). */
int db1_rl_assemble(const float* obs_float, int n_float, int n_img_slots, const float* act_float,
                    const long long* act_disc, int act_len, int T, int n_frames, int text_vocab, int n_disc, int n_cont,
                    int overlap_with_text, int seq_len, int prepend_trans_num, long long* tensor_seq, long long* label,
                    float* loss_mask, long long* position_id);
/* my_collate_fn (src/data/data_samplers.py:28-42): samples grouped by task type in order of first appearance; perm[n] =
 * sample indices group by group, group_type / group_count one entry per group. Returns the number of groups. */
int db1_collate_plan(const int* type_ids, int n, int* perm, int* group_type, int* group_count);
/* The concatenation on dim 0 of one field of one group: n contiguous blocks copied back to back into dst. */
int db1_concat_rows(void* dst, const void* const* srcs, const long long* nbytes, int n);

#ifdef __cplusplus
}
#endif
#endif /* DB1_HOST_H_ */
