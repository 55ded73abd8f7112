/*
 * gru_b200.h -- C ABI of the batched GRU text generator (ggml-experiments_b200/host/gru.cpp).
 *
 * The reference (/root/reference/rnn_text_gen/rnn_text_generation.cpp) generates ONE character stream: load_model
 * (:97-164), one GRU-cell graph (gru_forward :186-263), a 200-step greedy loop on the host (inference :266-314).
 * BASELINE.json config 5 batches the cell over B independent streams.  This program builds the same cell from ggml_* calls
 * for [.., B] tensors, keeps the loop on the device (argmax + state feedback) and runs through
 * ggml_graph_compute_with_ctx like the reference.
 *
 *   gru_load      <- load_model                (rnn.cpp:97-164; same gru.bin layout, kernels pre-transposed on the host)
 *   gru_generate  <- inference + gru_forward   (rnn.cpp:186-314), B streams, greedy
 */
#ifndef GRU_B200_H
#define GRU_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gru_model gru_model;

gru_model * gru_load(const char * gru_bin_path);   /* NULL if the file cannot be read */
void        gru_free(gru_model * m);
int         gru_vocab(const gru_model * m);        /* 66 */
int         gru_units(const gru_model * m);        /* 1024 */

/* Feed `first_tokens[b]` to stream b (zero initial state, rnn.cpp:283 + App. C #12), then feed back the greedy argmax for
 * `steps` steps in total.  out_tokens[t*B + b] = token chosen by stream b after step t (so row t-1 is the input of step t).
 * final_state (may be NULL): [B][units] f32.  Returns the device time of the loop in milliseconds (CUDA events), < 0 on error. */
float gru_generate(gru_model * m, const int32_t * first_tokens, int B, int steps, int32_t * out_tokens, float * final_state);

#ifdef __cplusplus
}
#endif
#endif
