/* <ggml.h> as included by /root/reference/rnn_text_gen/rnn_text_generation.cpp:1 */
#include "ggml/ggml.h"
