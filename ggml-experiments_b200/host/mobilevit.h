// mobilevit.h -- host-side MobileViT program: model structs, loader and graph builder.
//
// Mirrors the layer structure of /root/reference/mobilevit/main.cpp (structs at :56-213) so that the
// parity tests read like the reference, but:
//   * every forward() takes a batch (N images) and any H x W that is a multiple of 32
//     (the reference hard-codes (256,256,3,1), main.cpp:612,737,742,757-764,976);
//   * channel counts / stage counts come from the tensors in the file, not from compile-time hparams
//     (main.cpp:35-53,335-336), so S / XS / XXS files all load;
//   * graphs are cached per input shape and replayed.
// All arithmetic is expressed as ggml_* calls (include/ggml/ggml.h) and runs on the GPU.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "ggml/ggml.h"

namespace mvit {

struct hparams {  // main.cpp:35-53 (the fields that influence the forward pass)
    int   patch_size          = 2;
    int   num_attention_heads = 4;
    float layer_norm_eps      = 1.0e-5f;
};

struct conv_layer {  // main.cpp:56-73
    ggml_tensor * kernel = nullptr, * gamma = nullptr, * beta = nullptr, * moving_mean = nullptr, * moving_variance = nullptr;
    ggml_tensor * forward(ggml_context * ctx, ggml_tensor * input, int stride, bool use_normalization, bool use_activation,
                          bool depthwise) const;
    int in_channels() const { return (int)kernel->ne[1]; }   // file (KH,KW,IC,OC) -> ne=(OC,IC,KW,KH)
    int out_channels() const { return (int)kernel->ne[0]; }
};

struct inverted_residual {  // main.cpp:75-87
    int        strides = 1;
    conv_layer expand_1x1, conv_3x3, reduce_1x1;
    ggml_tensor * forward(ggml_context * ctx, ggml_tensor * inp) const;
};

struct transformer_layer {  // main.cpp:108-140
    ggml_tensor *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b, *up_w, *up_b, *down_w, *down_b;
    ggml_tensor *ln_before_g, *ln_before_b, *ln_after_g, *ln_after_b;
    ggml_tensor * forward(ggml_context * ctx, ggml_tensor * hidden, float eps, int num_heads) const;
};

struct vit_block {  // mobile_vit_layer, main.cpp:152-177
    inverted_residual              downsampling;
    conv_layer                     conv_kxk, conv_1x1, conv_projection, fusion;
    std::vector<transformer_layer> layers;
    ggml_tensor *                  ln_g = nullptr, * ln_b = nullptr;
    ggml_tensor * forward(ggml_context * ctx, ggml_tensor * inp, const hparams & hp) const;
};

ggml_tensor * unfolding(ggml_context * ctx, ggml_tensor * features, int patch_size);                       // main.cpp:721-747
ggml_tensor * folding(ggml_context * ctx, ggml_tensor * tokens, int patch_size, int n_patch_w, int n_patch_h,
                      int batch);                                                                          // main.cpp:750-768

struct forward_graph {
    ggml_context * ctx = nullptr;
    ggml_cgraph *  gf  = nullptr;
    ggml_tensor *  input_hwc = nullptr;  // ne = (3, W, H, N): the caller's HWC images, uploaded as they are
    ggml_tensor *  features  = nullptr;  // ne = (W/32, H/32, C, N)
    ggml_tensor *  pooled    = nullptr;  // ne = (1, 1, C, N)
    ggml_tensor *  logits    = nullptr;  // ne = (classes, N); only when the weight file carries a classifier
    uint8_t *      input_u8 = nullptr;   // pinned staging of raw u8 images (mvit_*_u8), [N][src_h][src_w][3]
    size_t         input_u8_bytes = 0;
    void *         pinned_arena = nullptr;  // page-locked backing store of ctx (input staging + output shadows)
    std::vector<ggml_tensor *> stages;   // stem, layer_1..layer_5, conv_1x1_exp outputs (debug taps, MVIT_DEBUG_STAGES=1)
    // Small requests are split into concurrent lanes (sub-batches on their own streams, see graph_for): then gf == nullptr, the
    // tensors above are plain host buffers and every lane is a complete forward graph aliasing its slice of them.
    std::vector<forward_graph> lanes;
    bool slot_upload = true, slot_download = true;  // mvit_slot_set_transfers
};

struct model {  // mobilevit_model, main.cpp:202-213
    hparams                              hp;
    conv_layer                           conv_stem, conv_1x1_exp;
    std::vector<inverted_residual>       layer_1, layer_2;  // mobile_net_layer x2 (main.cpp:89-106)
    vit_block                            layer_3, layer_4, layer_5;
    ggml_tensor *                        classifier_w = nullptr, * classifier_b = nullptr;  // optional head: (640,1000) kernel + bias
    ggml_context *                       ctx_w = nullptr;
    std::map<std::string, ggml_tensor *> tensors;
    int64_t                              total_weights = 0;
    std::map<std::tuple<int, int, int, int>, forward_graph> graphs;  // key: (n, h, w, slot)

    bool            load(const std::string & path);                  // load_model_v2, main.cpp:314-515
    forward_graph & graph_for(int n, int h, int w, int slot = 0);    // builds (once) the batched forward graph
    ggml_tensor *   build_forward(ggml_context * ctx, ggml_tensor * images_hwc, ggml_tensor ** pooled,
                                  std::vector<ggml_tensor *> * stages = nullptr) const;
    void            release(int n, int h, int w);
    ~model();
};

}  // namespace mvit
