/*
 * mobilevit_b200.h -- C ABI of the host-side MobileViT program (ggml-experiments_b200/host/mobilevit.cpp).
 *
 * The reference's entry points are C++ methods of `mobilevit_model` in /root/reference/mobilevit/main.cpp;
 * this header flattens them to plain C so that ctypes / cgo / JNI can bind them:
 *
 *   mvit_load              <- load_model_v2 + read_all_weights + assign_weights   (main.cpp:218-515,872-942)
 *   mvit_extract_features  <- mobilevit_model::extract_features                    (main.cpp:604-646)
 *   mvit_free              <- ggml_free(model.ctx_w)                               (main.cpp:699)
 *
 * extract_features is generalised from the reference's hard-coded (256,256,3,1) input (main.cpp:612) to a
 * batch of N images of H x W (multiples of 64).  The graph is still built from ggml_* calls
 * (include/ggml/ggml.h) and runs through ggml_graph_compute_with_ctx on the GPU; there is no CPU path.
 */
#ifndef MOBILEVIT_B200_H
#define MOBILEVIT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvit_model mvit_model;

/* Parse a weight file in the convert-tf-to-ggml.py layout. Returns NULL if the file cannot be opened or
 * is malformed; aborts (std::out_of_range in the reference, main.cpp:225) if a required tensor is missing. */
mvit_model * mvit_load(const char * weight_path);
void         mvit_free(mvit_model * m);

int     mvit_num_tensors(const mvit_model * m);   /* 313 for every variant */
int64_t mvit_num_weights(const mvit_model * m);   /* floats in the file */
int     mvit_out_channels(const mvit_model * m);  /* 640 / 384 / 320 for S / XS / XXS */

/* images_hwc: N*H*W*3 floats in [0,1], HWC per image (sam_image_f32, main.cpp:22-27).
 * features  : N*C*(H/32)*(W/32) floats, per image the ggml tensor ne=(W/32,H/32,C,1) the reference returns
 *             (main.cpp:645), i.e. [N][C][H/32][W/32]; may be NULL.
 * pooled    : N*C floats, mean of the feature map over space (the build's "logits"); may be NULL.
 * Returns 0 on success, non-zero for invalid arguments (n <= 0, H or W not a multiple of 64). */
int mvit_extract_features(mvit_model * m, const float * images_hwc, int n, int h, int w, float * features, float * pooled);

/* ---- zero-copy host variant: the reference's own flow (write inp->data, compute, read output->data;
 * main.cpp:627-634,640,703).  The host buffers belong to the library (pinned); valid until mvit_release. ---- */
float * mvit_host_input(mvit_model * m, int n, int h, int w);      /* [N,H,W,3] f32, fill in place */
int     mvit_compute(mvit_model * m, int n, int h, int w);         /* H2D + forward + D2H, synchronous */
const float * mvit_host_features(mvit_model * m, int n, int h, int w);  /* [N,C,H/32,W/32] after mvit_compute */
const float * mvit_host_pooled(mvit_model * m, int n, int h, int w);    /* [N,C] after mvit_compute */

/* ---- pipelined host variant (throughput serving): `slot` in 0..7, each with its own pinned input buffer, output buffers,
 * device arena and stream.  Fill slot s, submit it (returns immediately), fill and submit slot s+1, then wait for s:
 * the H2D/D2H copies of one slot overlap the kernels of the other.  Every submit still performs its own H2D + forward +
 * D2H; only the waiting is deferred. ---- */
float *       mvit_slot_input(mvit_model * m, int n, int h, int w, int slot);
int           mvit_slot_submit(mvit_model * m, int n, int h, int w, int slot);
int           mvit_slot_set_transfers(mvit_model * m, int n, int h, int w, int slot, int upload_inputs, int download_outputs);
int           mvit_slot_wait(mvit_model * m, int n, int h, int w, int slot);
const float * mvit_slot_features(mvit_model * m, int n, int h, int w, int slot);
const float * mvit_slot_pooled(mvit_model * m, int n, int h, int w, int slot);

/* ---- classification head.  The reference stops at the feature map (main.cpp:645); when the weight file also
 * carries `.../classifier/kernel:0` (C, classes) and `.../classifier/bias:0` (TFMobileViTForImageClassification naming, kernel
 * stored (in, out) like every dense kernel of the file) the graph ends in logits = pooled . kernel + bias. ---- */
int           mvit_num_classes(const mvit_model * m);                  /* 0: the file has no classifier */
/* logits: N*classes floats (may be NULL); top1: N class ids (may be NULL).  Returns 0, 1 (bad arguments) or 2 (no classifier). */
int           mvit_classify(mvit_model * m, const float * images_hwc, int n, int h, int w, float * logits, int32_t * top1);
const float * mvit_host_logits(mvit_model * m, int n, int h, int w);  /* [N,classes] after mvit_compute; NULL without a classifier */
const float * mvit_slot_logits(mvit_model * m, int n, int h, int w, int slot);

/* ---- uint8 images with the preprocessing on the device.  Replaces sam_image_preprocess (main.cpp:538-601) + the
 * HWC copy (main.cpp:627-634): every source image [src_h][src_w][3] u8 is resized so that its longer side fills the H x W
 * input (bilinear, the reference's arithmetic and rounding to u8), divided by 255 and written top-left into the zeroed f32
 * input; rows use stride W (the reference writes stride nx3, which shears non-square images: SURVEY App. C #3).
 * 4x fewer bytes cross PCIe than with f32 images. ---- */
uint8_t * mvit_host_input_u8(mvit_model * m, int n, int h, int w, int src_h, int src_w);   /* pinned [N,src_h,src_w,3], fill in place */
int       mvit_compute_u8(mvit_model * m, int n, int h, int w, int src_h, int src_w);      /* H2D(u8) + preprocess + forward + D2H */
uint8_t * mvit_slot_input_u8(mvit_model * m, int n, int h, int w, int slot, int src_h, int src_w);
int       mvit_slot_submit_u8(mvit_model * m, int n, int h, int w, int slot, int src_h, int src_w);
/* preprocessing alone (parity tests): images -> out_hwc [N,H,W,3] f32 on the host.  Returns 0 on success. */
int       mvit_preprocess_u8(mvit_model * m, const uint8_t * images, int n, int src_h, int src_w, int h, int w, float * out_hwc);

/* ---- device-resident variant (bench.py `value`): inputs/outputs stay in HBM -------------------------- */
int    mvit_prepare(mvit_model * m, int n, int h, int w);           /* build graph + device plan; 0 on success */
void * mvit_device_input(mvit_model * m, int n, int h, int w);      /* device ptr, [N,H,W,3] f32 */
void * mvit_device_features(mvit_model * m, int n, int h, int w);   /* device ptr, [N,C,H/32,W/32] f32 */
void * mvit_device_pooled(mvit_model * m, int n, int h, int w);     /* device ptr, [N,C] f32 */
int    mvit_forward_device(mvit_model * m, int n, int h, int w);    /* async launch, ordered on the library stream */
void   mvit_release(mvit_model * m, int n, int h, int w);           /* drop the cached graph/plan for a shape */

struct mvit_plan_info {
    int     mode;          /* 0 fast, 1 exact (enum ggml_b200_mode) */
    int     graph_nodes;
    int     launches;      /* kernel launches per forward */
    int64_t arena_bytes;   /* device activation arena after liveness planning */
    int64_t naive_bytes;   /* what keeping every intermediate alive (the reference's arena) would need */
    int64_t weight_bytes;
    int     cuda_graph;
    int     lanes;         /* concurrent sub-batch lanes the request is split into (1: one plan); launches / arena are summed over them */
};
int mvit_plan_info(mvit_model * m, int n, int h, int w, struct mvit_plan_info * out);
/* JSON array of per-launch device times, see ggml_b200_graph_profile_json. */
int mvit_profile_json(mvit_model * m, int n, int h, int w, int reps, char * buf, size_t cap);

/* Debug tap (tests only): with the environment variable MVIT_DEBUG_STAGES=1 set before the first use of a shape,
 * the outputs of stem (0), encoder layers 1..5 (1..5) and conv_1x1_exp (6) are kept as graph outputs; this copies
 * one of them as [N][C][H][W] floats and its ggml ne into ne4.  Returns the element count, <0 on error. */
int64_t mvit_debug_stage(mvit_model * m, int n, int h, int w, int idx, float * out, int64_t cap_floats, int64_t * ne4);

#ifdef __cplusplus
}
#endif
#endif
