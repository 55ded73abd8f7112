// vit_stage.cuh -- K8: every transformer layer of one MobileViT block (main.cpp:988-1172, called n times from
// mobile_vit_layer::forward, main.cpp:1196-1204) as ONE kernel launch; see vit_stage.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace b200 {

// host pointers to one layer's parameters as the weight file holds them: dense kernels are f32 [in][out] (element (k, n) at
// k * out + n, the layout match_dense() verified), biases / LayerNorm vectors f32
struct VitLayerHost {
    const float *ln1_g, *ln1_b, *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
};

struct VitStageLaunch {
    struct Params {
        int N, H, W, C, heads, d, F, L, n_layers;
        int n_seq, tiles, num_kb, NP, nch, w2, total_rows, stage_ok;
        int n_blk;                 // weight blocks per layer
        uint16_t blk_rows[64];     // their heights (rows of 128 B) in streaming order
        int   slot_bytes, ra_bytes;
        float eps, scale_log2;
        const float * x32;         // [N*H*W][C] residual stream entering the first layer
        float *       out32;       // [N*H*W][C] residual stream after the last layer (any of the three may be null)
        __half *      out16;
        float *       stats;       // [N*H*W][2] (sum, sum of squares) of the final rows: the LayerNorm that follows is folded into its consumer
        const uint8_t * blob;      // pre-tiled, pre-swizzled f16 weight blocks, layer after layer
        long long     layer_blob_bytes;
        const float * vec;         // n_layers + 1 bias blocks of vec_stride floats: [pending b2 256 | bqkv heads*3*DP | bo 256 | b1 nch*128]
        int vec_stride;
    } p;
    int    dp;
    int    grid;
    size_t smem_bytes;
};

// whether the fused kernel covers a stage of this shape (sequence length a power of two <= 64, C <= 256, head dim <= 64, ...).
// heads == 0 everywhere below selects the MLP-ONLY variant: LN -> up-projection + SiLU -> down-projection + residual of ONE layer (the second
// half of transformer_layer::forward, main.cpp:1113-1165) for stages whose sequences do not fit a tile (L = 256 / 1024); only ln2 / w1 / b1 /
// w2 / b2 of VitLayerHost are read.
bool vit_stage_supported(int N, int H, int W, int C, int heads, int F);
// constant folding (plan time, host): all layers' weights tiled into the kernel's streaming order
void vit_stage_pack(const VitLayerHost * layers, int n_layers, int C, int heads, int F, std::vector<uint8_t> & blob, std::vector<float> & vec);
bool vit_stage_prepare(VitStageLaunch & L, const float * x32, int N, int H, int W, int C, int heads, int F, int n_layers, float eps,
                       const uint8_t * blob_dev, const float * vec_dev, float * out32, __half * out16, float * stats);
void vit_stage_launch(const VitStageLaunch & L, cudaStream_t st);

}  // namespace b200
