// Internal interface of the tcgen05 3x3 convolution (conv_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

namespace ic {
namespace tc {

// kernel parameters (device pointers)
struct ConvTcParams {
    const uint8_t* weights;     // 36 stages x [2 planes][4 chunks][128 cout][8 cin] fp16
    const float* scale;         // [128] BN scale / weight pre-scale
    const float* shift;         // [128]
    const __half* res1;         // [2][N][16][H][W][8] or nullptr
    const __half* res2;
    __half* out;                // [2][N][16][H][W][8]
    int N, H, W, relu;
};

struct ConvTcArgs {
    const __half* in;           // [planes][N][16][H][W][8]
    const __half* weights;
    const float *scale, *shift;
    const __half *res1, *res2;
    __half* out;
    int N, H, W, relu;
    int exact;                  // 1: hi/lo planes, 3 MMAs per product; 0: hi plane only
};

int launch_conv3x3_tc(const ConvTcArgs& a, cudaStream_t s);
int launch_split_from_nhwc(const float* in, int N, int H, int W, __half* out, int write_lo, cudaStream_t s);
int launch_merge_to_nhwc(const __half* in, int N, int H, int W, float* out, int has_lo, cudaStream_t s);
void pack_weights_3x3(const float* w_hwio, std::vector<__half>& packed, float* inv_scale_out);

}  // namespace tc
}  // namespace ic
