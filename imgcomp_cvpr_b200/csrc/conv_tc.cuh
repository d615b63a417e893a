// Internal interface of the tcgen05 convolution (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"

namespace ic {
namespace tc {

// which taps each 64-channel input group contributes (built by pack_weights)
struct GroupTable {
    int ngroups;             // A-buffer loads per super tile
    int nstages;             // weight stages per super tile (one per (group, tap))
    int eff_ksteps;          // (stage, k-step) pairs with non-zero weights: accumulate steps per output / products
    uint8_t ntaps[16];
    uint8_t taps[16][9];     // dy*3+dx halo offsets
    uint8_t img_off[16];     // added to the image coordinate of the TMA load (0 for 2-D convs)
    uint8_t chunk0[16];      // first channel chunk of the group
};

// kernel parameters (device pointers)
struct ConvTcParams {
    const uint8_t* weights;     // nstages x [2 planes][4 chunks][NOUT rows][8 cin] fp16
    const float* scale;         // [NOUT] BN scale / weight pre-scale
    const float* shift;         // [NOUT]
    const __half* res1;         // same layout as out, or nullptr
    const __half* res2;
    __half* out;                // OUTMODE 0: [2][N][NOUT/8][H][W][8]
    float* out_f32;             // OUTMODE 1: [N][H][W][cout]
    int N, H, W, relu, cout;
    float acc_gain;             // compensates the round-toward-zero bias of the tensor core's fp32 accumulation
    int halo0;                  // origin of the halo tile relative to the output tile (-1: SAME 3x3-like, 0: VALID)
    // input image coordinate = n * img_mul + (n / img_div) * img_div_mul + img_off[group]   (img_div = 0: off)
    int img_mul, img_div, img_div_mul;
    // ... + img_base, with img_off[group] multiplied by img_off_mul (depth-major training volumes: one depth slice = N images;
    // the data gradient reads slice s - 1, i.e. img_base = -N: negative image coordinates are TMA zero fill)
    int img_off_mul, img_base;
    // resident-weight (context model) kernels with <= 24 real input channels: the third 8-channel chunk of two consecutive
    // taps forms ONE k-step (see the issue loop) instead of two half-empty ones
    int pair_c2;
    int aslots;                 // activation-tile slots in shared memory (2 ... 6)
    int pc_static;              // context model: compile-time tap schedule (1: group 1 = taps 0..4, 2: taps 4..8), 0: table-driven
    // context model, depth walk (inference layers): a CTA follows one (image, tile) along the depth axis and fetches every
    // input slice ONCE -- its A tile feeds the 5 taps of filter depth 1 (finishing output slice j - 1) and the 9 taps of
    // filter depth 0 (starting output slice j) in the same MMAs, N = 4 NOUT over two adjacent accumulators of a 4-deep
    // TMEM ring.  walk = outputs per depth segment (0: off), walk_nseg = segments per column.
    int walk, walk_nseg;
    // geometry of res1 (context model: a crop of a larger tensor); res2 always has the output geometry
    int res_H, res_W, res_dy, res_dx, res_div_mul, res_img_off;
    size_t res_plane;
    // OUTMODE 2 (context-model head): 0 logits, 1 bit cost, 2 coder frequencies (+ bit cost sums)
    int head;
    const int64_t* symbols;
    int64_t* out_freqs;
    double* bits_sum;
    int store_chunks;           // OUTMODE 0: chunks >= store_chunks are not stored (0: store all)
    int out_s2d;                // OUTMODE 0: write the output in space-to-depth form [plane][N][4*NOUT/8][H/2][W/2][8]
    // OUTMODE 0 depth-to-space (transposed convs): column block = (phase, chunk); output [plane][N][d2s_cch][2H][2W][8]
    int d2s_cch, d2s_ph0;
    // OUTMODE 3 (h13): NCHW float32 image (+ truncated uint8 copy), optional denormalise + clip
    int denorm;
    uint8_t* out_u8;
    // IC_TC_DBG=1: per CTA, cycles the MMA issuer spent waiting on [acc_empty, a_full, w_full] and its total loop time
    unsigned long long* dbg;
};

struct ConvTcArgs {
    const __half* in;           // [planes][Nimg][in_chunks][Hin][Win][8]
    int Nimg, in_chunks, Hin, Win;
    int in_chunks_valid;        // chunks that hold data (0 = in_chunks): the rest of a group's box is TMA zero fill, never read
    int store_chunks;           // OUTMODE 0: output chunks that are stored (0 = all NOUT / 8): all-zero padding chunks are skipped
    const __half* weights;
    const __half* weights_pair;   // pair-packed copy (cta_group::2 path) or nullptr
    const __half* weights_cat;    // B-concatenated copy (conv_cat_kernel) or nullptr
    const __half* weights_cat_pair;   // B-concatenated, pair-packed copy (conv_cat_kernel<PAIR>) or nullptr
    const GroupTable* groups;
    const float *scale, *shift;
    const __half *res1, *res2;
    __half* out;
    float* out_f32;
    int N, H, W;                // output tile grid
    int relu, cout, nout;       // nout: padded output channels the weights were packed for (128 or 48)
    int halo0, img_mul, img_div, img_div_mul;
    int img_off_mul, img_base;  // see ConvTcParams (img_off_mul = 0 means 1)
    int pair_c2;                // see ConvTcParams (the input has <= 24 channels: chunk 3 is all zero)
    int pc_f32;                 // context-model layer (resident weights, B-concatenation) with float32 NHWC output (training)
    int res_H, res_W, res_dy, res_dx, res_div_mul, res_img_off;
    size_t res_plane;
    int out_s2d, d2s_cch, d2s_ph0, denorm;
    uint8_t* out_u8;
    int head;                   // -1: none
    const int64_t* symbols;
    int64_t* out_freqs;
    double* bits_sum;
    int cpg;                    // chunks per group: 8 (64 channels) or 4 (context model)
    int exact;                  // 1: hi/lo planes, 3 MMAs per product; 0: hi plane only
    int prof_class;
};

int launch_conv_tc(const ConvTcArgs& a, cudaStream_t s);
int launch_split_from_nhwc(const float* in, int N, int H, int W, int C, int s2d, __half* out, int write_lo, cudaStream_t s,
                           const float* d_mul = nullptr);
int encode_planes_map(CUtensorMap* map, const __half* base, int planes, int Nimg, int chunks, int H, int W, int box_w, int box_h,
                      int box_chunks);
int launch_merge_to_nhwc(const __half* in, int N, int H, int W, int C, float* out, int has_lo, cudaStream_t s,
                         const float* d_mul = nullptr, const float* d_add = nullptr);
int launch_s2d_planes(const __half* in, int N, int H, int W, int C, int planes, __half* out, cudaStream_t s);
int pack_weights_tconv(const float* w, int k, int cin, int cout, const int* phases, int nphases, int nout,
                       std::vector<__half>& packed, GroupTable& gt, float* inv_scale_out);
int launch_split_from_nchw(const float* in, int N, int C, int H, int W, __half* out, int write_lo, cudaStream_t s);
void repack_pair(const std::vector<__half>& packed, int nstages, std::vector<__half>& out);
void repack_cat(const std::vector<__half>& packed, int nstages, std::vector<__half>& out);
void repack_cat_pair(const std::vector<__half>& packed, int nstages, std::vector<__half>& out);
int pack_weights_h1(const float* w_hwio, int cin, int cout, int nout, std::vector<__half>& packed, GroupTable& gt,
                    float* inv_scale_out);
int launch_conv_h1(const void* x, int x_is_u8, int N, int H, int W, int normalize, const __half* weights, const float* scale,
                   const float* shift, int relu, __half* out, int exact, cudaStream_t s);
int pack_weights_h1_im2col(const float* w_hwio, int cin, int cout, std::vector<__half>& packed, float* inv_scale_out);
int pack_weights_pc(const float* w, int ci, int co, int nout, std::vector<__half>& packed, GroupTable& gt,
                    float* inv_scale_out);
int pack_weights(const float* w_hwio, int k, int stride, int cin, int cout, int nout, std::vector<__half>& packed,
                 GroupTable& gt, float* inv_scale_out);

}  // namespace tc
}  // namespace ic
