// Shared helpers for the imgcomp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "imgcomp_b200.h"

namespace ic {

void set_error(const char* fmt, ...);

#define IC_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ic::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return IC_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define IC_CHECK_LAUNCH() IC_CHECK_CUDA(cudaPeekAtLastError())

#define IC_REQUIRE(cond, code, ...)       \
    do {                                  \
        if (!(cond)) {                    \
            ic::set_error(__VA_ARGS__);   \
            return (code);                \
        }                                 \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// TF 'SAME' padding: out = ceil(n/s), pad_before = total / 2.
static inline int same_pad_before(int n, int k, int s) {
    int out = (n + s - 1) / s;
    int total = (out - 1) * s + k - n;
    if (total < 0) total = 0;
    return total / 2;
}

// launch accounting (profile.cu): counts launches; when profiling is enabled, brackets the
// scope with CUDA events on `s`.
struct ProfScope {
    ProfScope(int cls, cudaStream_t s, int n_launches = 1);
    ~ProfScope();
    int cls_;
    cudaStream_t s_;
    bool on_;
    cudaEvent_t a_, b_;
};

// bump allocator over the caller's workspace
struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* p, size_t c) : base((char*)p), cap(c), off(0) {}
    template <typename T>
    T* get(size_t n) {
        size_t o = align_up(off, 256);
        off = o + n * sizeof(T);
        return (T*)(base + o);     // validity checked by ok()
    }
    bool ok() const { return off <= cap; }
};

// ------------------------------------------------------------------ SIMT conv
// Generic float32 NHWC implicit-GEMM convolution / transposed convolution with a
// fused epilogue (BN scale+shift, ReLU, up to two residual adds, optional
// denormalise+clip+NCHW store).  This is the IC_MODE_FP32 path and the on-device
// reference the tensor-core kernels are validated against.
struct ConvDesc {
    const float* in;      // N,Hi,Wi,Cin   (Cin % 4 == 0)
    const float* w;       // [KH*KW*Cin][ldw] row-major, ldw % 4 == 0, zero padded
    const float* scale;   // [ldw]
    const float* shift;   // [ldw]
    const float* res1;    // N,Ho,Wo,Cout or nullptr
    const float* res2;    // N,Ho,Wo,Cout or nullptr
    float* out;           // N,Ho,Wo,Cout (NHWC) or N,Cout,Ho,Wo when out_nchw
    uint8_t* out_u8;      // optional truncated copy (NCHW), only with out_nchw
    int N, Hi, Wi, Cin, Ho, Wo, Cout, ldw;
    int KH, KW, stride, pad_t, pad_l;
    int transposed;       // 0: iy = oy*stride - pad_t + ky ; 1: iy = (oy + pad_t - ky) / stride
    int relu;
    int out_nchw;         // 1: denormalise (FIXED) + clip[0,255] + NCHW store  (h13 epilogue)
    int denorm;
};
int launch_conv_simt(const ConvDesc& d, cudaStream_t stream);

// elementwise helpers (elementwise.cu)
int launch_prep_input(const void* x, int is_u8, int N, int H, int W, int normalize, float* out_nhwc4,
                      cudaStream_t s);
int launch_prep_input_s2d(const void* x, int is_u8, int N, int H, int W, int normalize, __half* out, int write_lo,
                          cudaStream_t s);
int launch_nchw_to_nhwc(const float* in, int N, int C, int H, int W, float* out, cudaStream_t s);
int launch_heatmap_quantize(const float* bn_nhwc, int N, int h, int w, int C, int heatmap,
                            const float* centers, int L,
                            float* z, float* hm, float* qbar, float* qhard, int64_t* sym, uint8_t* sym8,
                            float* qsoft, cudaStream_t s, int cb_stride = 0);
int launch_quantize(const float* x, const float* centers, int L, float sigma, int64_t n,
                    float* qsoft, float* qhard, int64_t* sym, cudaStream_t s);

// ms-ssim (msssim.cu)
size_t msssim_bwd_workspace_bytes(int N, int H, int W);
int msssim_tf_bwd(const float* a, const float* b, int N, int H, int W, float grad_out, float* d_b, float* value_out, void* ws,
                  size_t ws_bytes, cudaStream_t s);

}  // namespace ic
