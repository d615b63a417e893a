/*
 * imgcomp_b200 -- C ABI of the B200-native hot path of fab-jul/imgcomp-cvpr.
 *
 * The reference has no FFI: its boundary is the Python object API that
 * code/val.py and code/train.py call while building a TF graph.
 * Each entry point below replaces the *execution* of one of those calls; the
 * Python mirror classes in imgcomp_cvpr_b200/ keep the reference's names and
 * argument order and forward to these functions through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ic_status otherwise;
 *     ic_last_error() gives a thread-local message.
 *   - pointers named d_* are DEVICE pointers (owned by the caller, e.g. torch
 *     tensors), h_* are HOST pointers.  No torch / C++ types cross the boundary.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     nothing synchronises unless stated.
 *   - layouts at the boundary are the reference's: images / latents NCHW
 *     float32, symbols int64, conv2d weights HWIO, conv2d_transpose weights
 *     [kh,kw,Cout,Cin], conv3d weights [D,H,W,in,out].
 *   - handles own only the (re-laid-out) weights; activations live in a caller
 *     provided workspace sized by the matching *_workspace_bytes().
 *   - handles are immutable after creation: any number of threads/streams may
 *     use one handle concurrently with distinct workspaces.
 */
#ifndef IMGCOMP_B200_H_
#define IMGCOMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IC_ABI_VERSION 1

typedef enum {
    IC_OK = 0,
    IC_ERR_INVALID = -1,     /* bad argument (shape, dtype, NULL, not multiple of 8 ...) */
    IC_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
    IC_ERR_WORKSPACE = -3,   /* workspace too small */
    IC_ERR_UNSUPPORTED = -4, /* configuration outside the hot path (arch, kernel size ...) */
    IC_ERR_STATE = -5        /* call order violated (e.g. decoder not fed) */
} ic_status;

/* Precision mode of the tensor-core convolutions (ic_encode_fwd/ic_decode_fwd):
 *   IC_MODE_FP32   plain float32 FFMA kernels (reference arithmetic class)
 *   IC_MODE_EXACT  tcgen05 fp16 x3 split (hi*hi + hi*lo + lo*hi, fp32 accumulate):
 *                  float32-class results, the mode the parity gate runs in
 *   IC_MODE_FAST   tcgen05 single fp16 pass (reported with its symbol flip rate) */
typedef enum { IC_MODE_FP32 = 0, IC_MODE_EXACT = 1, IC_MODE_FAST = 2 } ic_mode;

const char* ic_last_error(void);
int ic_abi_version(void);
/* 1 if a CUDA device with compute capability 10.x is usable from this process. */
int ic_device_ok(void);

/* ------------------------------------------------------------------ autoencoder
 * replaces: autoencoder.get_network_cls(config)(config)      code/autoencoder.py:26-43
 * config attributes used by the reference: arch, num_chan_bn, heatmap,
 * normalization, arch_param_B, num_centers (code/autoencoder.py:32-43,218-268). */
typedef struct ic_ae ic_ae_t;

typedef struct {
    int32_t num_chan_bn;      /* C: 32 (cvpr/low, med) or 64 (cvpr/hi) */
    int32_t arch_param_B;     /* 5 */
    int32_t num_centers;      /* L = 6 */
    int32_t heatmap;          /* 1 (ae_configs/base:5) */
    int32_t normalization;    /* 1 = FIXED, 0 = OFF (ae_configs/base:3-4) */
} ic_ae_config;

/* Weight tensors are passed as an array of host pointers in a fixed order;
 * ic_ae_tensor_name(i) is the TF variable name of entry i,
 * ic_ae_tensor_numel(i) its element count.  i in [0, ic_ae_num_tensors). */
int ic_ae_num_tensors(const ic_ae_config* cfg);
const char* ic_ae_tensor_name(const ic_ae_config* cfg, int i);
int64_t ic_ae_tensor_numel(const ic_ae_config* cfg, int i);
int ic_ae_create(const ic_ae_config* cfg, const float* const* h_tensors, int n_tensors, ic_ae_t** out);
void ic_ae_destroy(ic_ae_t* ae);

/* replaces: ae.encode(x, is_training=False) -> EncoderOutput(qbar,qhard,symbols,z,heatmap)
 *           code/autoencoder.py:50-58,218-244 (+ quantizer.py:37-95 inside)
 * d_x: N x 3 x H x W, float32 in [0,255] (x_is_u8 = 0) or uint8 (x_is_u8 = 1; the
 *      tf.to_float of val.py:83 is fused).  H, W multiples of 8 (val.py:157 pads).
 * outputs N x C x H/8 x W/8; any output pointer may be NULL to skip it.
 * d_symbols_u8 is an extra compact copy of `symbols` for the context model. */
size_t ic_encode_workspace_bytes(const ic_ae_t* ae, int N, int H, int W, int mode);
int ic_encode_fwd(const ic_ae_t* ae, const void* d_x, int x_is_u8, int N, int H, int W,
                  float* d_z, float* d_heatmap, float* d_qbar, float* d_qhard,
                  int64_t* d_symbols, uint8_t* d_symbols_u8, float* d_qsoft,
                  void* d_workspace, size_t workspace_bytes, int mode, void* stream);

/* replaces: ae.decode(q, is_training=False) -> x_out      code/autoencoder.py:60-63,246-268
 * d_q: N x C x h x w float32; d_x_out: N x 3 x 8h x 8w float32, clipped to [0,255];
 * d_x_out_u8 (optional): tf.cast(x_out, uint8) of val.py:91 (truncation). */
size_t ic_decode_workspace_bytes(const ic_ae_t* ae, int N, int h, int w, int mode);
int ic_decode_fwd(const ic_ae_t* ae, const float* d_q, int N, int h, int w,
                  float* d_x_out, uint8_t* d_x_out_u8,
                  void* d_workspace, size_t workspace_bytes, int mode, void* stream);

/* The centres variable, copied to a device buffer of num_centers floats
 * (ae.get_centers_variable(), code/autoencoder.py:65-68). */
int ic_ae_centers(const ic_ae_t* ae, float* d_centers, void* stream);

/* replaces: quantizer.quantize(x, centers, sigma) -> (qsoft, qhard, symbols)
 *           code/quantizer.py:37-95.  Elementwise over n values; layout-agnostic. */
int ic_quantize_fwd(const float* d_x, const float* d_centers, int L, float sigma, int64_t n,
                    float* d_qsoft, float* d_qhard, int64_t* d_symbols, void* stream);

/* ------------------------------------------------------------------- probclass
 * replaces: probclass.get_network_cls(pc_config)(pc_config, num_centers)
 *           code/probclass.py:11-15,30-41 ('res_shallow', kernel_size 3). */
typedef struct ic_pc ic_pc_t;

typedef struct {
    int32_t kernel_size;      /* 3 (pc_configs/cvpr/res_shallow:4) */
    int32_t arch_param_k;     /* 24 (pc_configs/base:19) or 64 */
    int32_t num_centers;      /* L = 6 */
} ic_pc_config;

int ic_pc_num_tensors(const ic_pc_config* cfg);                 /* 8: 4 x (weights, biases) */
const char* ic_pc_tensor_name(const ic_pc_config* cfg, int i);
int64_t ic_pc_tensor_numel(const ic_pc_config* cfg, int i);
int ic_pc_create(const ic_pc_config* cfg, const float* const* h_tensors, int n_tensors, ic_pc_t** out);
void ic_pc_destroy(ic_pc_t* pc);

/* replaces: pc.bitcost(q, target_symbols, is_training=False, pad_value)   code/probclass.py:63-106
 * d_q N x C x h x w float32, d_symbols int64 (same shape) -> d_bits float32 (same shape).
 * d_bits_sum (optional, N doubles): per-image sum of bits (feeds bits.bitcost_to_bpp,
 * code/bits.py:4-14). */
size_t ic_pc_workspace_bytes(const ic_pc_t* pc, int N, int D, int H, int W);  /* D,H,W of the q volume */
int ic_pc_bitcost_fwd(const ic_pc_t* pc, const float* d_q, const int64_t* d_symbols, float pad_value,
                      int N, int C, int h, int w, float* d_bits, double* d_bits_sum,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* replaces: pc.logits(q, is_training=False) on an UN-padded N x D x H x W (x1) volume
 *           code/probclass.py:130-135 -> N x (D-4) x (H-8) x (W-8) x L float32. */
int ic_pc_logits_fwd(const ic_pc_t* pc, const float* d_q, int N, int D, int H, int W, float* d_logits,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* replaces the per-symbol loop of probclass.PredictionNetwork.get_freqs
 *           (code/probclass.py:441-476, driven by code/bit_counter.py:103-134):
 * ONE batched pass: symbols (N x C x h x w, int64) are padded with symbol 0,
 * gathered through `centers`, run through the context model; output
 * N x C x h x w x L int64 = max(int64(softmax * 1e9), 1) in the coder's raster order.
 * d_bits_sum (optional, N doubles) = sum -log2 p[symbol] (ProbclassNetworkTesting,
 * code/probclass.py:393-421). */
int ic_pc_freqs_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* d_centers,
                    int N, int C, int h, int w, int64_t* d_freqs, double* d_bits_sum,
                    void* d_workspace, size_t workspace_bytes, void* stream);

/* Codec pair for a real bitstream (the decoder half of code/bit_counter.py:137-163).
 *
 * ic_pc_codec_freqs_fwd: same contract as ic_pc_freqs_fwd, but always computed by the float32
 * kernels whose operation order the sequential decoder reproduces: the tables an encoder
 * feeds to the range coder must be the tables the decoder will derive, bit for bit.
 *
 * ic_pc_decode_fwd replaces `_decode` (code/bit_counter.py:137-163; README.md:65 "~200 s"):
 * N independent bitstreams (concatenated in d_stream, stream n = bytes
 * [d_stream_offsets[n], d_stream_offsets[n+1])), each written by ArithmeticEncoder
 * (code/arithmetic_coding.py:127-159) over the symbols of one C x h x w volume in raster
 * order, first symbol passed as side information (d_first_sym, code/bit_counter.py:118-121).
 * One CTA per image decodes symbol by symbol: context model with cached activations
 * (README.md:72-73) + range decoder on the device.  -> d_symbols N x C x h x w uint8.
 * Debug hooks: d_force_symbols (same shape, optional) replaces range decoding by teacher
 * forcing; d_freqs_seen (optional, N x C x h x w x L int64) receives every table used. */
int ic_pc_codec_freqs_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* d_centers,
                          int N, int C, int h, int w, int64_t* d_freqs, double* d_bits_sum,
                          void* d_workspace, size_t workspace_bytes, void* stream);
/* The same codec tables as 32-bit words (every entry is <= 1e9 < 2^30, code/probclass.py:443-444,473): half the bytes
 * that cross PCIe per image (4.7 MB instead of 9.4 MB for a Kodak latent).  Takes the centres as a HOST array
 * (L floats), so the call neither reads back nor synchronises: a compress pipeline can enqueue the next batch while
 * host threads code the previous one (code/bit_counter.py:103-134 is the loop this feeds). */
int ic_pc_codec_freqs_u32_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* h_centers,
                              int N, int C, int h, int w, uint32_t* d_freqs, double* d_bits_sum,
                              void* d_workspace, size_t workspace_bytes, void* stream);
size_t ic_pc_decode_workspace_bytes(const ic_pc_t* pc, int N, int C, int h, int w);
int ic_pc_decode_fwd(const ic_pc_t* pc, const uint8_t* d_stream, const int64_t* d_stream_offsets,
                     const int32_t* d_first_sym, const float* d_centers, int N, int C, int h, int w,
                     uint8_t* d_symbols, const uint8_t* d_force_symbols, int64_t* d_freqs_seen,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* replaces ONE PredictionNetwork.get_freqs call (code/probclass.py:441-444,461-476): an UN-padded
 * block of context symbols N x D x H x W (the reference feeds D,H,W = 5,9,9) is gathered through
 * `centers`, run through the context model -> int64 freqs N x (D-4) x (H-8) x (W-8) x L.  Same
 * arithmetic as ic_pc_freqs_fwd: the table of a position is bit-identical either way. */
int ic_pc_context_freqs_fwd(const ic_pc_t* pc, const int64_t* d_ctx_symbols, const float* d_centers,
                            int N, int D, int H, int W, int64_t* d_freqs,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------- training primitives
 * What tf.gradients and the two Adam optimisers of code/train.py:339-349 execute for the graph of
 * code/train.py:86-132 (cfg 3), as explicit float32 forward / backward kernels; the host side
 * (imgcomp_cvpr_b200/train.py) strings them together in the order of the reference graph.
 * Activations NHWC float32 with channel counts padded to a multiple of 4; convolution weights
 * [KH][KW][Cin][Cout] in the orientation of the op (slim.conv2d's HWIO; for slim.conv2d_transpose the
 * TF variable [kh][kw][out][in] with its last two axes swapped).  valid = 0: TF SAME padding
 * (SURVEY.md A.2), transposed = 1: conv2d_transpose (output 2x); valid = 1: VALID, forward kind only. */
size_t ic_nn_conv2d_workspace_bytes(int N, int Hi, int Wi, int Cin, int KH, int KW, int stride, int Cout,
                                    int transposed, int valid);
/* slim.conv2d / slim.conv2d_transpose without normaliser or activation (code/autoencoder.py:222-237,251-265) */
int ic_nn_conv2d_fwd(const float* d_x, const float* d_w, int N, int Hi, int Wi, int Cin, int KH, int KW,
                     int stride, int Cout, int transposed, int valid, float* d_y, void* stream);
/* Conv2DBackpropInput / its transposed twin: d_dy (output shape of the op) -> d_dx (input shape) */
int ic_nn_conv2d_bwd_data(const float* d_dy, const float* d_w, int N, int Hi, int Wi, int Cin, int KH, int KW,
                          int stride, int Cout, int transposed, int valid, float* d_dx,
                          void* d_workspace, size_t workspace_bytes, void* stream);
/* Conv2DBackpropFilter: d_x (input), d_dy (output gradient) -> d_dw [KH][KW][Cin][Cout] */
int ic_nn_conv2d_bwd_filter(const float* d_x, const float* d_dy, int N, int Hi, int Wi, int Cin, int KH, int KW,
                            int stride, int Cout, int transposed, int valid, float* d_dw,
                            void* d_workspace, size_t workspace_bytes, void* stream);
/* The 3x3 stride-1 SAME 128 -> 128 convolutions of the residual trunks (code/autoencoder.py:274-287) on the tcgen05
 * kernel in EXACT (fp16 hi/lo, fp32 accumulate) arithmetic, float32 NHWC in / out: data_grad = 0 is ic_nn_conv2d_fwd,
 * data_grad = 1 is ic_nn_conv2d_bwd_data (d_x = the output gradient) for that geometry.  d_w: float32 [3][3][128][128]
 * on the device; it is re-packed for the kernel on the device at every call (weights change every step). */
size_t ic_nn_conv3x3_tc_workspace_bytes(int N, int H, int W);
int ic_nn_conv3x3_tc(const float* d_x, const float* d_w, int N, int H, int W, int data_grad, float* d_y,
                     void* d_workspace, size_t workspace_bytes, void* stream);
/* Backward of y = conv3x3(x, w) for the same geometry, both gradients on tcgen05 (EXACT arithmetic; x, dy and w are
 * pre-scaled by per-tensor powers of two so that small gradients keep float32-class precision in the fp16 hi/lo split):
 * d_dx (optional) = ic_nn_conv2d_bwd_data, d_dw = ic_nn_conv2d_bwd_filter [3][3][128][128].  The filter gradient is a
 * GEMM over pixels on MN-major operands (see csrc/train_tc.cu) with a fixed-order reduction over pixel splits. */
/* ic_nn_conv3x3_tc with the forward pass kept for the backward pass: d_x_planes_keep (optional, 2*N*H*W*128 fp16
 * elements, device memory) receives the pre-scaled hi/lo planes of x, d_scales_keep (optional, 4 floats) {scale_w, scale_x, 1/scale_w,
 * 1/scale_x}.  Handing both to ic_nn_conv3x3_tc_bwd_ex saves it the maximum search over w and x and the split of x
 * (tf.gradients reuses the forward activations the same way; code/train.py:339-349). */
int ic_nn_conv3x3_tc_ex(const float* d_x, const float* d_w, int N, int H, int W, int data_grad, float* d_y,
                        void* d_x_planes_keep, float* d_scales_keep, void* d_workspace, size_t workspace_bytes,
                        void* stream);
size_t ic_nn_conv3x3_tc_bwd_workspace_bytes(int N, int H, int W);
int ic_nn_conv3x3_tc_bwd(const float* d_x, const float* d_dy, const float* d_w, int N, int H, int W, float* d_dx, float* d_dw,
                         void* d_workspace, size_t workspace_bytes, void* stream);
int ic_nn_conv3x3_tc_bwd_ex(const float* d_x, const float* d_dy, const float* d_w, int N, int H, int W, float* d_dx,
                            float* d_dw, const void* d_x_planes, const float* d_scales, void* d_workspace,
                            size_t workspace_bytes, void* stream);
/* The other convolutions of the training step on the tcgen05 kernels (EXACT arithmetic), forward and data gradient, with
 * the weights re-packed on the device per call through a per-layer index map (built once: a "plan"):
 *   op_kind 0  slim.conv2d 5x5 stride 2 SAME        (code/autoencoder.py:223 h2)
 *   op_kind 1  slim.conv2d_transpose 5x5 stride 2   (code/autoencoder.py:264-265 h12, h13)
 *   op_kind 3  masked (2,3,3) VALID conv3d          (code/probclass.py:227-261; "other" mask, depth-major volume (D,N,H,W,C))
 * data_grad = 1 plans the gradient w.r.t. the op's input.  d_w is the op's float32 weight array as ic_nn_conv2d_* take it
 * ([5][5][ceil4 Cin][ceil4 Cout], [2][3][3][ceil4 Cin][ceil4 Cout]).  ic_nn_tc_plan_create returns IC_ERR_UNSUPPORTED for
 * shapes without a tensor-core kernel (h1, from_bn, the gradients of to_bn, the context model's first layer): use
 * ic_nn_conv2d_* there.
 * ic_nn_tc_plan_run: D, N, H, W are the dimensions of d_x (D only for op_kind 3); d_y is (N, H/2, W/2, C') / (N, 2H, 2W, C')
 * / (D-1, N, H-2, W-2, C') forward, (D+1, N, H+2, W+2, C') context-model data gradient. */
typedef struct ic_tc_plan ic_tc_plan_t;
int ic_nn_tc_plan_create(int op_kind, int data_grad, int op_cin, int op_cout, ic_tc_plan_t** out);
void ic_nn_tc_plan_destroy(ic_tc_plan_t* plan);
int64_t ic_nn_tc_plan_map(const ic_tc_plan_t* plan, int* h_map_out, int64_t capacity);
size_t ic_nn_tc_plan_workspace_bytes(const ic_tc_plan_t* plan, int D, int N, int H, int W);
int ic_nn_tc_plan_run(const ic_tc_plan_t* plan, const float* d_x, const float* d_w, int D, int N, int H, int W, float* d_y,
                      void* d_workspace, size_t workspace_bytes, void* stream);
/* Filter gradients of the same ops on tcgen05 (a GEMM whose reduction runs over pixels, MN-major operands; the stride-2 ops
 * through the space-to-depth form of their finer tensor).  d_x: the op's input, d_dy: the gradient w.r.t. its output,
 * D/N/H/W: dimensions of d_x; d_dw in the layout of the op's weight array (masked context-model taps: 0).
 * IC_ERR_UNSUPPORTED from _create: use ic_nn_conv2d_bwd_filter. */
typedef struct ic_tc_wgrad_plan ic_tc_wgrad_plan_t;
int ic_nn_tc_wgrad_plan_create(int op_kind, int op_cin, int op_cout, ic_tc_wgrad_plan_t** out);
void ic_nn_tc_wgrad_plan_destroy(ic_tc_wgrad_plan_t* plan);
size_t ic_nn_tc_wgrad_plan_workspace_bytes(const ic_tc_wgrad_plan_t* plan, int D, int N, int H, int W);
int ic_nn_tc_wgrad_plan_run(const ic_tc_wgrad_plan_t* plan, const float* d_x, const float* d_dy, int D, int N, int H, int W,
                            float* d_dw, void* d_workspace, size_t workspace_bytes, void* stream);
/* slim.batch_norm(is_training=True, fused) (code/autoencoder.py:115-125): batch mean / biased variance over
 * the M = N*H*W rows, out = relu?((x - mean) * invstd * gamma + beta) (+ res1) (+ res2); d_mean / d_invstd are
 * kept for the backward pass; d_mov_mean / d_mov_var (optional) get the decay-0.9 moving-average update with
 * the unbiased variance.  use_stats = 0: plain affine layer with caller-supplied d_mean / d_invstd (bias + ReLU
 * of the context model: mean 0, invstd 1, gamma 1, beta = bias). */
size_t ic_nn_bn_workspace_bytes(int64_t M, int C);
int ic_nn_bn_train_fwd(const float* d_x, int64_t M, int C, const float* d_gamma, const float* d_beta, float eps,
                       int relu, int use_stats, const float* d_res1, const float* d_res2, float* d_mean,
                       float* d_invstd, float* d_mov_mean, float* d_mov_var, float* d_out,
                       void* d_workspace, size_t workspace_bytes, void* stream);
/* The trunk layers of the training step fused: the batch norm writes the fp16 hi/lo planes the next 3x3 conv reads
 * (d_planes_out), the conv's output pass accumulates the next batch norm's statistics (d_bn_partial -> d_partial_in), and the
 * power-of-two weight scales of all trunk convs come from ONE launch per step (ic_nn_weight_scales; d_scales n x 4 floats,
 * row i = {scale, 1, 1/scale, 1} = the d_scales argument of ic_nn_conv3x3_tc_bwd_ex).  Same arithmetic as the unfused
 * sequence ic_nn_conv3x3_tc + ic_nn_bn_train_fwd, except that the activations are split unscaled (as at inference). */
int ic_nn_weight_scales(const float* d_base, const int64_t* d_offsets, int n, int64_t count, float* d_scales, void* stream);
size_t ic_nn_conv3x3_tc_fused_workspace_bytes(int N, int H, int W);
size_t ic_nn_bn_partial_bytes(int64_t M);
/* ic_nn_pack3x3_all: the packed weight images of all n trunk convs, forward (entry 2 i) and data gradient (2 i + 1), in one
 * launch per step; entries of ic_nn_conv3x3_tc_prepared_bytes() bytes, handed to the two calls below as d_prepared
 * (optional: NULL = pack per call). */
size_t ic_nn_conv3x3_tc_prepared_bytes(void);
int ic_nn_pack3x3_all(const float* d_base, const int64_t* d_offsets, const float* d_scales, int n, int W, void* d_prepared, void* stream);
int ic_nn_conv3x3_tc_fused(const void* d_x_planes, const float* d_w, const float* d_wscale4, const void* d_prepared, int N, int H,
                           int W, float* d_y, double* d_bn_partial, void* d_workspace, size_t workspace_bytes, void* stream);
int ic_nn_bn_train_fwd_ex(const float* d_x, int64_t M, int C, const float* d_gamma, const float* d_beta, float eps, int relu,
                          int use_stats, const float* d_res1, const float* d_res2, float* d_mean, float* d_invstd,
                          float* d_mov_mean, float* d_mov_var, float* d_out, const double* d_partial_in, void* d_planes_out,
                          int64_t hw, void* d_workspace, size_t workspace_bytes, void* stream);
int ic_nn_bn_train_bwd(const float* d_x, const float* d_dy, int64_t M, int C, const float* d_gamma,
                       const float* d_beta, int relu, int use_stats, const float* d_mean, const float* d_invstd,
                       float* d_dx, float* d_dgamma, float* d_dbeta,
                       void* d_workspace, size_t workspace_bytes, void* stream);
/* Backward of a trunk layer fused the same way: the batch norm's dx goes out as pre-scaled fp16 planes (d_dx_planes, C = 128,
 * batch statistics; d_scale_out = {s, 1/s}, s = the power of two that brings an upper bound of max |dx| -- from per-channel
 * maxima gathered in the statistics pass -- into [2^5, 2^6); d_dx NULL), and ic_nn_conv3x3_tc_bwd_planes consumes them. */
int ic_nn_bn_train_bwd_ex(const float* d_x, const float* d_dy, int64_t M, int C, const float* d_gamma, const float* d_beta, int relu,
                          int use_stats, const float* d_mean, const float* d_invstd, float* d_dx, float* d_dgamma, float* d_dbeta,
                          void* d_dx_planes, float* d_scale_out, int64_t hw, void* d_workspace, size_t workspace_bytes,
                          void* stream);
int ic_nn_conv3x3_tc_bwd_planes(const void* d_dy_planes, const float* d_dy_scale, const float* d_w, const void* d_prepared_dgrad,
                                int N, int H, int W, float* d_dx,
                                const float* d_dx_add /* optional: d_dx = gradient + d_dx_add */, float* d_dw,
                                const void* d_x_planes, const float* d_scales, void* d_workspace, size_t workspace_bytes,
                                void* stream);
/* backward of _get_heatmap3D + _mask_with_heatmap + the soft quantizer with
 * qbar = qsoft + stop_gradient(qhard - qsoft) (code/autoencoder.py:127-134,171-200; quantizer.py:60-100):
 * d_bn N,h,w,Cb (channel 0 = heatmap logit, 1..C = features), d_dq N,h,w,C = gradient w.r.t. qbar,
 * d_dhm N,C,h,w (optional) = gradient w.r.t. heatmap3D (from H_mask, code/train.py:311-314)
 * -> d_dbn N,h,w,Cb, d_dcenters (L). */
size_t ic_nn_hq_workspace_bytes(int64_t npix);
int ic_nn_hq_bwd(const float* d_bn, int N, int h, int w, int C, int Cb, int heatmap, const float* d_centers, int L,
                 const float* d_dq, const float* d_dhm, float* d_dbn, float* d_dcenters,
                 void* d_workspace, size_t workspace_bytes, void* stream);
/* _denormalize + _clip_to_image_range (code/autoencoder.py:146-158) on the decoder's last NHWC (4-channel) map
 * -> x_out NCHW, and the gradient back (passes where the un-clipped value is inside [0,255]). */
int ic_nn_denorm_clip_fwd(const float* d_v_nhwc4, int N, int H, int W, float* d_x_out_nchw, void* stream);
int ic_nn_denorm_clip_bwd(const float* d_v_nhwc4, const float* d_dx_out_nchw, int N, int H, int W,
                          float* d_dv_nhwc4, void* stream);
/* layout changes between the reference's NCHW tensors and the NHWC maps of the kernels (Cs = padded channels) */
int ic_nn_nhwc_to_nchw(const float* d_in, int N, int C, int Cs, int64_t hw, float* d_out, void* stream);
int ic_nn_nchw_to_nhwc(const float* d_in, int N, int C, int Cs, int64_t hw, float* d_out, void* stream);
/* out = a * x + b * y (y optional; out may alias), out = x * y */
int ic_nn_axpby(float a, const float* d_x, float b, const float* d_y, int64_t n, float* d_out, void* stream);
int ic_nn_mul(const float* d_x, const float* d_y, int64_t n, float* d_out, void* stream);
/* tf.train.AdamOptimizer._apply_dense on a flat tensor (code/training_helpers.py:38-48): g = grad + l2 * w
 * (slim l2 regulariser, code/autoencoder.py:101-102), optional 0/1 mask on g, step counts from 1. */
int ic_nn_adam_step(float* d_w, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                    float beta2, float eps, int64_t step, float l2, const float* d_mask, void* stream);
/* same update, bias-corrected step size lr * sqrt(1 - beta2^t) / (1 - beta1^t) read from d_lr_t[0] (CUDA-graph friendly) */
int ic_nn_adam_step_dev(float* d_w, const float* d_grad, float* d_m, float* d_v, int64_t n, const float* d_lr_t, float beta1,
                        float beta2, float eps, float l2, const float* d_mask, void* stream);

/* _normalize (code/autoencoder.py:136-144,160-169): x NCHW (uint8 or float32 in [0,255]) -> normalised NHWC with 4
 * channels (the 4th is zero padding) */
int ic_nn_normalize_fwd(const void* d_x_nchw, int x_is_u8, int N, int H, int W, float* d_out_nhwc4, void* stream);
/* forward twin of ic_nn_hq_bwd (code/autoencoder.py:127-134,171-200): d_bn N,h,w,Cb -> NCHW z, heatmap3D, qbar, qhard,
 * qsoft (each optional) and int64 symbols */
int ic_nn_hq_fwd(const float* d_bn, int N, int h, int w, int C, int Cb, int heatmap, const float* d_centers, int L, float* d_z,
                 float* d_heatmap, float* d_qbar, float* d_qhard, float* d_qsoft, int64_t* d_symbols, void* stream);
/* Context model in training mode (code/probclass.py:63-106 with is_training=True).  The padded volume is depth-major,
 * (C+4, N, h+8, w+8, 4 channels: value, 0, 0, 0), so that the (2,3,3) VALID conv3d (code/probclass.py:227-261) is two
 * VALID ic_nn_conv2d_* passes over the contiguous slice ranges [0, D-1) and [1, D), accumulated with ic_nn_axpby.
 * ic_nn_pc_pad_fwd = pad_for_probclass3d (code/probclass.py:268-292). */
int ic_nn_pc_pad_fwd(const float* d_q_nchw, int N, int C, int h, int w, float pad_value, const float* d_pad_value, float* d_out,
                     void* stream);   /* d_pad_value (optional): the pad value is read from device memory (centers[0]) */
/* softmax_cross_entropy_with_logits * log2(e) (code/probclass.py:99-104).  d_logits: rows in (C, N, h, w) order, Cs
 * floats per row, the first L valid; d_symbols / d_heatmap / d_bc_nchw in the reference's NCHW order.
 * backward: d_dlogits[row][k] = (coef_real + coef_mask * heatmap) * log2(e) * (softmax_k - [k == symbol]); these are
 * the gradients of beta * max(0.5 * (mean(bc * heatmap) + mean(bc)) - H_target, 0) (code/train.py:309-316) with
 * coef_* = beta * 0.5 / numel when the hinge is active, else 0. */
int ic_nn_pc_xent_fwd(const float* d_logits, int Cs, int L, const int64_t* d_symbols, int N, int C, int h, int w, float* d_bc_nchw,
                      void* stream);
int ic_nn_pc_xent_bwd(const float* d_logits, int Cs, int L, const int64_t* d_symbols, const float* d_heatmap, int N, int C, int h,
                      int w, float coef_real, float coef_mask, const float* d_coef, float* d_dlogits, void* stream);
/* the hinge coefficient on the device (so that a training step needs no host read-back in its middle and can be
 * captured in a CUDA graph): d_coef[0] = beta * 0.5 / n if 0.5 * (sums[1] / n + sums[0] / n) > H_target else 0, with
 * d_sums from ic_masked_sums_fwd; pass d_coef to ic_nn_pc_xent_bwd (overrides coef_real / coef_mask) */
int ic_nn_rate_coef(const double* d_sums, int64_t n, float beta, float h_target, int has_heatmap, float* d_coef, void* stream);
/* out = d_a[0] * x, scalar in device memory */
int ic_nn_scale_dev(const float* d_a, const float* d_x, int64_t n, float* d_out, void* stream);
/* the residual crop of the 3-D residual block, x[:, 2:, 2:-2, 2:-2, :] (code/probclass.py:185-196), on A slices of
 * H x W x C (C % 4 == 0): out = in[:, crop:-crop, crop:-crop, :]; backward ACCUMULATES into d_dx */
int ic_nn_crop_fwd(const float* d_in, int64_t A, int H, int W, int C, int crop, float* d_out, void* stream);
int ic_nn_crop_bwd_add(const float* d_dy, int64_t A, int H, int W, int C, int crop, float* d_dx, void* stream);

/* --------------------------------------------------------------------- MS-SSIM
 * replaces: ms_ssim.MultiScaleSSIM(img1, img2, data_format='NCHW')   code/ms_ssim.py:115-186
 * float32, ONE scalar for the batch.  d_out: 1 float; d_levels (optional): 10 floats
 * = ssim_l (5) then cs_l (5).  Returns IC_ERR_INVALID where the reference raises
 * (a level whose height is below the blur's tap count, code/ms_ssim.py:24-29). */
size_t ic_msssim_workspace_bytes(int N, int H, int W, int is_double);
int ic_msssim_tf_fwd(const float* d_img1, const float* d_img2, int N, int H, int W,
                     float* d_out, float* d_levels,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* gradient of ms_ssim.MultiScaleSSIM w.r.t. img2 (what tf.gradients derives from code/ms_ssim.py:16-186 for the
 * distortion loss of code/train.py:392,431): d_dimg2 = grad_out * d value / d img2 (N,3,H,W); d_value (optional)
 * receives the forward value.  The forward pass is recomputed inside. */
size_t ic_msssim_bwd_workspace_bytes(int N, int H, int W);
int ic_msssim_tf_bwd(const float* d_img1, const float* d_img2, int N, int H, int W, float grad_out, float* d_dimg2,
                     float* d_value, void* d_workspace, size_t workspace_bytes, void* stream);

/* replaces: ms_ssim_np.MultiScaleSSIM on uint8 (through tf_msssim_np, val.py:93)
 *           code/ms_ssim_np.py:25-110.  float64, one value PER IMAGE (d_out: N doubles). */
int ic_msssim_np_fwd(const uint8_t* d_img1, const uint8_t* d_img2, int N, int H, int W,
                     double* d_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------- training-loss forward
 * replaces the reductions of train.get_loss (code/train.py:309-311: reduce_mean(bc), reduce_mean(bc * heatmap))
 * d_out: 2 doubles = [sum bc, sum bc*heatmap] (heatmap may be NULL -> second sum 0). */
size_t ic_loss_workspace_bytes(void);
int ic_masked_sums_fwd(const float* d_bc, const float* d_heatmap, int64_t n, double* d_out,
                       void* d_workspace, size_t workspace_bytes, void* stream);
/* replaces Distortions.get_mse_per_img (code/train.py:400-418): per-image mean squared error of two float32
 * NCHW batches, optionally after the int32 cast (truncation) the reference applies outside of training. */
int ic_mse_per_image_fwd(const float* d_x, const float* d_x_out, int N, int64_t per_image, int cast_to_int,
                         float* d_out, void* stream);
/* Gradient w.r.t. x_out of the distortion term when config.distortion_to_minimize is 'mse' (psnr = 0) or 'psnr' (psnr = 1)
 * (code/train.py:381-397): d_mse = ic_mse_per_image_fwd(x, x_out, ..., cast_to_int = 0) on the device, N floats. */
int ic_nn_distortion_bwd(const float* d_x, const float* d_x_out, const float* d_mse, int N, int64_t per_image, int psnr,
                         float* d_dx_out, void* stream);

/* ------------------------------------------------------------- test hooks
 * One fused 3x3 128->128 residual conv (conv + BN + ReLU + residual adds) of the
 * encoder/decoder through the path selected by `mode`; fp32 NHWC (N,H,W,128) in/out.
 * Lets tests compare the tcgen05 kernel with the FFMA kernel layer by layer.
 * Workspace: 4 * N*H*W*128*4 bytes (+1 KB) for the tensor-core modes. */
int ic_debug_conv3x3(const ic_ae_t* ae, int decoder, int layer, const float* d_in, const float* d_res1,
                     const float* d_res2, int N, int H, int W, float* d_out,
                     void* d_workspace, size_t workspace_bytes, int mode, void* stream);

/* ------------------------------------------------------- launch accounting
 * Not part of the reference's surface: lets bench.py count this library's kernel
 * launches (`gpu_launches`) and time one kernel class live with CUDA events on the
 * launching stream (`roofline.achieved`). */
enum {
    IC_PROF_CONV3X3 = 0,     /* the 3x3 128->128 residual convs (95 % of the FLOPs) */
    IC_PROF_CONV_OTHER = 1,  /* h1, h2, to_bn, from_bn, h12, h13 */
    IC_PROF_ELEMENTWISE = 2, /* input prep, heatmap+quantizer, layout changes */
    IC_PROF_PROBCLASS = 3,
    IC_PROF_MSSSIM = 4,
    IC_PROF_NUM_CLASSES = 5
};
long long ic_launch_count(void);
void ic_profile_enable(int on);
void ic_profile_reset(void);
int ic_profile_get(int cls, double* total_ms, long long* launches);

/* ------------------------------------------------------ host arithmetic coder
 * restates: arithmetic_coding.ArithmeticEncoder / ArithmeticDecoder with
 * SimpleFrequencyTable (code/arithmetic_coding.py:39-222,323-424), 32-bit state.
 * Pure host code (the north star keeps the coder on the host). */
/* CRC-32C of a host buffer (TensorFlow tensor bundles store it, masked, per tensor: code/saver.py restores such files) */
uint32_t ic_crc32c(const void* h_data, int64_t n);

typedef struct ic_ac_enc ic_ac_enc_t;
typedef struct ic_ac_dec ic_ac_dec_t;
int ic_ac_enc_create(ic_ac_enc_t** out);
/* write n symbols; h_freqs is n x L int64 (one table per symbol, as produced by ic_pc_freqs_fwd) */
int ic_ac_enc_write(ic_ac_enc_t* e, const int64_t* h_freqs, int L, const int64_t* h_symbols, int64_t n);
/* finish(): flushes like ArithmeticEncoder.finish + BitOutputStream.close; returns the byte
 * stream (owned by the encoder until destroy) and the exact bit count before byte padding. */
/* the same over 32-bit tables and 8-bit symbols (what ic_pc_codec_freqs_u32_fwd / ic_encode_fwd's uint8 symbols give) */
int ic_ac_enc_write_u32(ic_ac_enc_t* e, const uint32_t* h_freqs, int L, const uint8_t* h_symbols, int64_t n);
int ic_ac_enc_finish(ic_ac_enc_t* e, const uint8_t** h_bytes, int64_t* n_bytes, int64_t* n_bits);
void ic_ac_enc_destroy(ic_ac_enc_t* e);
int ic_ac_dec_create(const uint8_t* h_bytes, int64_t n_bytes, ic_ac_dec_t** out);
/* decode n symbols with n tables (teacher-forced / pre-computed tables) */
int ic_ac_dec_read(ic_ac_dec_t* d, const int64_t* h_freqs, int L, int64_t* h_symbols, int64_t n);
void ic_ac_dec_destroy(ic_ac_dec_t* d);

#ifdef __cplusplus
}
#endif
#endif /* IMGCOMP_B200_H_ */
