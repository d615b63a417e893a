// Internal interface of the context-model kernels (probclass.cu).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"

namespace ic {

// Device weights, masked taps only, raster (fd,fy,fx) tap order:
//   w0 [13][K]      first mask (code/probclass.py:150-162), Cin = 1
//   w1,w2 [14][K][K] other mask (code/probclass.py:164-176)
//   w3 [14][K][L]
struct PcWeights {
    int K, L;
    const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3;
    // tensor-core path (K = 24): layers 1..3 packed for conv_tc (fp16 hi/lo, 32-channel stages)
    bool tc = false;
    const __half* wt[3] = {nullptr, nullptr, nullptr};
    const float* scale_t[3] = {nullptr, nullptr, nullptr};   // [128] weight pre-scale inverse
    const float* shift_t[3] = {nullptr, nullptr, nullptr};   // [128] bias, zero padded
    tc::GroupTable gt[3];
};

struct PcInput {
    int N, D, H, W;              // un-padded volume
    int pad_d, pad_hw;           // 4,4 for bitcost/freqs (pad_for_probclass3d), 0,0 for logits()
    const float* q;              // float source (N,D,H,W) or nullptr
    float pad_value;
    const int64_t* symbols;      // symbol source (gathered through centers) or nullptr
    float centers_host[8];
    const int64_t* target_symbols;   // for the bitcost / freqs heads (N,D,H,W)
};

enum { PC_HEAD_LOGITS = 0, PC_HEAD_BITCOST = 1, PC_HEAD_FREQS = 2 };

size_t pc_workspace_bytes(int KC, int N, int D, int H, int W, int pad_d, int pad_hw);
// canonical = true: always the float32 FFMA kernels, whose per-output fmaf chain the sequential
// decoder (pc_decode.cu) reproduces bit for bit
// out_freqs32 (canonical only): write the tables as uint32 instead of int64 (every entry is <= 1e9)
int pc_forward(const PcWeights& w, const PcInput& in, int head, float* out_f, int64_t* out_freqs,
               double* bits_sum, void* ws, size_t ws_bytes, cudaStream_t s, bool canonical = false,
               uint32_t* out_freqs32 = nullptr);

// sequential decode of N bitstreams (pc_decode.cu); all pointers are device pointers
struct PcDecodeInput {
    int N, C, h, w;
    float centers_host[8];
    const uint8_t* stream;          // concatenated bitstreams
    const int64_t* stream_off;      // [N + 1] byte offsets into `stream`
    const int32_t* first_sym;       // [N]
    uint8_t* sym_out;               // N,C,h,w
    const uint8_t* force_sym;       // optional: teacher forcing (no range decoding)
    int64_t* freqs_out;             // optional: N,C,h,w,L tables the decoder used
};
size_t pc_decode_workspace_bytes(int N, int C, int h, int w);
int pc_decode(const PcWeights& w, const PcDecodeInput& in, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace ic
