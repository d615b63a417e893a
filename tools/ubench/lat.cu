// Single-warp latency / issue-rate probes used to budget the sequential decoder (pc_decode.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096
__global__ void k(long long* out, float fa, double da, unsigned ua) {
    const int lane = threadIdx.x;
    long long t0, t1;
    // 1. dependent FFMA chain
    float x = fa + lane;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = fmaf(x, fa, 1.0f);
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    // 2. two interleaved FFMA chains fed by SHFL (the decoder's layer loop shape)
    float a = fa, b = fa * 2, v = x;
    t0 = clock64();
#pragma unroll 24
    for (int i = 0; i < N; ++i) {
        float s = __shfl_sync(0xffffffffu, v, i & 31);
        a = fmaf(s, fa, a);
        b = fmaf(s, x, b);
    }
    t1 = clock64();
    if (lane == 0) out[1] = t1 - t0;
    // 3. dependent SHFL chain (latency)
    float w = a + b;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) w = __shfl_sync(0xffffffffu, w, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) out[2] = t1 - t0;
    // 4. independent SHFLs (issue rate)
    float acc = 0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) acc += __shfl_sync(0xffffffffu, w, i & 31);
    t1 = clock64();
    if (lane == 0) out[3] = t1 - t0;
    // 5. dependent DFMA chain
    double d = da + lane;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) d = fma(d, da, 1.0);
    t1 = clock64();
    if (lane == 0) out[4] = t1 - t0;
    // 6. u64 -> f64 -> u64 conversion chain
    unsigned long long q = ua + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) q = __double2ull_rz(__ull2double_rz(q) * da) + 1;
    t1 = clock64();
    if (lane == 0) out[5] = t1 - t0;
    // 7. f32 -> s64 conversion chain
    float g = fa * 1e9f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) g = (float)((long long)g | 1);
    t1 = clock64();
    if (lane == 0) out[6] = t1 - t0;
    // 8. expf chain
    float e = fa;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) e = expf(e) - 1.0f;
    t1 = clock64();
    if (lane == 0) out[7] = t1 - t0;
    // 9. __fdiv_rn chain
    float h = fa + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) h = __fdiv_rn(h, fa + 1.5f) + 2.0f;
    t1 = clock64();
    if (lane == 0) out[8] = t1 - t0;
    // 10. __drcp_rn chain
    double r = da + 2;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) r = __drcp_rn(r) + 2.0;
    t1 = clock64();
    if (lane == 0) out[9] = t1 - t0;
    // 11. u64 / u32-ish division chain
    unsigned long long u = 0x7fffffffffffull + lane;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) u = u / (ua | 0x40000000u) + 0x7fffffffffffull;
    t1 = clock64();
    if (lane == 0) out[10] = t1 - t0;
    // 12. shared memory broadcast round trip: STS, syncwarp, LDS
    __shared__ float sh[64];
    float m = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        sh[lane] = m;
        __syncwarp();
        m = sh[(lane + 1) & 31] + 1.0f;
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[11] = t1 - t0;
    // 13. eight independent FFMA chains (issue rate of one warp)
    float c0 = fa, c1 = fa + 1, c2 = fa + 2, c3 = fa + 3, c4 = fa + 4, c5 = fa + 5, c6 = fa + 6, c7 = fa + 7;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        c0 = fmaf(c0, fa, 1.f); c1 = fmaf(c1, fa, 1.f); c2 = fmaf(c2, fa, 1.f); c3 = fmaf(c3, fa, 1.f);
        c4 = fmaf(c4, fa, 1.f); c5 = fmaf(c5, fa, 1.f); c6 = fmaf(c6, fa, 1.f); c7 = fmaf(c7, fa, 1.f);
    }
    t1 = clock64();
    if (lane == 0) out[12] = (t1 - t0) / 8;
    // 14. independent broadcast LDS.128 (issue interval)
    __shared__ float4 sh4[64];
    sh4[lane] = make_float4(x, a, b, w);
    __syncwarp();
    float s4 = 0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        float4 v4 = sh4[i & 31];
        s4 += v4.x + v4.w;
    }
    t1 = clock64();
    if (lane == 0) out[13] = t1 - t0;
    // 15. u64 multiply chain
    unsigned long long mm = u | 1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) mm = mm * (ua | 3) + 7;
    t1 = clock64();
    if (lane == 0) out[14] = t1 - t0;
    if (x + a + b + w + acc + d + q + g + e + h + r + u + m + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7 + s4 + mm == 12345.678) out[15] = 1;
}

// two warps ping-pong through named barriers (the decode / coder hand-over)
__global__ void pingpong(long long* out) {
    const int warp = threadIdx.x >> 5;
    __shared__ volatile int box[2];
    long long t0 = clock64();
    if (warp == 0) {
        for (int i = 0; i < N; ++i) {
            box[0] = i;
            asm volatile("bar.arrive 1, 64;" ::: "memory");
            asm volatile("bar.sync 2, 64;" ::: "memory");
        }
    } else {
        for (int i = 0; i < N; ++i) {
            asm volatile("bar.sync 1, 64;" ::: "memory");
            box[1] = box[0];
            asm volatile("bar.arrive 2, 64;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
    long long* d;
    cudaMalloc(&d, 16 * sizeof(long long));
    k<<<1, 32>>>(d, 1.0001f, 1.0000001, 12345u);
    k<<<1, 32>>>(d, 1.0001f, 1.0000001, 12345u);
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    pingpong<<<1, 64>>>(d);
    long long pp = 0;
    cudaMemcpy(&pp, d, sizeof(pp), cudaMemcpyDeviceToHost);
    const char* names[] = {"dependent FFMA", "SHFL + 2 FFMA chains (per step)", "dependent SHFL", "independent SHFL (+FADD)",
                           "dependent DFMA", "u64->f64, DMUL, f64->u64, +1", "f32->s64->f32", "expf, -1", "__fdiv_rn, +2",
                           "__drcp_rn, +2", "u64 / u32 division, +c", "STS, syncwarp, LDS, syncwarp", "independent FFMA (issue interval)",
                           "independent broadcast LDS.128 (+2 FADD)", "u64 multiply-add chain"};
    for (int i = 0; i < 15; ++i) printf("%-36s %7.1f cycles/iter\n", names[i], (double)h[i] / N);
    printf("%-36s %7.1f cycles/iter\n", "named-barrier round trip, 2 warps", (double)pp / N);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
