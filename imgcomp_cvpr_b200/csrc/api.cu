// C ABI of imgcomp_b200 (include/imgcomp_b200.h): handles, weight re-layout,
// layer orchestration.  No kernels here.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "probclass.cuh"

namespace ic {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

size_t msssim_workspace_bytes(int N, int H, int W, int is_double);
int msssim_tf(const float* a, const float* b, int N, int H, int W, float* out, float* levels, void* ws,
              size_t ws_bytes, cudaStream_t s);
int msssim_np(const uint8_t* a, const uint8_t* b, int N, int H, int W, double* out, void* ws, size_t ws_bytes,
              cudaStream_t s);

namespace {

constexpr int kArchN = 128;       // code/autoencoder.py:210
constexpr float kBnEps = 1e-5f;   // code/autoencoder.py:118

struct LayerSpec {
    std::string scope;
    int k, stride, cin, cout;
    bool transposed, relu;
};

// execution-order list of the autoencoder convs (code/autoencoder.py:218-268)
std::vector<LayerSpec> ae_layers(const ic_ae_config& c, bool decoder) {
    std::vector<LayerSpec> v;
    const int n = kArchN, C = c.num_chan_bn, CB = c.heatmap ? C + 1 : C;
    const std::string P = decoder ? "autoencoder/decoder" : "autoencoder/encoder";
    const char* tag = decoder ? "dec" : "enc";
    if (!decoder) {
        v.push_back({P + "/h1", 5, 2, 3, n / 2, false, true});
        v.push_back({P + "/h2", 5, 2, n / 2, n, false, true});
    } else {
        v.push_back({P + "/from_bn", 3, 2, C, n, true, true});
    }
    char buf[128];
    for (int b = 0; b < c.arch_param_B; ++b)
        for (int i = 1; i <= 3; ++i)
            for (int j = 1; j <= 2; ++j) {
                snprintf(buf, sizeof(buf), "/res_block_%s_%d/%s_%d_%d/conv%d", tag, b, tag, b, i, j);
                v.push_back({P + buf, 3, 1, n, n, false, j == 1});
            }
    const char* fin = decoder ? "/dec_after_res" : "/res_block_enc_final";
    v.push_back({P + fin + "/conv1", 3, 1, n, n, false, false});   // activation_fn=None for both (autoencoder.py:232-233)
    v.push_back({P + fin + "/conv2", 3, 1, n, n, false, false});
    if (!decoder) {
        v.push_back({P + "/to_bn", 5, 2, n, CB, false, false});
    } else {
        v.push_back({P + "/h12", 5, 2, n, n / 2, true, true});
        v.push_back({P + "/h13", 5, 2, n / 2, 3, true, false});
    }
    return v;
}

const char* kBnNames[4] = {"gamma", "beta", "moving_mean", "moving_variance"};

struct TensorInfo {
    std::string name;
    int64_t numel;
};

std::vector<TensorInfo> ae_tensors(const ic_ae_config& c) {
    std::vector<TensorInfo> t;
    for (int dec = 0; dec < 2; ++dec) {
        for (const auto& l : ae_layers(c, dec)) {
            t.push_back({l.scope + "/weights", (int64_t)l.k * l.k * l.cin * l.cout});
            for (int i = 0; i < 4; ++i) t.push_back({l.scope + "/BatchNorm/" + kBnNames[i], l.cout});
        }
        if (!dec) t.push_back({"autoencoder/encoder/centers", c.num_centers});
    }
    return t;
}

struct DevLayer {
    LayerSpec spec;
    int cin_pad, ldw;
    float *w = nullptr, *scale = nullptr, *shift = nullptr;
    // tensor-core path (3x3 128->128 layers only): packed fp16 hi/lo weights and the BN scale
    // divided by the power-of-two weight pre-scale
    __half* w_tc = nullptr;
    __half* w_tc_pair = nullptr;   // pair-packed copy for the cta_group::2 kernel (IC_CONV_PAIR=1)
    __half* w_h1 = nullptr;        // h1 only: im2col-packed weights of the dedicated kernel (conv_h1.cu)
    float* scale_h1 = nullptr;
    __half* w_tc_cat = nullptr;    // B-concatenated copy for conv_cat_kernel (EXACT mode; IC_CONV_CAT=0 disables, =2: CTA pairs)
    __half* w_tc_cat_pair = nullptr;
    float* scale_tc = nullptr;
    tc::GroupTable gt;
    int nout_tc = 0;
    // transposed convs (decoder): depth-to-space packing; from_bn needs two launches (output rows 2m and 2m+1)
    __half* w_tc_b = nullptr;
    tc::GroupTable gt_b;
};

int upload(const std::vector<float>& h, float** d) {
    IC_CHECK_CUDA(cudaMalloc((void**)d, h.size() * sizeof(float)));
    IC_CHECK_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return IC_OK;
}

bool cfg_ok(const ic_ae_config* c) {
    return c && (c->num_chan_bn % 4 == 0) && c->num_chan_bn > 0 && c->num_chan_bn <= 256 && c->arch_param_B >= 0 &&
           c->arch_param_B <= 16 && c->num_centers >= 1 && c->num_centers <= 8;
}

}  // namespace
}  // namespace ic

using namespace ic;

struct ic_ae {
    ic_ae_config cfg;
    std::vector<DevLayer> enc, dec;
    float* d_centers = nullptr;
    float h_centers[8];
};

struct ic_pc {
    ic_pc_config cfg;
    PcWeights w;
    std::vector<float*> owned;
    std::vector<__half*> owned_h;
};

extern "C" {

const char* ic_last_error(void) { return g_err; }
int ic_abi_version(void) { return IC_ABI_VERSION; }

int ic_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

// ------------------------------------------------------------------ autoencoder
int ic_ae_num_tensors(const ic_ae_config* cfg) { return cfg_ok(cfg) ? (int)ae_tensors(*cfg).size() : IC_ERR_INVALID; }

const char* ic_ae_tensor_name(const ic_ae_config* cfg, int i) {
    static thread_local std::string s;
    if (!cfg_ok(cfg)) return nullptr;
    auto t = ae_tensors(*cfg);
    if (i < 0 || i >= (int)t.size()) return nullptr;
    s = t[i].name;
    return s.c_str();
}

int64_t ic_ae_tensor_numel(const ic_ae_config* cfg, int i) {
    if (!cfg_ok(cfg)) return IC_ERR_INVALID;
    auto t = ae_tensors(*cfg);
    if (i < 0 || i >= (int)t.size()) return IC_ERR_INVALID;
    return t[i].numel;
}

static int ae_build(ic_ae* ae, const ic_ae_config* cfg, const float* const* h_tensors, int n_tensors) {
    IC_REQUIRE(cfg_ok(cfg) && h_tensors, IC_ERR_INVALID, "ic_ae_create: bad config / NULL argument");
    IC_REQUIRE(cfg->normalization == 0 || cfg->normalization == 1, IC_ERR_UNSUPPORTED, "normalization must be OFF(0) or FIXED(1)");
    auto tensors = ae_tensors(*cfg);
    IC_REQUIRE(n_tensors == (int)tensors.size(), IC_ERR_INVALID, "ic_ae_create: expected %zu tensors, got %d",
               tensors.size(), n_tensors);
    for (int i = 0; i < n_tensors; ++i) IC_REQUIRE(h_tensors[i], IC_ERR_INVALID, "ic_ae_create: tensor %d (%s) is NULL", i, tensors[i].name.c_str());
    ae->cfg = *cfg;
    int ti = 0;
    for (int dec = 0; dec < 2; ++dec) {
        auto& dst = dec ? ae->dec : ae->enc;
        for (const auto& l : ae_layers(*cfg, dec)) {
            dst.emplace_back();            // owned by the handle from here on: early returns cannot leak
            DevLayer& d = dst.back();
            d.spec = l;
            d.cin_pad = (int)align_up(l.cin, 4);
            d.ldw = (int)align_up(l.cout, 4);
            const float* w = h_tensors[ti];
            const float *g = h_tensors[ti + 1], *be = h_tensors[ti + 2], *mu = h_tensors[ti + 3], *var = h_tensors[ti + 4];
            ti += 5;
            // GEMM B matrix [(ky,kx,ci)][co]; conv2d weights are HWIO, conv2d_transpose weights [kh,kw,Cout,Cin]
            std::vector<float> wm((size_t)l.k * l.k * d.cin_pad * d.ldw, 0.f), sc(d.ldw, 0.f), sh(d.ldw, 0.f);
            for (int t = 0; t < l.k * l.k; ++t)
                for (int ci = 0; ci < l.cin; ++ci)
                    for (int co = 0; co < l.cout; ++co) {
                        float v = l.transposed ? w[((size_t)t * l.cout + co) * l.cin + ci] : w[((size_t)t * l.cin + ci) * l.cout + co];
                        wm[((size_t)t * d.cin_pad + ci) * d.ldw + co] = v;
                    }
            for (int co = 0; co < l.cout; ++co) {   // fused batch norm, inference
                float inv = 1.0f / sqrtf(var[co] + kBnEps);
                sc[co] = g[co] * inv;
                sh[co] = be[co] - mu[co] * sc[co];
            }
            int rc = upload(wm, &d.w);
            if (rc == IC_OK) rc = upload(sc, &d.scale);
            if (rc == IC_OK) rc = upload(sh, &d.shift);
            // tensor-core path: the 3x3 128->128 residual convs, h2 (5x5 s2 64->128) and to_bn (5x5 s2 128->C+1, padded to 48 or 80)
            const bool tc_res = l.k == 3 && l.stride == 1 && !l.transposed && l.cin == 128 && l.cout == 128;
            const bool tc_s2 = l.k == 5 && l.stride == 2 && !l.transposed && l.cin % 64 == 0 && (l.cout == 128 || l.cout <= 80);
            const bool tc_t = l.transposed && l.stride == 2 && l.cin % 32 == 0 &&
                              ((l.k == 3 && l.cout == 128) || (l.k == 5 && (l.cout == 64 || l.cout == 3)));
            if (rc == IC_OK && tc_t) {
                std::vector<float> sct(128, 0.f), sht(128, 0.f);
                float inv = 1.f;
                const int all[4] = {0, 1, 2, 3};
                std::vector<__half> packed;
                d.nout_tc = l.cout == 3 ? 16 : 256;
                if (l.cout == 128) {       // from_bn: 4 phases x 128 = 512 columns -> two launches of 256
                    rc = tc::pack_weights_tconv(w, l.k, l.cin, l.cout, all, 2, 256, packed, d.gt, &inv);
                    if (rc == IC_OK) {
                        std::vector<__half> pb;
                        float inv2;
                        rc = tc::pack_weights_tconv(w, l.k, l.cin, l.cout, all + 2, 2, 256, pb, d.gt_b, &inv2);
                        if (rc == IC_OK) {
                            IC_CHECK_CUDA(cudaMalloc((void**)&d.w_tc_b, pb.size() * sizeof(__half)));
                            IC_CHECK_CUDA(cudaMemcpy(d.w_tc_b, pb.data(), pb.size() * sizeof(__half), cudaMemcpyHostToDevice));
                        }
                    }
                } else {
                    rc = tc::pack_weights_tconv(w, l.k, l.cin, l.cout, all, 4, d.nout_tc, packed, d.gt, &inv);
                }
                if (rc == IC_OK) {
                    for (int co = 0; co < l.cout; ++co) {
                        sct[co] = sc[co] * inv;
                        sht[co] = sh[co];
                    }
                    IC_CHECK_CUDA(cudaMalloc((void**)&d.w_tc, packed.size() * sizeof(__half)));
                    IC_CHECK_CUDA(cudaMemcpy(d.w_tc, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
                    rc = upload(sct, &d.scale_tc);
                    if (rc == IC_OK) {
                        cudaFree(d.shift);
                        rc = upload(sht, &d.shift);
                    }
                } else {
                    d.nout_tc = 0;
                    rc = IC_OK;
                }
            }
            const bool tc_h1 = l.k == 5 && l.stride == 2 && !l.transposed && l.cin == 3 && l.cout == 64;
            if (rc == IC_OK && tc_h1) {
                std::vector<__half> ph1;
                float inv1 = 1.f;
                if (tc::pack_weights_h1_im2col(w, l.cin, l.cout, ph1, &inv1) == IC_OK) {
                    std::vector<float> s1(64, 0.f);
                    for (int co = 0; co < l.cout; ++co) s1[co] = sc[co] * inv1;         // exact: inv1 is a power of two
                    IC_CHECK_CUDA(cudaMalloc((void**)&d.w_h1, ph1.size() * sizeof(__half)));
                    IC_CHECK_CUDA(cudaMemcpy(d.w_h1, ph1.data(), ph1.size() * sizeof(__half), cudaMemcpyHostToDevice));
                    rc = upload(s1, &d.scale_h1);
                }
            }
            if (rc == IC_OK && (tc_res || tc_s2 || tc_h1)) {
                std::vector<__half> packed;
                float inv = 1.f;
                d.nout_tc = tc_h1 ? 64 : (l.cout == 128 ? 128 : (l.cout <= 48 ? 48 : 80));
                rc = tc_h1 ? tc::pack_weights_h1(w, l.cin, l.cout, d.nout_tc, packed, d.gt, &inv)
                           : tc::pack_weights(w, l.k, l.stride, l.cin, l.cout, d.nout_tc, packed, d.gt, &inv);
                if (rc == IC_OK) {
                    std::vector<float> sct(128, 0.f), sht(128, 0.f);
                    for (int co = 0; co < l.cout; ++co) {
                        sct[co] = sc[co] * inv;               // exact: inv is a power of two
                        sht[co] = sh[co];
                    }
                    IC_CHECK_CUDA(cudaMalloc((void**)&d.w_tc, packed.size() * sizeof(__half)));
                    IC_CHECK_CUDA(cudaMemcpy(d.w_tc, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
                    // 128-column convs run as CTA pairs (cta_group::2: M = 256 over two CTAs, each CTA holds half of B and
                    // fetches 6 KB instead of 8 KB of operands per MMA); IC_CONV_PAIR=0 selects the single-CTA kernel
                    const char* pair_env = getenv("IC_CONV_PAIR");
                    if (d.nout_tc == 128 && !(pair_env && atoi(pair_env) == 0)) {
                        std::vector<__half> pp;
                        tc::repack_pair(packed, d.gt.nstages, pp);
                        IC_CHECK_CUDA(cudaMalloc((void**)&d.w_tc_pair, pp.size() * sizeof(__half)));
                        IC_CHECK_CUDA(cudaMemcpy(d.w_tc_pair, pp.data(), pp.size() * sizeof(__half), cudaMemcpyHostToDevice));
                    }
                    // opt-in (IC_CONV_CAT=1: one CTA, =2: CTA pairs): measured slower than conv_tc_kernel on B200 because the
                    // single-buffered accumulators expose the epilogue (profiles/r2_summary.md)
                    const int cat_mode = getenv("IC_CONV_CAT") ? atoi(getenv("IC_CONV_CAT")) : 0;
                    if (d.nout_tc == 128 && cat_mode) {
                        std::vector<__half> pc_;
                        if (cat_mode == 2) tc::repack_cat_pair(packed, d.gt.nstages, pc_);
                        else tc::repack_cat(packed, d.gt.nstages, pc_);
                        __half** dst = cat_mode == 2 ? &d.w_tc_cat_pair : &d.w_tc_cat;
                        IC_CHECK_CUDA(cudaMalloc((void**)dst, pc_.size() * sizeof(__half)));
                        IC_CHECK_CUDA(cudaMemcpy(*dst, pc_.data(), pc_.size() * sizeof(__half), cudaMemcpyHostToDevice));
                    }
                    rc = upload(sct, &d.scale_tc);
                    if (rc == IC_OK) {       // padded shift for the tensor-core epilogue
                        cudaFree(d.shift);
                        std::vector<float> shp(std::max<size_t>(128, sh.size()), 0.f);
                        for (size_t i = 0; i < sh.size(); ++i) shp[i] = sh[i];
                        rc = upload(shp, &d.shift);
                    }
                } else {
                    d.nout_tc = 0;
                    rc = IC_OK;
                }
            }
            if (rc != IC_OK) return rc;
        }
        if (!dec) {
            std::vector<float> c(h_tensors[ti], h_tensors[ti] + cfg->num_centers);
            memset(ae->h_centers, 0, sizeof(ae->h_centers));
            memcpy(ae->h_centers, c.data(), sizeof(float) * cfg->num_centers);
            ++ti;
            int rc = upload(c, &ae->d_centers);
            if (rc != IC_OK) return rc;
        }
    }
    return IC_OK;
}

void ic_ae_destroy(ic_ae_t* ae);

int ic_ae_create(const ic_ae_config* cfg, const float* const* h_tensors, int n_tensors, ic_ae_t** out) {
    IC_REQUIRE(out, IC_ERR_INVALID, "ic_ae_create: NULL argument");
    ic_ae* ae = new ic_ae();
    int rc = ae_build(ae, cfg, h_tensors, n_tensors);
    if (rc != IC_OK) {
        ic_ae_destroy(ae);
        return rc;
    }
    *out = ae;
    return IC_OK;
}

void ic_ae_destroy(ic_ae_t* ae) {
    if (!ae) return;
    for (auto* v : {&ae->enc, &ae->dec})
        for (auto& l : *v) {
            cudaFree(l.w);
            cudaFree(l.scale);
            cudaFree(l.shift);
            cudaFree(l.w_tc);
            cudaFree(l.w_tc_pair);
            cudaFree(l.w_tc_cat);
            cudaFree(l.w_h1);
            cudaFree(l.scale_h1);
            cudaFree(l.w_tc_cat_pair);
            cudaFree(l.w_tc_b);
            cudaFree(l.scale_tc);
        }
    cudaFree(ae->d_centers);
    delete ae;
}

int ic_ae_centers(const ic_ae_t* ae, float* d_centers, void* stream) {
    IC_REQUIRE(ae && d_centers, IC_ERR_INVALID, "ic_ae_centers: NULL argument");
    IC_CHECK_CUDA(cudaMemcpyAsync(d_centers, ae->d_centers, sizeof(float) * ae->cfg.num_centers, cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
    return IC_OK;
}

}  // extern "C"

namespace {

ConvDesc make_desc(const DevLayer& L, const float* in, int N, int Hi, int Wi, float* out) {
    ConvDesc d;
    memset(&d, 0, sizeof(d));
    const LayerSpec& s = L.spec;
    d.in = in;
    d.w = L.w;
    d.scale = L.scale;
    d.shift = L.shift;
    d.out = out;
    d.N = N;
    d.Hi = Hi;
    d.Wi = Wi;
    d.Cin = L.cin_pad;
    d.Cout = s.cout;
    d.ldw = L.ldw;
    d.KH = d.KW = s.k;
    d.stride = s.stride;
    d.transposed = s.transposed;
    d.relu = s.relu;
    if (!s.transposed) {
        d.Ho = (Hi + s.stride - 1) / s.stride;
        d.Wo = (Wi + s.stride - 1) / s.stride;
        d.pad_t = same_pad_before(Hi, s.k, s.stride);
        d.pad_l = same_pad_before(Wi, s.k, s.stride);
    } else {   // gradient of the SAME conv from stride*n -> n
        d.Ho = Hi * s.stride;
        d.Wo = Wi * s.stride;
        d.pad_t = same_pad_before(d.Ho, s.k, s.stride);
        d.pad_l = same_pad_before(d.Wo, s.k, s.stride);
    }
    return d;
}

// 15 residual blocks (+5 group skips) + final no-ReLU block + long skip
// (code/autoencoder.py:224-234 / :252-262).  `layers` points at the first 3x3 conv.
// pool: 5 trunk-sized buffers, pool[0] holds the input on entry.  `conv(layer, in, out, res1, res2)`
// launches one fused conv.  Returns the index of the output buffer.
template <typename ConvFn>
int run_res_stack(const DevLayer* layers, int B, ConvFn conv, int* result) {
    bool busy[5] = {true, false, false, false, false};
    auto grab = [&]() {
        for (int i = 0; i < 5; ++i)
            if (!busy[i]) {
                busy[i] = true;
                return i;
            }
        return -1;
    };
    const int r0 = 0;
    int x = 0, li = 0;
    for (int b = 0; b <= B; ++b) {
        const bool final_block = (b == B);
        const int rb = x;
        for (int i = 0; i < (final_block ? 1 : 3); ++i) {
            int t1 = grab();
            int rc = conv(layers[li++], x, t1, -1, -1);
            if (rc != IC_OK) return rc;
            int y = grab();
            const bool last = final_block || i == 2;
            // residual_block: x + residual_input ; then net = net + residual_input_{b,0}
            rc = conv(layers[li++], t1, y, x, last ? (final_block ? r0 : rb) : -1);
            if (rc != IC_OK) return rc;
            busy[t1] = false;
            if (x != rb && x != r0) busy[x] = false;
            x = y;
        }
        if (rb != r0 && rb != x) busy[rb] = false;
    }
    *result = x;
    return IC_OK;
}

int run_res_stack_simt(const DevLayer* layers, int B, int N, int H, int W, float* pool[5], float** result, cudaStream_t s) {
    int idx = 0;
    int rc = run_res_stack(layers, B, [&](const DevLayer& L, int in, int out, int r1, int r2) {
        ConvDesc d = make_desc(L, pool[in], N, H, W, pool[out]);
        d.res1 = r1 >= 0 ? pool[r1] : nullptr;
        d.res2 = r2 >= 0 ? pool[r2] : nullptr;
        return launch_conv_simt(d, s);
    }, &idx);
    *result = pool[idx];
    return rc;
}

int conv_tc_layer(const DevLayer& L, const __half* in, int in_chunks, int Hin, int Win, __half* out, float* out_f32,
                  const __half* r1, const __half* r2, int N, int H, int W, bool exact, cudaStream_t s, int out_s2d = 0) {
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in;
    a.Nimg = N;
    a.in_chunks = in_chunks;
    a.Hin = Hin;
    a.Win = Win;
    a.weights = L.w_tc;
    a.weights_pair = L.w_tc_pair;
    a.weights_cat = L.w_tc_cat;
    a.weights_cat_pair = L.w_tc_cat_pair;
    a.groups = &L.gt;
    a.scale = L.scale_tc;
    a.shift = L.shift;
    a.res1 = r1;
    a.res2 = r2;
    a.out = out;
    a.out_f32 = out_f32;
    a.N = N;
    a.H = H;
    a.W = W;
    a.relu = L.spec.relu;
    a.cout = L.spec.cout;
    a.nout = L.nout_tc;
    a.halo0 = -1;
    a.img_mul = 1;
    a.head = -1;
    a.out_s2d = out_s2d;
    a.cpg = 4;
    a.exact = exact;
    a.prof_class = L.spec.k == 3 ? IC_PROF_CONV3X3 : IC_PROF_CONV_OTHER;
    return tc::launch_conv_tc(a, s);
}

// tensor-core res stack over fp16 hi/lo plane buffers; pool[0] holds the input.  Returns the output index.
// one depth-to-space launch of a stride-2 transposed conv on tensor cores (see tc::pack_weights_tconv)
int tconv_tc_layer(const DevLayer& L, bool second, const __half* in, int in_chunks, int N, int Hin, int Win, __half* out,
                   float* out_img, uint8_t* out_u8, int denorm, bool exact, cudaStream_t s) {
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in;
    a.Nimg = N;
    a.in_chunks = in_chunks;
    a.Hin = Hin;
    a.Win = Win;
    a.weights = second ? L.w_tc_b : L.w_tc;
    a.groups = second ? &L.gt_b : &L.gt;
    a.scale = L.scale_tc;
    a.shift = L.shift;
    a.out = out;
    a.out_f32 = out_img;
    a.out_u8 = out_u8;
    a.denorm = denorm;
    a.N = N;
    a.H = Hin;
    a.W = Win;
    a.relu = L.spec.relu;
    a.cout = L.spec.cout;
    a.nout = L.nout_tc;
    a.halo0 = -1;
    a.img_mul = 1;
    a.head = -1;
    a.cpg = 4;
    a.exact = exact;
    a.prof_class = IC_PROF_CONV_OTHER;
    if (out) {
        a.d2s_cch = L.spec.cout / 8;
        a.d2s_ph0 = second ? 2 : 0;
    }
    return tc::launch_conv_tc(a, s);
}

// last_s2d: the final conv writes its output in space-to-depth form (the input layout of the stride-2 to_bn)
int run_res_stack_tc(const DevLayer* layers, int B, int N, int H, int W, __half* pool[5], bool exact, int* result,
                     cudaStream_t s, bool last_s2d = false) {
    const DevLayer* last = layers + (6 * B + 2 - 1);
    return run_res_stack(layers, B, [&](const DevLayer& L, int in, int out, int r1, int r2) {
        return conv_tc_layer(L, pool[in], 16, H, W, pool[out], nullptr, r1 >= 0 ? pool[r1] : nullptr,
                             r2 >= 0 ? pool[r2] : nullptr, N, H, W, exact, s, (last_s2d && &L == last) ? 1 : 0);
    }, result);
}

}  // namespace

extern "C" {

size_t ic_encode_workspace_bytes(const ic_ae_t* ae, int N, int H, int W, int mode) {
    if (!ae || N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t n = N;
    const int CB = ae->cfg.heatmap ? ae->cfg.num_chan_bn + 1 : ae->cfg.num_chan_bn;
    size_t b = 0;
    b += align_up(n * H * W * 4 * 4, 256);
    b += align_up(n * (H / 2) * (W / 2) * 64 * 4, 256);
    (void)mode;
    b += 5 * align_up(n * (H / 4) * (W / 4) * 128 * 4, 256);
    b += align_up(n * (H / 8) * (W / 8) * CB * 4, 256);
    return b + 4096;
}

int ic_encode_fwd(const ic_ae_t* ae, const void* d_x, int x_is_u8, int N, int H, int W, float* d_z, float* d_heatmap,
                  float* d_qbar, float* d_qhard, int64_t* d_symbols, uint8_t* d_symbols_u8, float* d_qsoft,
                  void* d_workspace, size_t workspace_bytes, int mode, void* stream) {
    IC_REQUIRE(ae && d_x && d_workspace, IC_ERR_INVALID, "ic_encode_fwd: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, IC_ERR_INVALID,
               "ic_encode_fwd: N=%d H=%d W=%d; H and W must be positive multiples of the subsampling factor 8", N, H, W);
    IC_REQUIRE(mode == IC_MODE_FP32 || mode == IC_MODE_EXACT || mode == IC_MODE_FAST, IC_ERR_INVALID, "ic_encode_fwd: bad mode %d", mode);
    cudaStream_t s = (cudaStream_t)stream;
    const ic_ae_config& c = ae->cfg;
    const int CB = c.heatmap ? c.num_chan_bn + 1 : c.num_chan_bn;
    Arena ar(d_workspace, workspace_bytes);
    const size_t n = N;
    float* xin = ar.get<float>(n * H * W * 4);
    float* a1 = ar.get<float>(n * (H / 2) * (W / 2) * 64);
    float* pool[5];
    for (int i = 0; i < 5; ++i) pool[i] = ar.get<float>(n * (H / 4) * (W / 4) * 128);
    float* bn = ar.get<float>(n * (H / 8) * (W / 8) * CB);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_encode_fwd: workspace too small: need %zu, have %zu", ar.off, workspace_bytes);

    const DevLayer* L = ae->enc.data();
    const DevLayer& tobn = ae->enc.back();
    int rc;
    if (mode == IC_MODE_FP32) {
        rc = launch_prep_input(d_x, x_is_u8, N, H, W, c.normalization, xin, s);
        if (rc != IC_OK) return rc;
        rc = launch_conv_simt(make_desc(L[0], xin, N, H, W, a1), s);
        if (rc != IC_OK) return rc;
        float* trunk = nullptr;
        rc = launch_conv_simt(make_desc(L[1], a1, N, H / 2, W / 2, pool[0]), s);
        if (rc != IC_OK) return rc;
        rc = run_res_stack_simt(L + 2, c.arch_param_B, N, H / 4, W / 4, pool, &trunk, s);
        if (rc != IC_OK) return rc;
        rc = launch_conv_simt(make_desc(tobn, trunk, N, H / 4, W / 4, bn), s);
        if (rc != IC_OK) return rc;
    } else {
        const bool exact = mode == IC_MODE_EXACT;
        const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
        __half* hp[5];
        for (int i = 0; i < 5; ++i) hp[i] = reinterpret_cast<__half*>(pool[i]);
        // h1: IC_H1_GENERIC=0 selects the dedicated kernel (conv_h1.cu).  Default is the prep pass + grouped-tap kernel: the
        // dedicated kernel's on-chip im2col builder still costs what the saved HBM pass does (16.4 ms vs
        // 16.6 ms per step)
        const char* h1_env = getenv("IC_H1_GENERIC");
        const bool h1_generic = !h1_env || atoi(h1_env);
        if (L[0].w_h1 && !h1_generic) {
            // h1 from the image itself: normalisation, im2col and the conv in one kernel (conv_h1.cu), output (64 ch at
            // H/2) written space-to-depth [pl][N][32][H/4][W/4][8] into pool[1..2]
            rc = tc::launch_conv_h1(d_x, x_is_u8, N, H, W, c.normalization, L[0].w_h1, L[0].scale_h1, L[0].shift, L[0].spec.relu,
                                    hp[1], exact, s);
            if (rc != IC_OK) return rc;
        } else {
            // image -> normalised, space-to-depth hi/lo planes [pl][N][4][H/2][W/2][8]  (fits in `a1`), then h1 as 9
            // stride-1 taps over one 32-channel group on the generic kernel
            __half* x2 = reinterpret_cast<__half*>(a1);
            rc = launch_prep_input_s2d(d_x, x_is_u8, N, H, W, c.normalization, x2, exact, s);
            if (rc != IC_OK) return rc;
            rc = conv_tc_layer(L[0], x2, 4, H2, W2, hp[1], nullptr, nullptr, nullptr, N, H2, W2, exact, s, 1);
            if (rc != IC_OK) return rc;
        }
        rc = conv_tc_layer(L[1], hp[1], 32, H4, W4, hp[0], nullptr, nullptr, nullptr, N, H4, W4, exact, s);   // h2
        if (rc != IC_OK) return rc;
        int ti = 0;
        rc = run_res_stack_tc(L + 2, c.arch_param_B, N, H4, W4, hp, exact, &ti, s, tobn.nout_tc != 0);
        if (rc != IC_OK) return rc;
        const int f1 = (ti + 1) % 5;
        if (tobn.nout_tc) {
            // the last residual conv already wrote [pl][N][64][H/8][W/8][8]: to_bn on tensor cores -> fp32 NHWC
            rc = conv_tc_layer(tobn, hp[ti], 64, H / 8, W / 8, nullptr, bn, nullptr, nullptr, N, H / 8, W / 8, exact, s);
        } else {
            rc = tc::launch_merge_to_nhwc(hp[ti], N, H4, W4, 128, pool[f1], exact, s);
            if (rc != IC_OK) return rc;
            rc = launch_conv_simt(make_desc(tobn, pool[f1], N, H4, W4, bn), s);
        }
        if (rc != IC_OK) return rc;
    }
    return launch_heatmap_quantize(bn, N, H / 8, W / 8, c.num_chan_bn, c.heatmap, ae->d_centers, c.num_centers, d_z,
                                   d_heatmap, d_qbar, d_qhard, d_symbols, d_symbols_u8, d_qsoft, s);
}

size_t ic_decode_workspace_bytes(const ic_ae_t* ae, int N, int h, int w, int mode) {
    if (!ae || N <= 0 || h <= 0 || w <= 0) return 0;
    const size_t n = N;
    size_t b = 0;
    b += align_up(n * h * w * ae->cfg.num_chan_bn * 4, 256);
    b += (mode == IC_MODE_FP32 ? 5 : 7) * align_up(n * (2 * h) * (2 * w) * 128 * 4, 256);
    b += align_up(n * (4 * h) * (4 * w) * 64 * 4, 256);
    return b + 4096;
}

int ic_decode_fwd(const ic_ae_t* ae, const float* d_q, int N, int h, int w, float* d_x_out, uint8_t* d_x_out_u8,
                  void* d_workspace, size_t workspace_bytes, int mode, void* stream) {
    IC_REQUIRE(ae && d_q && d_x_out && d_workspace, IC_ERR_INVALID, "ic_decode_fwd: NULL argument");
    IC_REQUIRE(N > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_decode_fwd: bad shape N=%d h=%d w=%d", N, h, w);
    IC_REQUIRE(mode == IC_MODE_FP32 || mode == IC_MODE_EXACT || mode == IC_MODE_FAST, IC_ERR_INVALID, "ic_decode_fwd: bad mode %d", mode);
    cudaStream_t s = (cudaStream_t)stream;
    const ic_ae_config& c = ae->cfg;
    Arena ar(d_workspace, workspace_bytes);
    const size_t n = N;
    float* qn = ar.get<float>(n * h * w * c.num_chan_bn);
    float* pool[7];
    for (int i = 0; i < (mode == IC_MODE_FP32 ? 5 : 7); ++i) pool[i] = ar.get<float>(n * (2 * h) * (2 * w) * 128);
    float* a12 = ar.get<float>(n * (4 * h) * (4 * w) * 64);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_decode_fwd: workspace too small: need %zu, have %zu", ar.off, workspace_bytes);
    int rc = IC_OK;
    const DevLayer* L = ae->dec.data();
    float* trunk = nullptr;
    if (mode == IC_MODE_FP32) {
        rc = launch_nchw_to_nhwc(d_q, N, c.num_chan_bn, h, w, qn, s);
        if (rc != IC_OK) return rc;
        rc = launch_conv_simt(make_desc(L[0], qn, N, h, w, pool[0]), s);
        if (rc != IC_OK) return rc;
        rc = run_res_stack_simt(L + 1, c.arch_param_B, N, 2 * h, 2 * w, pool, &trunk, s);
    } else {
        const bool exact = mode == IC_MODE_EXACT;
        const size_t nl = ae->dec.size();
        __half* hp[5];
        for (int i = 0; i < 5; ++i) hp[i] = reinterpret_cast<__half*>(pool[i]);
        const bool all_tc = L[0].nout_tc && L[nl - 2].nout_tc && L[nl - 1].nout_tc;
        if (all_tc) {
            // q (NCHW fp32) -> hi/lo planes; from_bn / h12 / h13 as depth-to-space convs on tensor cores
            __half* qp = reinterpret_cast<__half*>(qn);
            rc = tc::launch_split_from_nchw(d_q, N, c.num_chan_bn, h, w, qp, exact, s);
            if (rc != IC_OK) return rc;
            for (int half = 0; half < 2 && rc == IC_OK; ++half)
                rc = tconv_tc_layer(L[0], half == 1, qp, c.num_chan_bn / 8, N, h, w, hp[0], nullptr, nullptr, 0, exact, s);
            if (rc != IC_OK) return rc;
            int ti = 0;
            rc = run_res_stack_tc(L + 1, c.arch_param_B, N, 2 * h, 2 * w, hp, exact, &ti, s);
            if (rc != IC_OK) return rc;
            const int a0 = ti >= 2 ? 0 : 3;           // two adjacent free trunk buffers hold the 64-ch 4h x 4w tensor
            rc = tconv_tc_layer(L[nl - 2], false, hp[ti], 16, N, 2 * h, 2 * w, hp[a0], nullptr, nullptr, 0, exact, s);
            if (rc != IC_OK) return rc;
            return tconv_tc_layer(L[nl - 1], false, hp[a0], 8, N, 4 * h, 4 * w, nullptr, d_x_out, d_x_out_u8, c.normalization,
                                  exact, s);
        }
        rc = launch_nchw_to_nhwc(d_q, N, c.num_chan_bn, h, w, qn, s);
        if (rc != IC_OK) return rc;
        rc = launch_conv_simt(make_desc(L[0], qn, N, h, w, pool[5]), s);
        if (rc != IC_OK) return rc;
        rc = tc::launch_split_from_nhwc(pool[5], N, 2 * h, 2 * w, 128, 0, hp[0], exact, s);
        if (rc != IC_OK) return rc;
        int ti = 0;
        rc = run_res_stack_tc(L + 1, c.arch_param_B, N, 2 * h, 2 * w, hp, exact, &ti, s);
        if (rc != IC_OK) return rc;
        trunk = pool[6];
        rc = tc::launch_merge_to_nhwc(hp[ti], N, 2 * h, 2 * w, 128, trunk, exact, s);
    }
    if (rc != IC_OK) return rc;
    const size_t nl = ae->dec.size();
    rc = launch_conv_simt(make_desc(L[nl - 2], trunk, N, 2 * h, 2 * w, a12), s);
    if (rc != IC_OK) return rc;
    ConvDesc d = make_desc(L[nl - 1], a12, N, 4 * h, 4 * w, d_x_out);
    d.out_nchw = 1;
    d.denorm = c.normalization;
    d.out_u8 = d_x_out_u8;
    return launch_conv_simt(d, s);
}

// test hook: one fused 3x3 128->128 conv of the encoder (layer index into the 32 residual convs),
// fp32 NHWC in / out, through the chosen path.  Workspace: 4 trunk-sized buffers.
int ic_debug_conv3x3(const ic_ae_t* ae, int decoder, int layer, const float* d_in, const float* d_res1, const float* d_res2,
                     int N, int H, int W, float* d_out, void* d_workspace, size_t workspace_bytes, int mode, void* stream) {
    IC_REQUIRE(ae && d_in && d_out, IC_ERR_INVALID, "ic_debug_conv3x3: NULL argument");
    const auto& layers = decoder ? ae->dec : ae->enc;
    const int first = decoder ? 1 : 2;
    IC_REQUIRE(layer >= 0 && first + layer < (int)layers.size() && layers[first + layer].w_tc && layers[first + layer].spec.k == 3, IC_ERR_INVALID,
               "ic_debug_conv3x3: layer %d is not a 3x3 128->128 conv", layer);
    const DevLayer& L = layers[first + layer];
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == IC_MODE_FP32) {
        ConvDesc d = make_desc(L, d_in, N, H, W, d_out);
        d.res1 = d_res1;
        d.res2 = d_res2;
        return launch_conv_simt(d, s);
    }
    const bool exact = mode == IC_MODE_EXACT;
    Arena ar(d_workspace, workspace_bytes);
    const size_t elems = (size_t)N * H * W * 128 * 2;
    __half* bi = ar.get<__half>(elems);
    __half* b1 = ar.get<__half>(elems);
    __half* b2 = ar.get<__half>(elems);
    __half* bo = ar.get<__half>(elems);
    IC_REQUIRE(d_workspace && ar.ok(), IC_ERR_WORKSPACE, "ic_debug_conv3x3: workspace too small (need %zu)", ar.off);
    int rc = tc::launch_split_from_nhwc(d_in, N, H, W, 128, 0, bi, exact, s);
    if (rc == IC_OK && d_res1) rc = tc::launch_split_from_nhwc(d_res1, N, H, W, 128, 0, b1, exact, s);
    if (rc == IC_OK && d_res2) rc = tc::launch_split_from_nhwc(d_res2, N, H, W, 128, 0, b2, exact, s);
    if (rc == IC_OK)
        rc = conv_tc_layer(L, bi, 16, H, W, bo, nullptr, d_res1 ? b1 : nullptr, d_res2 ? b2 : nullptr, N, H, W, exact, s);
    if (rc == IC_OK) rc = tc::launch_merge_to_nhwc(bo, N, H, W, 128, d_out, exact, s);
    return rc;
}

int ic_quantize_fwd(const float* d_x, const float* d_centers, int L, float sigma, int64_t n, float* d_qsoft,
                    float* d_qhard, int64_t* d_symbols, void* stream) {
    IC_REQUIRE(d_x && d_centers && n >= 0 && L >= 1, IC_ERR_INVALID, "ic_quantize_fwd: bad argument");
    return launch_quantize(d_x, d_centers, L, sigma, n, d_qsoft, d_qhard, d_symbols, (cudaStream_t)stream);
}

// ------------------------------------------------------------------- probclass
static const char* kPcScopes[4] = {"probclass3d/logits/conv3d_conv0_mask", "probclass3d/logits/res1/conv3d_conv1_mask",
                                   "probclass3d/logits/res1/conv3d_conv2_mask", "probclass3d/logits/conv3d_conv2_mask"};

static bool pc_cfg_ok(const ic_pc_config* c) {
    return c && c->kernel_size == 3 && (c->arch_param_k == 24 || c->arch_param_k == 64) && c->num_centers >= 1 &&
           c->num_centers <= 8;
}

static void pc_dims(const ic_pc_config* c, int layer, int* ci, int* co) {
    *ci = layer == 0 ? 1 : c->arch_param_k;
    *co = layer == 3 ? c->num_centers : c->arch_param_k;
}

int ic_pc_num_tensors(const ic_pc_config* cfg) { return pc_cfg_ok(cfg) ? 8 : IC_ERR_UNSUPPORTED; }

const char* ic_pc_tensor_name(const ic_pc_config* cfg, int i) {
    static thread_local std::string s;
    if (!pc_cfg_ok(cfg) || i < 0 || i >= 8) return nullptr;
    s = std::string(kPcScopes[i / 2]) + (i % 2 ? "/biases" : "/weights");
    return s.c_str();
}

int64_t ic_pc_tensor_numel(const ic_pc_config* cfg, int i) {
    if (!pc_cfg_ok(cfg) || i < 0 || i >= 8) return IC_ERR_INVALID;
    int ci, co;
    pc_dims(cfg, i / 2, &ci, &co);
    return i % 2 ? co : (int64_t)18 * ci * co;
}

int ic_pc_create(const ic_pc_config* cfg, const float* const* h_tensors, int n_tensors, ic_pc_t** out) {
    IC_REQUIRE(cfg && h_tensors && out, IC_ERR_INVALID, "ic_pc_create: NULL argument");
    IC_REQUIRE(pc_cfg_ok(cfg), IC_ERR_UNSUPPORTED,
               "ic_pc_create: only arch res_shallow with kernel_size 3, arch_param__k in {24,64}, num_centers <= 8 is on the hot path");
    IC_REQUIRE(n_tensors == 8, IC_ERR_INVALID, "ic_pc_create: expected 8 tensors, got %d", n_tensors);
    for (int i = 0; i < 8; ++i) IC_REQUIRE(h_tensors[i], IC_ERR_INVALID, "ic_pc_create: tensor %d is NULL", i);
    ic_pc* pc = new ic_pc();
    pc->cfg = *cfg;
    pc->w.K = cfg->arch_param_k;
    pc->w.L = cfg->num_centers;
    const float** wdst[4] = {&pc->w.w0, &pc->w.w1, &pc->w.w2, &pc->w.w3};
    const float** bdst[4] = {&pc->w.b0, &pc->w.b1, &pc->w.b2, &pc->w.b3};
    for (int l = 0; l < 4; ++l) {
        int ci, co;
        pc_dims(cfg, l, &ci, &co);
        const float* w = h_tensors[2 * l];     // [2][3][3][ci][co]
        std::vector<float> packed;
        // keep only the taps the mask leaves (code/probclass.py:150-176), raster (fd,fy,fx) order
        for (int fd = 0; fd < 2; ++fd)
            for (int fy = 0; fy < 3; ++fy)
                for (int fx = 0; fx < 3; ++fx) {
                    bool masked = fd == 1 && (fy > 1 || (fy == 1 && (l == 0 ? fx >= 1 : fx > 1)));
                    if (masked) continue;
                    const float* src = w + ((size_t)(fd * 3 + fy) * 3 + fx) * ci * co;
                    packed.insert(packed.end(), src, src + (size_t)ci * co);
                }
        std::vector<float> bias(h_tensors[2 * l + 1], h_tensors[2 * l + 1] + co);
        float *dw = nullptr, *db = nullptr;
        int rc = upload(packed, &dw);
        if (rc == IC_OK) rc = upload(bias, &db);
        pc->owned.push_back(dw);
        pc->owned.push_back(db);
        if (rc != IC_OK) {
            ic_pc_destroy(pc);
            return rc;
        }
        *wdst[l] = dw;
        *bdst[l] = db;
        // tensor-core packing of layers 1..3 (arch_param__k = 24 only)
        if (l >= 1 && cfg->arch_param_k == 24 && !getenv("IC_PC_FFMA")) {
            std::vector<float> masked((size_t)18 * ci * co);
            memcpy(masked.data(), w, masked.size() * sizeof(float));
            std::vector<__half> packed;
            float inv = 1.f;
            const int nout = l == 3 ? 16 : 32;
            if (tc::pack_weights_pc(masked.data(), ci, co, nout, packed, pc->w.gt[l - 1], &inv) == IC_OK) {
                __half* dp = nullptr;
                IC_CHECK_CUDA(cudaMalloc((void**)&dp, packed.size() * sizeof(__half)));
                IC_CHECK_CUDA(cudaMemcpy(dp, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
                pc->owned_h.push_back(dp);
                std::vector<float> sct(128, 0.f), sht(128, 0.f);
                for (int o = 0; o < co; ++o) {
                    sct[o] = inv;
                    sht[o] = bias[o];
                }
                float *ds = nullptr, *dh = nullptr;
                rc = upload(sct, &ds);
                if (rc == IC_OK) rc = upload(sht, &dh);
                pc->owned.push_back(ds);
                pc->owned.push_back(dh);
                if (rc != IC_OK) {
                    ic_pc_destroy(pc);
                    return rc;
                }
                pc->w.wt[l - 1] = dp;
                pc->w.scale_t[l - 1] = ds;
                pc->w.shift_t[l - 1] = dh;
                if (l == 3) pc->w.tc = true;
            }
        }
    }
    *out = pc;
    return IC_OK;
}

void ic_pc_destroy(ic_pc_t* pc) {
    if (!pc) return;
    for (float* p : pc->owned) cudaFree(p);
    for (__half* p : pc->owned_h) cudaFree(p);
    delete pc;
}

size_t ic_pc_workspace_bytes(const ic_pc_t* pc, int N, int D, int H, int W) {
    if (!pc || N <= 0) return 0;
    // sized for the padded case (bitcost / freqs); logits() on a bare volume needs less
    return pc_workspace_bytes(pc->cfg.arch_param_k, N, D, H, W, 4, 4);
}

int ic_pc_bitcost_fwd(const ic_pc_t* pc, const float* d_q, const int64_t* d_symbols, float pad_value, int N, int C, int h,
                      int w, float* d_bits, double* d_bits_sum, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_q && d_symbols && d_bits && d_workspace, IC_ERR_INVALID, "ic_pc_bitcost_fwd: NULL argument");
    IC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_pc_bitcost_fwd: bad shape");
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = C; in.H = h; in.W = w;
    in.pad_d = 4; in.pad_hw = 4;
    in.q = d_q;
    in.pad_value = pad_value;
    in.target_symbols = d_symbols;
    return pc_forward(pc->w, in, PC_HEAD_BITCOST, d_bits, nullptr, d_bits_sum, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int ic_pc_logits_fwd(const ic_pc_t* pc, const float* d_q, int N, int D, int H, int W, float* d_logits, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_q && d_logits && d_workspace, IC_ERR_INVALID, "ic_pc_logits_fwd: NULL argument");
    IC_REQUIRE(N > 0 && D >= 5 && H >= 9 && W >= 9, IC_ERR_INVALID,
               "ic_pc_logits_fwd: volume %dx%dx%d is smaller than the 5x9x9 context", D, H, W);
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = D; in.H = H; in.W = W;
    in.q = d_q;
    return pc_forward(pc->w, in, PC_HEAD_LOGITS, d_logits, nullptr, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int ic_pc_freqs_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* d_centers, int N, int C, int h, int w,
                    int64_t* d_freqs, double* d_bits_sum, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_symbols && d_centers && d_freqs && d_workspace, IC_ERR_INVALID, "ic_pc_freqs_fwd: NULL argument");
    IC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_pc_freqs_fwd: bad shape");
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = C; in.H = h; in.W = w;
    in.pad_d = 4; in.pad_hw = 4;
    in.symbols = d_symbols;
    in.target_symbols = d_symbols;
    // centres are tiny: read them back once so the gather table can travel as a kernel argument
    IC_CHECK_CUDA(cudaMemcpyAsync(in.centers_host, d_centers, sizeof(float) * pc->cfg.num_centers, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    IC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return pc_forward(pc->w, in, PC_HEAD_FREQS, nullptr, d_freqs, d_bits_sum, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int ic_pc_codec_freqs_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* d_centers, int N, int C, int h, int w,
                          int64_t* d_freqs, double* d_bits_sum, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_symbols && d_centers && d_freqs && d_workspace, IC_ERR_INVALID, "ic_pc_codec_freqs_fwd: NULL argument");
    IC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_pc_codec_freqs_fwd: bad shape");
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = C; in.H = h; in.W = w;
    in.pad_d = 4; in.pad_hw = 4;
    in.symbols = d_symbols;
    in.target_symbols = d_symbols;
    IC_CHECK_CUDA(cudaMemcpyAsync(in.centers_host, d_centers, sizeof(float) * pc->cfg.num_centers, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    IC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return pc_forward(pc->w, in, PC_HEAD_FREQS, nullptr, d_freqs, d_bits_sum, d_workspace, workspace_bytes, (cudaStream_t)stream,
                      /*canonical=*/true);
}

int ic_pc_codec_freqs_u32_fwd(const ic_pc_t* pc, const int64_t* d_symbols, const float* h_centers, int N, int C, int h, int w,
                              uint32_t* d_freqs, double* d_bits_sum, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_symbols && h_centers && d_freqs && d_workspace, IC_ERR_INVALID, "ic_pc_codec_freqs_u32_fwd: NULL argument");
    IC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_pc_codec_freqs_u32_fwd: bad shape");
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = C; in.H = h; in.W = w;
    in.pad_d = 4; in.pad_hw = 4;
    in.symbols = d_symbols;
    in.target_symbols = d_symbols;
    for (int i = 0; i < pc->cfg.num_centers; ++i) in.centers_host[i] = h_centers[i];     // host copy: nothing to wait for
    return pc_forward(pc->w, in, PC_HEAD_FREQS, nullptr, nullptr, d_bits_sum, d_workspace, workspace_bytes, (cudaStream_t)stream,
                      /*canonical=*/true, d_freqs);
}

size_t ic_pc_decode_workspace_bytes(const ic_pc_t* pc, int N, int C, int h, int w) {
    if (!pc || N <= 0 || C <= 0 || h <= 0 || w <= 0) return 0;
    return pc_decode_workspace_bytes(N, C, h, w);
}

int ic_pc_decode_fwd(const ic_pc_t* pc, const uint8_t* d_stream, const int64_t* d_stream_offsets, const int32_t* d_first_sym,
                     const float* d_centers, int N, int C, int h, int w, uint8_t* d_symbols, const uint8_t* d_force_symbols,
                     int64_t* d_freqs_seen, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_stream && d_stream_offsets && d_first_sym && d_centers && d_symbols && d_workspace, IC_ERR_INVALID,
               "ic_pc_decode_fwd: NULL argument");
    IC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_pc_decode_fwd: bad shape");
    PcDecodeInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.C = C; in.h = h; in.w = w;
    in.stream = d_stream;
    in.stream_off = d_stream_offsets;
    in.first_sym = d_first_sym;
    in.sym_out = d_symbols;
    in.force_sym = d_force_symbols;
    in.freqs_out = d_freqs_seen;
    IC_CHECK_CUDA(cudaMemcpyAsync(in.centers_host, d_centers, sizeof(float) * pc->cfg.num_centers, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    IC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return pc_decode(pc->w, in, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int ic_pc_context_freqs_fwd(const ic_pc_t* pc, const int64_t* d_ctx_symbols, const float* d_centers, int N, int D, int H,
                            int W, int64_t* d_freqs, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(pc && d_ctx_symbols && d_centers && d_freqs && d_workspace, IC_ERR_INVALID, "ic_pc_context_freqs_fwd: NULL argument");
    IC_REQUIRE(N > 0 && D >= 5 && H >= 9 && W >= 9, IC_ERR_INVALID,
               "ic_pc_context_freqs_fwd: volume %dx%dx%d is smaller than the 5x9x9 context", D, H, W);
    PcInput in;
    memset(&in, 0, sizeof(in));
    in.N = N; in.D = D; in.H = H; in.W = W;
    in.symbols = d_ctx_symbols;
    IC_CHECK_CUDA(cudaMemcpyAsync(in.centers_host, d_centers, sizeof(float) * pc->cfg.num_centers, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    IC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return pc_forward(pc->w, in, PC_HEAD_FREQS, nullptr, d_freqs, nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

// --------------------------------------------------------------------- MS-SSIM
size_t ic_msssim_workspace_bytes(int N, int H, int W, int is_double) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return msssim_workspace_bytes(N, H, W, is_double) + 256 + align_up(sizeof(double) * 10 * (size_t)N, 256);
}

int ic_msssim_tf_fwd(const float* d_img1, const float* d_img2, int N, int H, int W, float* d_out, float* d_levels,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_img1 && d_img2 && d_out && d_workspace, IC_ERR_INVALID, "ic_msssim_tf_fwd: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_msssim_tf_fwd: bad shape");
    return msssim_tf(d_img1, d_img2, N, H, W, d_out, d_levels, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int ic_msssim_np_fwd(const uint8_t* d_img1, const uint8_t* d_img2, int N, int H, int W, double* d_out, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_img1 && d_img2 && d_out && d_workspace, IC_ERR_INVALID, "ic_msssim_np_fwd: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_msssim_np_fwd: bad shape");
    return msssim_np(d_img1, d_img2, N, H, W, d_out, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
