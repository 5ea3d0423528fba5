// extern "C" entry points of the convolution family; picks the engine per call.
#include <atomic>

#include "conv.cuh"

namespace b200gan {
// 0 = automatic (tcgen05 when the shape is eligible), 1 = CUDA-core engine only, 2 = automatic but without
// the halo-reuse variant (testing / A-B timing)
static std::atomic<int> g_conv_engine{0};
// launches per engine (include/b200gan.h B200GAN_ENGINE_*): the evidence that a test / a step ran on the tensor cores
static std::atomic<uint64_t> g_engine_launches[B200GAN_ENGINE_COUNT];
static std::atomic<int> g_last_engine{-1};
static inline int ran(int engine, int rc) {
    if (rc == 0) {
        g_engine_launches[engine].fetch_add(1, std::memory_order_relaxed);
        g_last_engine.store(engine, std::memory_order_relaxed);
    }
    return rc;
}
}  // namespace b200gan

extern "C" uint64_t b200gan_engine_launches(int engine) {
    if (engine < 0 || engine >= B200GAN_ENGINE_COUNT) return 0;
    return b200gan::g_engine_launches[engine].load(std::memory_order_relaxed);
}
extern "C" int b200gan_last_conv_engine(void) { return b200gan::g_last_engine.load(std::memory_order_relaxed); }

extern "C" int b200gan_set_conv_engine(int engine) {
    int prev = b200gan::g_conv_engine.exchange(engine & 0xff);
    return prev;
}


static int conv_fwd_dispatch(const void* x, const void* w, void* y, int dtype, const b200gan::ConvGeom& g_in,
                             const b200gan::ConvEp& ep, cudaStream_t st) {
    using namespace b200gan;
    ConvGeom g = g_in;
    g.f16 = dtype == B200GAN_F16;
    const int b = g.b;
    const bool packed = g.pack_in || g.pack_out;
    // the tcgen05 epilogues read the output-shaped side inputs (addend / gate) as 16-byte vectors
    const bool side_ok = (((uintptr_t)ep.addend | (uintptr_t)ep.gate) & 15) == 0;
    if (!packed && g_conv_engine.load() != 1 && b > 0 && b <= 65535 && conv_fwd_pointwise_eligible(dtype, g, x, w, y))
        return ran(B200GAN_ENGINE_FWD_POINTWISE, conv_fwd_pointwise(x, w, y, dtype, g, ep, st));
    // the halo kernel's epilogue reads bias / rowscale as float4
    if (g_conv_engine.load() == 0 && b > 0 && side_ok && (((uintptr_t)ep.bias | (uintptr_t)ep.rowscale) & 15) == 0 &&
        conv_fwd_halo_eligible(dtype, g, x, w, y))
        return ran(B200GAN_ENGINE_FWD_HALO, conv_fwd_halo(x, w, y, g, ep, st));
    if (!packed && g_conv_engine.load() != 1 && b > 0 && side_ok && conv_fwd_umma_eligible(dtype, g, x, w, y))
        return ran(B200GAN_ENGINE_FWD_UMMA, conv_fwd_umma(x, w, y, g, ep, st));
    return ran(B200GAN_ENGINE_FWD_SIMT, conv_fwd_simt(x, w, y, dtype, g, ep, st));
}

static int conv_wgrad_dispatch(const void* x, const void* gy, float* gw, int dtype, const b200gan::ConvGeom& g_in,
                               cudaStream_t st) {
    using namespace b200gan;
    ConvGeom g = g_in;
    g.f16 = dtype == B200GAN_F16;
    const int b = g.b;
    const bool packed = g.pack_in || g.pack_out;
    if (!packed && g_conv_engine.load() != 1 && b > 0 && b <= 65535 && conv_wgrad_pointwise_eligible(dtype, g, x, gy))
        return ran(B200GAN_ENGINE_WGRAD_POINTWISE, conv_wgrad_pointwise(x, gy, gw, dtype, g, st));
    if (g_conv_engine.load() == 0 && b > 0 && conv_wgrad_halo_eligible(dtype, g, x, gy))
        return ran(B200GAN_ENGINE_WGRAD_HALO, conv_wgrad_halo(x, gy, gw, g, st));
    if (!packed && g_conv_engine.load() != 1 && b > 0 && conv_wgrad_umma_eligible(dtype, g, x, gy))
        return ran(B200GAN_ENGINE_WGRAD_UMMA, conv_wgrad_umma(x, gy, gw, g, st));
    return ran(B200GAN_ENGINE_WGRAD_SIMT, conv_wgrad_simt(x, gy, gw, dtype, g, st));
}

extern "C" int b200gan_conv_fwd(const void* x, const void* w, void* y, int dtype, int b, int in_h, int in_w, int ic,
                                int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                int w_per_sample, const float* bias, const float* rowscale, const void* noise,
                                const float* noise_w, float slope, float gain, void* stream) {
    b200gan::ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    const b200gan::ConvEp ep{bias, rowscale, noise, noise_w, slope, gain, nullptr, nullptr};
    return conv_fwd_dispatch(x, w, y, dtype, g, ep, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_fwd_ex(const void* x, const void* w, void* y, int dtype, int b, int in_h, int in_w, int ic,
                                   int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                   int w_per_sample, int pack_in, int pack_out, const b200gan_conv_epilogue* epilogue,
                                   void* stream) {
    b200gan::ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    g.pack_in = pack_in != 0;
    g.pack_out = pack_out != 0;
    if ((g.pack_in || g.pack_out) && (up != 1 || down != 1)) {
        b200gan::set_error("conv_fwd_ex: packed views need up = down = 1");
        return B200GAN_EINVAL;
    }
    b200gan::ConvEp ep{nullptr, nullptr, nullptr, nullptr, 1.f, 1.f, nullptr, nullptr};
    if (epilogue) ep = *epilogue;
    return conv_fwd_dispatch(x, w, y, dtype, g, ep, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_fwd_packed(const void* x, const void* w, void* y, int dtype, int b, int in_h, int in_w,
                                       int ic, int out_h, int out_w, int oc, int kh, int kw, int pad0,
                                       int w_per_sample, int pack_in, int pack_out, const float* bias,
                                       const float* rowscale, const void* noise, const float* noise_w, float slope,
                                       float gain, void* stream) {
    b200gan::ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, 1, 1, pad0, w_per_sample};
    g.pack_in = pack_in != 0;
    g.pack_out = pack_out != 0;
    const b200gan::ConvEp ep{bias, rowscale, noise, noise_w, slope, gain, nullptr, nullptr};
    return conv_fwd_dispatch(x, w, y, dtype, g, ep, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_wgrad(const void* x, const void* gy, float* gw, int dtype, int b, int in_h, int in_w,
                                  int ic, int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                  int w_per_sample, void* stream) {
    b200gan::ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    return conv_wgrad_dispatch(x, gy, gw, dtype, g, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_wgrad_packed(const void* x, const void* gy, float* gw, int dtype, int b, int in_h, int in_w,
                                         int ic, int out_h, int out_w, int oc, int kh, int kw, int pad0,
                                         int w_per_sample, int pack_x, int pack_gy, void* stream) {
    b200gan::ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, 1, 1, pad0, w_per_sample};
    g.pack_in = pack_x != 0;
    g.pack_out = pack_gy != 0;
    return conv_wgrad_dispatch(x, gy, gw, dtype, g, (cudaStream_t)stream);
}
