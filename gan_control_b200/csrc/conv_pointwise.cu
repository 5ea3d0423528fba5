// 1x1 convolutions with a tiny channel count on one side: ToRGB (C -> 3, gan_model.py:421-426), its data
// gradient (3 -> C), the discriminator's from_rgb ConvLayer (3 -> C, gm.py:943) and their weight
// gradients.  These are pure streaming passes over a 1024^2 activation (AI ~ 3 FLOP/B, SURVEY.md
// App. B): no tensor cores, no tiling -- one coalesced 16-byte-vector read of the wide tensor, the
// narrow tensor and the (per-sample) weights ride in registers / shared memory.
#include <stdlib.h>

#include "conv.cuh"

namespace b200gan {

constexpr int kPwMaxSmall = 4;
// B200GAN_PW_MODE (timing / test experiments): 0 = best kernel per shape (default), 1 = no mma.sync kernels,
// 2 = only the generic shared-memory-weight kernels
static const int g_pw_mode = [] { const char* e = getenv("B200GAN_PW_MODE"); return e ? atoi(e) : 0; }();
static const bool g_pw_generic = g_pw_mode >= 2;

struct PwEpilogue {
    const float* bias;
    const float* rowscale;
    const void* noise;
    const float* noise_w;
    float slope, gain;
    const void* addend;      // output-shaped (include/b200gan.h b200gan_conv_epilogue)
    const void* gate;        // output-shaped: backward mode
    int on;
};

// output element (pix, o) of `oc_total` channels
template <typename T>
__device__ __forceinline__ float pw_epilogue(const PwEpilogue& e, float v, int b, int oc_total, int o, int64_t pix, float nz) {
    if (!e.on) return v;
    if (e.addend) v += io<T>::ld((const T*)e.addend + pix * oc_total + o);
    if (e.rowscale) v *= e.rowscale[(int64_t)b * oc_total + o];
    if (e.gate) return v * (io<T>::ld((const T*)e.gate + pix * oc_total + o) > 0.f ? e.gain : e.gain * e.slope);
    v += nz + (e.bias ? e.bias[o] : 0.f);
    return e.gain * (v > 0.f ? v : v * e.slope);
}

// ---- OC <= 4:  y[p][o] = sum_c x[p][c] * w[wb][o][c] ------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pw_small_oc_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          T* __restrict__ y, int npix, int ic, int oc,
                                                          int per_sample, PwEpilogue ep) {
    extern __shared__ float wsm[];                      // [oc][ic]
    const int b = blockIdx.y;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    for (int i = threadIdx.x; i < oc * ic; i += blockDim.x) wsm[i] = io<T>::ld(wb + i);
    __syncthreads();
    const int nv = ic / VEC;
    int L = 1;
    while (L < 32 && L < nv) L <<= 1;                   // lanes cooperating on one pixel
    const int lane = threadIdx.x & 31, sub = lane % L, pw_ = lane / L, ppw = 32 / L;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const float nw = (ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    for (int p0 = warp * ppw; p0 < npix; p0 += warps * ppw) {
        const int p = p0 + pw_;
        float acc[kPwMaxSmall] = {0.f, 0.f, 0.f, 0.f};
        if (p < npix) {
            const T* xp = x + ((int64_t)b * npix + p) * ic;
            for (int v = sub; v < nv; v += L) {
                const Pack<T, VEC> xv = *reinterpret_cast<const Pack<T, VEC>*>(xp + v * VEC);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float xf = io<T>::ld(&xv.v[j]);
#pragma unroll
                    for (int o = 0; o < kPwMaxSmall; ++o)
                        if (o < oc) acc[o] = fmaf(xf, wsm[o * ic + v * VEC + j], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kPwMaxSmall; ++o)
            for (int off = L >> 1; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
        if (sub == 0 && p < npix) {
            const int64_t pix = (int64_t)b * npix + p;
            const float nz = ep.noise ? nw * io<T>::ld((const T*)ep.noise + pix) : 0.f;
            for (int o = 0; o < oc; ++o) io<T>::st(y + pix * oc + o, pw_epilogue<T>(ep, acc[o], b, oc, o, pix, nz));
        }
    }
}

// ---- IC <= 4:  y[p][o..o+VEC) = sum_c x[p][c] * w[wb][o][c] -----------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pw_small_ic_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                          T* __restrict__ y, int npix, int ic, int oc,
                                                          int per_sample, PwEpilogue ep) {
    extern __shared__ float wsm[];                      // [oc][ic]
    const int b = blockIdx.y;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    for (int i = threadIdx.x; i < oc * ic; i += blockDim.x) wsm[i] = io<T>::ld(wb + i);
    __syncthreads();
    const int nv = oc / VEC;
    const float nw = (ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    const int64_t total = (int64_t)npix * nv;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(idx % nv);
        const int p = (int)(idx / nv);
        const int64_t pix = (int64_t)b * npix + p;
        float xs[kPwMaxSmall];
#pragma unroll
        for (int c = 0; c < kPwMaxSmall; ++c) xs[c] = c < ic ? io<T>::ld(x + pix * ic + c) : 0.f;
        const float nz = ep.noise ? nw * io<T>::ld((const T*)ep.noise + pix) : 0.f;
        Pack<T, VEC> out;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int o = v * VEC + j;
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < kPwMaxSmall; ++c)
                if (c < ic) a = fmaf(xs[c], wsm[o * ic + c], a);
            io<T>::st(&out.v[j], pw_epilogue<T>(ep, a, b, oc, o, pix, nz));
        }
        *reinterpret_cast<Pack<T, VEC>*>(y + pix * oc + v * VEC) = out;
    }
}

// ---- fast paths: the weight slice a thread needs is loop-invariant -> registers ---------------------------------------
// The generic kernels above re-read their weights from shared memory for every vector (24 scalar LDS per 16 bytes of
// activation) and, in the IC <= 4 one, decode the (pixel, vector) pair with two 64-bit divisions per element: both were
// issue-bound at 1.2 - 1.7 TB/s (from_rgb forward at 1024^2: 0.91 ms for a 1.2 GB pass, profiles/r02_pointwise.md).
// Here the channel vector of a thread is fixed for the whole kernel (power-of-two vector counts), its weights, bias and
// demodulation row live in registers, side inputs are read as 16-byte vectors, and several pixels are in flight per thread.

// OC <= 4 (ToRGB forward, from_rgb data gradient): L = nv / SL lanes share one pixel, lane `sub` owns the vectors
// sub, sub + L, ... (SL of them); partial sums meet in a shuffle reduction.
template <typename T, int VEC, int SL>
__global__ void __launch_bounds__(256, 2) pw_small_oc_reg_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                              T* __restrict__ y, int npix, int ic, int oc, int per_sample,
                                                              int L, PwEpilogue ep) {
    constexpr int U = SL >= 4 ? 1 : (SL == 2 ? 2 : 4);           // pixel groups in flight per warp
    const int b = blockIdx.y;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    const int lane = threadIdx.x & 31, sub = lane & (L - 1), ppw = 32 / L, pw_ = lane / L;
    float wr[SL][kPwMaxSmall][VEC];
#pragma unroll
    for (int s = 0; s < SL; ++s)
#pragma unroll
        for (int o = 0; o < kPwMaxSmall; ++o)
#pragma unroll
            for (int j = 0; j < VEC; ++j) wr[s][o][j] = o < oc ? io<T>::ld(wb + (int64_t)o * ic + (sub + s * L) * VEC + j) : 0.f;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const float nw = (ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    const T* xb = x + (int64_t)b * npix * ic;
    for (int p0 = warp * ppw; p0 < npix; p0 += warps * ppw * U) {
        Pack<T, VEC> xv[U][SL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * warps * ppw + pw_;
#pragma unroll
            for (int s = 0; s < SL; ++s)
                if (p < npix) xv[u][s] = *reinterpret_cast<const Pack<T, VEC>*>(xb + (int64_t)p * ic + (sub + s * L) * VEC);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * warps * ppw + pw_;
            float acc[kPwMaxSmall] = {0.f, 0.f, 0.f, 0.f};
            if (p < npix) {
#pragma unroll
                for (int s = 0; s < SL; ++s)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const float xf = io<T>::ld(&xv[u][s].v[j]);
#pragma unroll
                        for (int o = 0; o < kPwMaxSmall; ++o) acc[o] = fmaf(xf, wr[s][o][j], acc[o]);
                    }
            }
#pragma unroll
            for (int o = 0; o < kPwMaxSmall; ++o)
                for (int off = L >> 1; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
            if (sub == 0 && p < npix) {
                const int64_t pix = (int64_t)b * npix + p;
                const float nz = ep.noise ? nw * io<T>::ld((const T*)ep.noise + pix) : 0.f;
#pragma unroll
                for (int o = 0; o < kPwMaxSmall; ++o)
                    if (o < oc) io<T>::st(y + pix * oc + o, pw_epilogue<T>(ep, acc[o], b, oc, o, pix, nz));
            }
        }
    }
}

// IC <= 4 (from_rgb forward, ToRGB data gradient): a thread owns output vector v = tid % nv and walks over pixels.
template <typename T, int VEC, int ICN>
__global__ void __launch_bounds__(256, 2) pw_small_ic_reg_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                              T* __restrict__ y, int npix, int ic, int oc, int per_sample,
                                                              int log_nv, PwEpilogue ep) {
    constexpr int U = 3;
    const int b = blockIdx.y;
    const int nv = 1 << log_nv;
    const int v = threadIdx.x & (nv - 1), pl = threadIdx.x >> log_nv, npl = blockDim.x >> log_nv;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    float wr[VEC][ICN], br[VEC], rr[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int o = v * VEC + j;
#pragma unroll
        for (int c = 0; c < ICN; ++c) wr[j][c] = c < ic ? io<T>::ld(wb + (int64_t)o * ic + c) : 0.f;
        br[j] = (ep.on && ep.bias && !ep.gate) ? ep.bias[o] : 0.f;
        rr[j] = (ep.on && ep.rowscale) ? ep.rowscale[(int64_t)b * oc + o] : 1.f;
    }
    const float nw = (ep.on && ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    const float g_pos = ep.gain, g_neg = ep.gain * ep.slope;
    const T* xb = x + (int64_t)b * npix * ic;
    const int stride = gridDim.x * npl;
    for (int p0 = blockIdx.x * npl + pl; p0 < npix; p0 += stride * U) {
        float xs[U][ICN], nz[U];
        Pack<T, VEC> sa[U], sg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * stride;
            nz[u] = 0.f;
#pragma unroll
            for (int c = 0; c < ICN; ++c) xs[u][c] = (c < ic && p < npix) ? io<T>::ld(xb + (int64_t)p * ic + c) : 0.f;
            if (ep.on && p < npix) {
                const int64_t pix = (int64_t)b * npix + p;
                if (ep.noise) nz[u] = nw * io<T>::ld((const T*)ep.noise + pix);
                if (ep.addend) sa[u] = *reinterpret_cast<const Pack<T, VEC>*>((const T*)ep.addend + pix * oc + v * VEC);
                if (ep.gate) sg[u] = *reinterpret_cast<const Pack<T, VEC>*>((const T*)ep.gate + pix * oc + v * VEC);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * stride;
            if (p >= npix) continue;
            Pack<T, VEC> out;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < ICN; ++c) a = fmaf(xs[u][c], wr[j][c], a);
                if (ep.on) {                                    // same order of operations as pw_epilogue
                    if (ep.addend) a += io<T>::ld(&sa[u].v[j]);
                    a *= rr[j];
                    if (ep.gate) {
                        a *= io<T>::ld(&sg[u].v[j]) > 0.f ? g_pos : g_neg;
                    } else {
                        a += nz[u] + br[j];
                        a = a > 0.f ? a * g_pos : a * g_neg;
                    }
                }
                io<T>::st(&out.v[j], a);
            }
            *reinterpret_cast<Pack<T, VEC>*>(y + ((int64_t)b * npix + p) * oc + v * VEC) = out;
        }
    }
}

// ---- 16-bit storage: the same two passes on mma.sync (m16n8k16 / m16n8k8, fp32 accumulate) --------------------------------
// The register-weight kernels above still spend ~100 issue slots per 16 bytes of activation (conversions, FFMAs, shuffle
// reductions) and measured 1.2 - 2.5 TB/s (profiles/r02_pointwise.md).  A warp-level MMA does the 16-pixel x 16-channel
// products in one instruction straight from the registers the 16-byte loads filled: the k index of a dot product is
// arbitrary as long as both operands agree, so each lane's 16-byte chunk (8 consecutive channels) IS its share of two
// k-steps, for the activation rows (g, g + 8) and for the weight row n = g alike (fragment layout of PTX
// mma.m16n8k16: lane = 4 g + t; the index algebra is checked by emulation in tests/test_pointwise_mapping_cpu.py).
// tcgen05 has no place here: M = 16 pixel granularity, K <= 32, nothing to stage in shared memory.
template <typename T> struct WarpMma;
template <> struct WarpMma<__nv_bfloat16> {
    static __device__ __forceinline__ void k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
    }
};
template <> struct WarpMma<__half> {
    static __device__ __forceinline__ void k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
    }
};
template <typename T> __device__ __forceinline__ uint32_t pw_pair(const T* p, bool lo_ok, bool hi_ok) {
    const unsigned short lo = lo_ok ? *reinterpret_cast<const unsigned short*>(p) : (unsigned short)0;
    const unsigned short hi = hi_ok ? *reinterpret_cast<const unsigned short*>(p + 1) : (unsigned short)0;
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

// OC <= 4, IC = 32 * NQ: a warp takes 16 pixels per step (rows g and g + 8 of the fragment); U steps in flight for the
// narrow layers, QC 32-channel chunks in flight for the wide ones (eight 16-byte loads per lane either way).
template <typename T, int NQ>
__global__ void __launch_bounds__(256) pw_small_oc_mma_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y,
                                                              int npix, int ic, int oc, int per_sample, PwEpilogue ep) {
    constexpr int U = NQ >= 4 ? 1 : 4 / NQ;
    constexpr int QC = NQ >= 4 ? 4 : NQ;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    uint4 wq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
        wq[q] = g < oc ? *reinterpret_cast<const uint4*>(wb + (int64_t)g * ic + q * 32 + t * 8) : make_uint4(0u, 0u, 0u, 0u);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int groups = (npix + 15) >> 4;
    const float nw = (ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    const T* xb = x + (int64_t)b * npix * ic;
    for (int g0 = warp; g0 < groups; g0 += warps * U) {
        float c[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0.f;
#pragma unroll
        for (int q0 = 0; q0 < NQ; q0 += QC) {
            uint4 xa[U][QC], xh[U][QC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p_lo = (g0 + u * warps) * 16 + g, p_hi = p_lo + 8;
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    xa[u][q] = p_lo < npix ? __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)p_lo * ic + (q0 + q) * 32 + t * 8))
                                           : make_uint4(0u, 0u, 0u, 0u);
                    xh[u][q] = p_hi < npix ? __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)p_hi * ic + (q0 + q) * 32 + t * 8))
                                           : make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    WarpMma<T>::k16(c[u], xa[u][q].x, xh[u][q].x, xa[u][q].y, xh[u][q].y, wq[q0 + q].x, wq[q0 + q].y);
                    WarpMma<T>::k16(c[u], xa[u][q].z, xh[u][q].z, xa[u][q].w, xh[u][q].w, wq[q0 + q].z, wq[q0 + q].w);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // c[0], c[1]: pixel row g, output channels 2t, 2t + 1;  c[2], c[3]: row g + 8
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = (g0 + u * warps) * 16 + g + 8 * h;
                if (p >= npix || 2 * t >= oc) continue;
                const int64_t pix = (int64_t)b * npix + p;
                const float nz = ep.noise ? nw * io<T>::ld((const T*)ep.noise + pix) : 0.f;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int o = 2 * t + e;
                    if (o < oc) io<T>::st(y + pix * oc + o, pw_epilogue<T>(ep, c[u][2 * h + e], b, oc, o, pix, nz));
                }
            }
        }
    }
}

// IC <= 4, OC = 32 * NQ: output-channel tiles are permuted so that lane (g, t) ends up with the 8 CONSECUTIVE channels
// 32 q + 8 t .. + 8 of rows g and g + 8 (one 16-byte store each): column n of tile (q, j) is channel 32 q + 8 (n / 2) + 2 j + n % 2.
template <typename T, int NQ>
__global__ void __launch_bounds__(256) pw_small_ic_mma_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y,
                                                              int npix, int ic, int oc, int per_sample, PwEpilogue ep) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const T* wb = w + (int64_t)(per_sample ? b : 0) * oc * ic;
    uint32_t bq[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = q * 32 + (g >> 1) * 8 + j * 2 + (g & 1);
            bq[q][j] = pw_pair<T>(wb + (int64_t)o * ic + 2 * t, 2 * t < ic, 2 * t + 1 < ic);
        }
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int groups = (npix + 15) >> 4;
    const float nw = (ep.on && ep.noise && ep.noise_w) ? *ep.noise_w : 0.f;
    const float g_pos = ep.gain, g_neg = ep.gain * ep.slope;
    const bool lrelu_ok = ep.gain > 0.f && ep.slope >= 0.f && ep.slope <= 1.f;      // gain * lrelu(u) == max(g u, g slope u)
    const T* xb = x + (int64_t)b * npix * ic;
    // (requesting the narrow operand of the NEXT step one iteration ahead was measured SLOWER: 0.48 vs 0.40 ms for
    // 3 -> 32 @1024^2, profiles/r02_pointwise.md; the pass is bound by its 16-byte stores, 32 resident warps cover the loads)
    for (int g0 = warp; g0 < groups; g0 += warps) {
        const int p_lo = g0 * 16 + g, p_hi = p_lo + 8;
        const uint32_t a0 = p_lo < npix ? pw_pair<T>(xb + (int64_t)p_lo * ic + 2 * t, 2 * t < ic, 2 * t + 1 < ic) : 0u;
        const uint32_t a1 = p_hi < npix ? pw_pair<T>(xb + (int64_t)p_hi * ic + 2 * t, 2 * t < ic, 2 * t + 1 < ic) : 0u;
        float nz[2] = {0.f, 0.f};
        if (ep.on && ep.noise) {
            if (p_lo < npix) nz[0] = nw * io<T>::ld((const T*)ep.noise + (int64_t)b * npix + p_lo);
            if (p_hi < npix) nz[1] = nw * io<T>::ld((const T*)ep.noise + (int64_t)b * npix + p_hi);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
                WarpMma<T>::k8(c[j], a0, a1, bq[q][j]);
            }
            const int o0 = q * 32 + t * 8;
            float br[8], rr[8];
            if (ep.on) {
#pragma unroll
                for (int e = 0; e < 8; e += 4) {
                    const float4 bv = (ep.bias && !ep.gate) ? __ldg(reinterpret_cast<const float4*>(ep.bias + o0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 rv = ep.rowscale ? __ldg(reinterpret_cast<const float4*>(ep.rowscale + (int64_t)b * oc + o0 + e))
                                                  : make_float4(1.f, 1.f, 1.f, 1.f);
                    br[e] = bv.x; br[e + 1] = bv.y; br[e + 2] = bv.z; br[e + 3] = bv.w;
                    rr[e] = rv.x; rr[e + 1] = rv.y; rr[e + 2] = rv.z; rr[e + 3] = rv.w;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = h ? p_hi : p_lo;
                if (p >= npix) continue;
                const int64_t off = ((int64_t)b * npix + p) * oc + o0;
                Pack<T, 8> out;
                Pack<T, 8> sa, sg;
                if (ep.on && ep.addend) sa = *reinterpret_cast<const Pack<T, 8>*>((const T*)ep.addend + off);
                if (ep.on && ep.gate) sg = *reinterpret_cast<const Pack<T, 8>*>((const T*)ep.gate + off);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float a = c[e >> 1][2 * h + (e & 1)];
                    if (ep.on) {                                // same order of operations as pw_epilogue
                        if (ep.addend) a += io<T>::ld(&sa.v[e]);
                        a *= rr[e];
                        if (ep.gate) {
                            a *= io<T>::ld(&sg.v[e]) > 0.f ? g_pos : g_neg;
                        } else {
                            a += nz[h] + br[e];
                            a = lrelu_ok ? fmaxf(a * g_pos, a * g_neg) : (a > 0.f ? a * g_pos : a * g_neg);
                        }
                    }
                    io<T>::st(&out.v[e], a);
                }
                *reinterpret_cast<Pack<T, 8>*>(y + off) = out;
            }
        }
    }
}

// launchers of the 16-bit MMA kernels; false when the shape is not theirs
template <typename T>
static bool pw_launch_mma(const T* x, const T* w, T* y, int npix, const ConvGeom& g, const PwEpilogue& ep, cudaStream_t st) {
    int bx = sm_count() * 8 / g.b;                                  // whole waves only (see conv_wgrad_pointwise)
    const int need = (int)cdiv(cdiv(npix, 16), 8);
    if (bx > need) bx = need;
    const dim3 grid(bx < 1 ? 1 : bx, g.b);
    if (g.oc <= kPwMaxSmall && g.ic % 32 == 0 && g.ic <= 512) {
        switch (g.ic / 32) {
            case 1: pw_small_oc_mma_kernel<T, 1><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 2: pw_small_oc_mma_kernel<T, 2><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 4: pw_small_oc_mma_kernel<T, 4><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 8: pw_small_oc_mma_kernel<T, 8><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 16: pw_small_oc_mma_kernel<T, 16><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            default: break;
        }
    }
    const bool side_aligned = (((uintptr_t)ep.bias | (uintptr_t)ep.rowscale | (uintptr_t)ep.addend | (uintptr_t)ep.gate) & 15) == 0;
    if (g.ic <= kPwMaxSmall && g.oc % 32 == 0 && g.oc <= 128 && side_aligned) {
        switch (g.oc / 32) {
            case 1: pw_small_ic_mma_kernel<T, 1><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 2: pw_small_ic_mma_kernel<T, 2><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            case 4: pw_small_ic_mma_kernel<T, 4><<<grid, 256, 0, st>>>(x, w, y, npix, g.ic, g.oc, g.w_per_sample, ep); return true;
            default: return false;
        }
    }
    return false;
}
template <>
bool pw_launch_mma<float>(const float*, const float*, float*, int, const ConvGeom&, const PwEpilogue&, cudaStream_t) { return false; }

static inline int pw_log2_exact(int v) {                       // log2 of a power of two, else -1
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// ---- weight gradient with one narrow side:  gw[wb][o][c] += sum_p gy[p][o] * x[p][c] ----------------
// `wide` has wc channels (vectorised), `narrow` has nc <= 4.  narrow_is_oc selects the output indexing.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const T* __restrict__ wide, const T* __restrict__ narrow,
                                                       float* __restrict__ gw, int npix, int wc, int nc,
                                                       int narrow_is_oc, int per_sample, int pix_per_block) {
    extern __shared__ float red[];                      // [nc][wc]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < nc * wc; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int nv = wc / VEC;                            // vectors per pixel (<= 256)
    const int v = threadIdx.x % nv, pl = threadIdx.x / nv, npl = blockDim.x / nv;
    const int p_begin = blockIdx.x * pix_per_block, p_end = min(npix, p_begin + pix_per_block);
    float acc[kPwMaxSmall][VEC];
#pragma unroll
    for (int s = 0; s < kPwMaxSmall; ++s)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[s][j] = 0.f;
    if (pl < npl) {
        // (four pixels in flight per thread were measured SLOWER, 0.53 vs 0.31 ms at 3 <-> 32 @1024^2: 102 registers, half
        // the resident warps -- profiles/r02_pointwise.md)
        for (int p = p_begin + pl; p < p_end; p += npl) {
            const int64_t pix = (int64_t)b * npix + p;
            const Pack<T, VEC> wv = *reinterpret_cast<const Pack<T, VEC>*>(wide + pix * wc + v * VEC);
            float ns[kPwMaxSmall];
#pragma unroll
            for (int s = 0; s < kPwMaxSmall; ++s) ns[s] = s < nc ? io<T>::ld(narrow + pix * nc + s) : 0.f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float wf = io<T>::ld(&wv.v[j]);
#pragma unroll
                for (int s = 0; s < kPwMaxSmall; ++s) acc[s][j] = fmaf(wf, ns[s], acc[s][j]);
            }
        }
#pragma unroll
        for (int s = 0; s < kPwMaxSmall; ++s)
            if (s < nc)
#pragma unroll
                for (int j = 0; j < VEC; ++j) atomicAdd(&red[s * wc + v * VEC + j], acc[s][j]);
    }
    __syncthreads();
    const int oc = narrow_is_oc ? nc : wc, ic = narrow_is_oc ? wc : nc;
    float* dst = gw + (int64_t)(per_sample ? b : 0) * oc * ic;
    for (int i = threadIdx.x; i < nc * wc; i += blockDim.x) {
        const int s = i / wc, cw = i % wc;
        const int o = narrow_is_oc ? s : cw, c = narrow_is_oc ? cw : s;
        atomicAdd(dst + (int64_t)o * ic + c, red[i]);
    }
}

// ---------------------------------------------------------------------------------------------
static bool pw_shape(const ConvGeom& g) {
    return g.kh == 1 && g.kw == 1 && g.up == 1 && g.down == 1 && g.pad0 == 0 && g.out_h == g.in_h && g.out_w == g.in_w;
}

bool conv_fwd_pointwise_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y) {
    if (!pw_shape(g)) return false;
    const int vec = dtype == B200GAN_F32 ? 4 : 8;
    if ((int64_t)g.oc * g.ic * 4 > 40 * 1024) return false;      // weights are staged in (default-limit) shared memory
    if (g.oc <= kPwMaxSmall && g.ic % vec == 0 && (uintptr_t)x % 16 == 0) return true;
    if (g.ic <= kPwMaxSmall && g.oc % vec == 0 && (uintptr_t)y % 16 == 0) return true;
    return false;
}

int conv_fwd_pointwise(const void* x, const void* w, void* y, int dtype, const ConvGeom& g, const ConvEp& e, cudaStream_t st) {
    PwEpilogue ep{e.bias, e.rowscale, e.noise, e.noise_w, e.slope, e.gain, e.addend, e.gate, ep_active(e) ? 1 : 0};
    const int npix = g.out_h * g.out_w;
    return B200_DISPATCH(dtype, [&] {
        constexpr int V = 16 / sizeof(T);
        const size_t smem = (size_t)g.oc * g.ic * sizeof(float);
        int bx = sm_count() * 8 / g.b;
        if (g_pw_mode == 0 && pw_launch_mma<T>((const T*)x, (const T*)w, (T*)y, npix, g, ep, st)) {
            count_launch();
            return check_launch("conv_fwd_pointwise");
        }
        if (g.oc <= kPwMaxSmall) {
            const int nv = g.ic / V, lnv = pw_log2_exact(nv);
            if (lnv >= 0 && nv <= 128 && !g_pw_generic) {
                // register-weight kernel: SL vectors per lane, L = nv / SL lanes per pixel
                const int sl = nv <= 32 ? 1 : nv / 32, L = nv / sl, ppw = 32 / L;
                int need = (int)cdiv(npix, 8 * ppw * (sl >= 4 ? 1 : 4 / sl));
                if (bx > need) bx = need;
                const dim3 grid(bx < 1 ? 1 : bx, g.b);
                if (sl == 1)
                    pw_small_oc_reg_kernel<T, V, 1><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc, g.w_per_sample, L, ep);
                else if (sl == 2)
                    pw_small_oc_reg_kernel<T, V, 2><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc, g.w_per_sample, L, ep);
                else
                    pw_small_oc_reg_kernel<T, V, 4><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc, g.w_per_sample, L, ep);
            } else {
                int need = (int)cdiv(npix, 64);
                if (bx > need) bx = need;
                pw_small_oc_kernel<T, V><<<dim3(bx < 1 ? 1 : bx, g.b), 256, smem, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc,
                                                                                      g.w_per_sample, ep);
            }
        } else {
            const int nv = g.oc / V, lnv = pw_log2_exact(nv);
            if (lnv >= 0 && nv <= 256 && !g_pw_generic) {
                const int npl = 256 >> lnv;
                int need = (int)cdiv(npix, npl * 3);
                if (bx > need) bx = need;
                const dim3 grid(bx < 1 ? 1 : bx, g.b);
                if (g.ic <= 3)
                    pw_small_ic_reg_kernel<T, V, 3><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc, g.w_per_sample, lnv, ep);
                else
                    pw_small_ic_reg_kernel<T, V, 4><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc, g.w_per_sample, lnv, ep);
            } else {
                int need = (int)cdiv((int64_t)npix * (g.oc / V), 256);
                if (bx > need) bx = need;
                pw_small_ic_kernel<T, V><<<dim3(bx < 1 ? 1 : bx, g.b), 256, smem, st>>>((const T*)x, (const T*)w, (T*)y, npix, g.ic, g.oc,
                                                                                      g.w_per_sample, ep);
            }
        }
        count_launch();
        return check_launch("conv_fwd_pointwise");
    });
}

bool conv_wgrad_pointwise_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy) {
    if (!pw_shape(g)) return false;
    const int vec = dtype == B200GAN_F32 ? 4 : 8;
    if (g.oc <= kPwMaxSmall && g.ic % vec == 0 && g.ic / vec <= 256 && (uintptr_t)x % 16 == 0) return true;
    if (g.ic <= kPwMaxSmall && g.oc % vec == 0 && g.oc / vec <= 256 && (uintptr_t)gy % 16 == 0) return true;
    return false;
}

int conv_wgrad_pointwise(const void* x, const void* gy, float* gw, int dtype, const ConvGeom& g, cudaStream_t st) {
    const int npix = g.out_h * g.out_w;
    return B200_DISPATCH(dtype, [&] {
        constexpr int V = 16 / sizeof(T);
        const bool narrow_oc = g.oc <= kPwMaxSmall;
        const T* wide = narrow_oc ? (const T*)x : (const T*)gy;
        const T* narrow = narrow_oc ? (const T*)gy : (const T*)x;
        const int wc = narrow_oc ? g.ic : g.oc, nc = narrow_oc ? g.oc : g.ic;
        // at most ONE wave of resident blocks (4 per SM): rounding the per-sample count up left a 16-block second wave at
        // batch 32 and doubled the kernel (1.06 ms vs 2 x 0.31 ms at batch 16, gpurun_out/r2p_variants.log)
        int blocks = sm_count() * 4 / g.b;
        int ppb = (int)cdiv(npix, blocks < 1 ? 1 : blocks);
        if (ppb < 256) ppb = 256;
        blocks = (int)cdiv(npix, ppb);
        const size_t smem = (size_t)nc * wc * sizeof(float);
        pw_wgrad_kernel<T, V><<<dim3(blocks, g.b), 256, smem, st>>>(wide, narrow, gw, npix, wc, nc, narrow_oc ? 1 : 0,
                                                                 g.w_per_sample, ppb);
        count_launch();
        return check_launch("conv_wgrad_pointwise");
    });
}

}  // namespace b200gan
