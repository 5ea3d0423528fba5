// upfirdn2d: zero-insert upsample -> pad/crop -> FIR (true convolution) -> decimate, NHWC.
// Replaces gan_model.py:45-50 / pytorch_upfirdn2d.py:9-51 (a six-pass composition of F.pad,
// F.conv2d with one channel and strided slicing) with one bandwidth-bound pass:
// 16-byte vectorised along the contiguous channel axis, taps broadcast from shared memory.
// The same kernel serves the adjoint (flip_kernel = 0, up <-> down), so it is closed under
// differentiation.
#include <cstring>

#include "common.cuh"

namespace b200gan {

constexpr int kMaxTaps = 16;

}  // namespace b200gan
#include "fir_epilogue.cuh"
namespace b200gan {


template <typename T, int VEC>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(
    const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ kernel, int n, int in_h,
    int in_w, int c, int out_h, int out_w, int kh, int kw, int up, int down, int pad0_y, int pad0_x,
    int flip, float gain, int64_t total_vec, FirEpilogue ep) {
    __shared__ float taps[kMaxTaps * kMaxTaps];
    for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) {
        int ky = i / kw, kx = i % kw;
        int sy = flip ? kh - 1 - ky : ky, sx = flip ? kw - 1 - kx : kx;
        taps[i] = kernel[sy * kw + sx] * gain;
    }
    __syncthreads();
    const int cv = c / VEC;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total_vec;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int ci = (int)(idx % cv);
        int64_t p = idx / cv;
        int ox = (int)(p % out_w);
        p /= out_w;
        int oy = (int)(p % out_h);
        int b = (int)(p / out_h);
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const int zy0 = oy * down - pad0_y, zx0 = ox * down - pad0_x;
        // polyphase: only taps that land on a real (non-inserted) sample, i.e. (z0 + k) % up == 0
        const int ky0 = ((-zy0 % up) + up) % up, kx0 = ((-zx0 % up) + up) % up;
        for (int ky = ky0; ky < kh; ky += up) {
            const int zy = zy0 + ky;
            if (zy < 0) continue;
            const int iy = zy / up;
            if (iy >= in_h) break;
            for (int kx = kx0; kx < kw; kx += up) {
                const int zx = zx0 + kx;
                if (zx < 0) continue;
                const int ix = zx / up;
                if (ix >= in_w) break;
                float f = taps[ky * kw + kx];
                const T* src = x + (((int64_t)b * in_h + iy) * in_w + ix) * c + (int64_t)ci * VEC;
                Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(src);
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = fmaf(f, io<T>::ld(&v.v[j]), acc[j]);
            }
        }
        if (ep.enabled) fir_epilogue<T, VEC>(ep, acc, b, ((int64_t)b * out_h + oy) * out_w + ox, c, ci * VEC);
        Pack<T, VEC> o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) io<T>::st(&o.v[j], acc[j]);
        *reinterpret_cast<Pack<T, VEC>*>(y + idx * VEC) = o;
    }
}

// RGB-sized tensors (C <= 4: the ToRGB skip path, gm.py:429, and images): one thread per output PIXEL,
// all channels in registers, polyphase tap skipping.
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_smallc_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                               const float* __restrict__ kernel, int n, int in_h,
                                                               int in_w, int c, int out_h, int out_w, int kh, int kw,
                                                               int up, int down, int pad0_y, int pad0_x, int flip,
                                                               float gain, int64_t total) {
    __shared__ float taps[kMaxTaps * kMaxTaps];
    for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) {
        int ky = i / kw, kx = i % kw;
        int sy = flip ? kh - 1 - ky : ky, sx = flip ? kw - 1 - kx : kx;
        taps[i] = kernel[sy * kw + sx] * gain;
    }
    __syncthreads();
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(idx % out_w);
        int64_t q = idx / out_w;
        const int oy = (int)(q % out_h);
        const int b = (int)(q / out_h);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int zy0 = oy * down - pad0_y, zx0 = ox * down - pad0_x;
        const int ky0 = ((-zy0 % up) + up) % up, kx0 = ((-zx0 % up) + up) % up;
        for (int ky = ky0; ky < kh; ky += up) {
            const int zy = zy0 + ky;
            if (zy < 0) continue;
            const int iy = zy / up;
            if (iy >= in_h) break;
            for (int kx = kx0; kx < kw; kx += up) {
                const int zx = zx0 + kx;
                if (zx < 0) continue;
                const int ix = zx / up;
                if (ix >= in_w) break;
                const float f = taps[ky * kw + kx];
                const T* src = x + (((int64_t)b * in_h + iy) * in_w + ix) * c;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < c) acc[j] = fmaf(f, io<T>::ld(src + j), acc[j]);
            }
        }
        T* dst = y + idx * c;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < c) io<T>::st(dst + j, acc[j]);
    }
}

// Fast path for the blur call sites (up = down = 1, <= 4x4 taps, vectorisable channel count): one
// thread owns an (x, 8-channel) column and walks ROWS output rows with a rolling set of KH row
// accumulators, so every input vector is loaded KW times (from L1) instead of KH*KW times, and a
// warp's loads are one contiguous 512-byte run of the NHWC row.
template <typename T, int VEC, int KH, int KW, int ROWS>
__global__ void __launch_bounds__(256) blur_rows_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                        const float* __restrict__ kernel, int n, int in_h, int in_w,
                                                        int c, int out_h, int out_w, int kh, int kw, int pad0_y,
                                                        int pad0_x, int flip, float gain, int64_t total,
                                                        FirEpilogue ep) {
    __shared__ float taps[KH * KW];
    __shared__ float tap_v[KH], tap_h[KW];
    __shared__ int separable;
    for (int i = threadIdx.x; i < KH * KW; i += blockDim.x) {
        int ky = i / KW, kx = i % KW;
        float v = 0.f;
        if (ky < kh && kx < kw) {
            int sy = flip ? kh - 1 - ky : ky, sx = flip ? kw - 1 - kx : kx;
            v = kernel[sy * kw + sx] * gain;
        }
        taps[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // the StyleGAN2 FIR is an outer product ([1,3,3,1] x [1,3,3,1], gm.py:60-68): detect rank 1 and
        // factor it, halving the FMA count (horizontal 4 taps, then vertical 4 taps)
        const float f00 = taps[0];
        int sep = fabsf(f00) > 1e-20f;
        for (int ky = 0; ky < KH && sep; ++ky)
            for (int kx = 0; kx < KW; ++kx) {
                const float want = taps[ky * KW] * taps[kx] / f00;
                if (fabsf(want - taps[ky * KW + kx]) > 1e-6f * fabsf(f00)) sep = 0;
            }
        separable = sep;
        for (int ky = 0; ky < KH; ++ky) tap_v[ky] = sep ? taps[ky * KW] / f00 : 0.f;
        for (int kx = 0; kx < KW; ++kx) tap_h[kx] = sep ? taps[kx] : 0.f;
    }
    __syncthreads();
    const bool sep = separable != 0;
    const int cv = c / VEC;
    const int row_blocks = (out_h + ROWS - 1) / ROWS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % cv);
        int64_t q = idx / cv;
        const int ox = (int)(q % out_w);
        q /= out_w;
        const int y0 = (int)(q % row_blocks) * ROWS;
        const int b = (int)(q / row_blocks);
        float acc[KH][VEC];
#pragma unroll
        for (int s = 0; s < KH; ++s)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[s][j] = 0.f;
        const T* xb = x + (int64_t)b * in_h * in_w * c + (int64_t)ci * VEC;
        T* yb = y + ((int64_t)b * out_h * out_w + ox) * c + (int64_t)ci * VEC;
#pragma unroll
        for (int r = 0; r < ROWS + KH - 1; ++r) {
            const int iy = y0 - pad0_y + r;
            const bool row_ok = iy >= 0 && iy < in_h;
            if (sep) {
                // horizontal pass of this input row into one vector, then scatter to the KH row accumulators
                float hrow[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) hrow[j] = 0.f;
#pragma unroll
                for (int kx = 0; kx < KW; ++kx) {
                    const int ix = ox - pad0_x + kx;
                    if (!(row_ok && ix >= 0 && ix < in_w)) continue;
                    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(xb + ((int64_t)iy * in_w + ix) * c);
                    const float f = tap_h[kx];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) hrow[j] = fmaf(f, io<T>::ld(&v.v[j]), hrow[j]);
                }
#pragma unroll
                for (int ky = 0; ky < KH; ++ky) {
                    const int orow = r - ky;
                    if (orow >= 0 && orow < ROWS) {
                        const float f = tap_v[ky];
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc[orow % KH][j] = fmaf(f, hrow[j], acc[orow % KH][j]);
                    }
                }
            } else {
#pragma unroll
                for (int kx = 0; kx < KW; ++kx) {
                    const int ix = ox - pad0_x + kx;
                    if (!(row_ok && ix >= 0 && ix < in_w)) continue;     // zero padding contributes nothing
                    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(xb + ((int64_t)iy * in_w + ix) * c);
                    float in[VEC];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) in[j] = io<T>::ld(&v.v[j]);
#pragma unroll
                    for (int ky = 0; ky < KH; ++ky) {
                        const int orow = r - ky;                   // compile-time after unrolling
                        if (orow >= 0 && orow < ROWS) {
                            const float f = taps[ky * KW + kx];
#pragma unroll
                            for (int j = 0; j < VEC; ++j) acc[orow % KH][j] = fmaf(f, in[j], acc[orow % KH][j]);
                        }
                    }
                }
            }
            if (r >= KH - 1) {
                const int orow = r - (KH - 1);
                if (y0 + orow < out_h) {
                    if (ep.enabled)
                        fir_epilogue<T, VEC>(ep, acc[orow % KH], b, ((int64_t)b * out_h + y0 + orow) * out_w + ox, c, ci * VEC);
                    Pack<T, VEC> o;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) io<T>::st(&o.v[j], acc[orow % KH][j]);
                    *reinterpret_cast<Pack<T, VEC>*>(yb + (int64_t)(y0 + orow) * out_w * c) = o;
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[orow % KH][j] = 0.f;
            }
        }
    }
}

template <typename T, int VEC>
static int launch_blur_rows(const void* x, void* y, const float* kernel, int n, int in_h, int in_w, int c, int out_h,
                            int out_w, int kh, int kw, int pad0_y, int pad0_x, int flip, float gain, const FirEpilogue& ep,
                            cudaStream_t st) {
    constexpr int ROWS = 8;
    int64_t total = (int64_t)n * ((out_h + ROWS - 1) / ROWS) * out_w * (c / VEC);
    if (total == 0) return 0;
    int64_t blocks = cdiv(total, 256);
    int64_t cap = (int64_t)sm_count() * 64;
    if (blocks > cap) blocks = cap;
    blur_rows_kernel<T, VEC, 4, 4, ROWS><<<(unsigned)blocks, 256, 0, st>>>(
        (const T*)x, (T*)y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, pad0_y, pad0_x, flip, gain, total, ep);
    count_launch();
    return check_launch("upfirdn2d(blur)");
}

template <typename T, int VEC>
static int launch_upfirdn(const void* x, void* y, const float* kernel, int n, int in_h, int in_w, int c,
                          int out_h, int out_w, int kh, int kw, int up, int down, int pad0_y,
                          int pad0_x, int flip, float gain, const FirEpilogue& ep, cudaStream_t st) {
    int64_t total = (int64_t)n * out_h * out_w * (c / VEC);
    if (total == 0) return 0;
    int64_t blocks = cdiv(total, 256);
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    upfirdn2d_kernel<T, VEC><<<(unsigned)blocks, 256, 0, st>>>(
        (const T*)x, (T*)y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down, pad0_y, pad0_x,
        flip, gain, total, ep);
    count_launch();
    return check_launch("upfirdn2d");
}

bool blur_tma_eligible(int dtype, int c, int kh, int kw, int up, int down, int out_h, int out_w, const void* x, const void* y);
int blur_tma(const void* x, void* y, const float* taps, int n, int in_h, int in_w, int c, int out_h, int out_w, int kh,
             int kw, int pad0_y, int pad0_x, int flip, float gain, const FirEpilogue& ep, cudaStream_t st);

bool fir_resample_tma_eligible(int dtype, int c, int kh, int kw, int up, int down, int out_h, int out_w, const void* x,
                               const void* y);
int fir_resample_tma(const void* x, void* y, const float* taps, int n, int in_h, int in_w, int c, int out_h, int out_w,
                     int kh, int kw, int up, int down, int pad0_y, int pad0_x, int flip, float gain, cudaStream_t st);

static int upfirdn2d_dispatch(const void* x, void* y, const float* kernel, int dtype, int n, int in_h, int in_w, int c,
                              int out_h, int out_w, int kh, int kw, int up, int down, int pad0_y, int pad0_x,
                              int flip_kernel, float gain, const FirEpilogue& ep, cudaStream_t st);

}  // namespace b200gan

extern "C" int b200gan_upfirdn2d(const void* x, void* y, const float* kernel, int dtype, int n, int in_h,
                                 int in_w, int c, int out_h, int out_w, int kh, int kw, int up,
                                 int down, int pad0_y, int pad0_x, int flip_kernel, float gain,
                                 void* stream) {
    b200gan::FirEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    return b200gan::upfirdn2d_dispatch(x, y, kernel, dtype, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down, pad0_y,
                                       pad0_x, flip_kernel, gain, ep, (cudaStream_t)stream);
}

extern "C" int b200gan_upfirdn2d_act(const void* x, void* y, const float* kernel, int dtype, int n, int in_h, int in_w,
                                     int c, int out_h, int out_w, int kh, int kw, int up, int down, int pad0_y,
                                     int pad0_x, int flip_kernel, float gain, const float* bias,
                                     const float* rowscale, const void* noise, const float* noise_w, float slope,
                                     float act_gain, void* stream) {
    b200gan::FirEpilogue ep;
    ep.bias = bias; ep.rowscale = rowscale; ep.noise = noise; ep.noise_w = noise_w;
    ep.slope = slope; ep.gain = act_gain; ep.enabled = 1;
    return b200gan::upfirdn2d_dispatch(x, y, kernel, dtype, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down, pad0_y,
                                       pad0_x, flip_kernel, gain, ep, (cudaStream_t)stream);
}

namespace b200gan {
static int upfirdn2d_dispatch(const void* x, void* y, const float* kernel, int dtype, int n, int in_h, int in_w, int c,
                              int out_h, int out_w, int kh, int kw, int up, int down, int pad0_y, int pad0_x,
                              int flip_kernel, float gain, const FirEpilogue& ep, cudaStream_t st) {
    B200_REQUIRE(kh >= 1 && kw >= 1 && kh <= kMaxTaps && kw <= kMaxTaps, "upfirdn2d: kernel %dx%d not in 1..16", kh, kw);
    B200_REQUIRE(up >= 1 && down >= 1, "upfirdn2d: up/down must be >= 1");
    B200_REQUIRE(n >= 0 && c >= 1 && in_h >= 1 && in_w >= 1 && out_h >= 0 && out_w >= 0, "upfirdn2d: bad shape");
    if (n > 0 && blur_tma_eligible(dtype, c, kh, kw, up, down, out_h, out_w, x, y))
        return blur_tma(x, y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, pad0_y, pad0_x, flip_kernel, gain, ep, st);
    if (n > 0 && !ep.enabled && fir_resample_tma_eligible(dtype, c, kh, kw, up, down, out_h, out_w, x, y))
        return fir_resample_tma(x, y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down, pad0_y, pad0_x,
                                flip_kernel, gain, st);
    return B200_DISPATCH(dtype, [&] {
        constexpr int V = 16 / sizeof(T);
        bool aligned = (c % V == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
        if (c <= 4 && c > 1 && !ep.enabled) {
            int64_t total = (int64_t)n * out_h * out_w;
            if (total == 0) return 0;
            int64_t blocks = cdiv(total, 256);
            int64_t cap = (int64_t)sm_count() * 32;
            if (blocks > cap) blocks = cap;
            upfirdn2d_smallc_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, (T*)y, kernel, n, in_h, in_w, c, out_h, out_w,
                                                                      kh, kw, up, down, pad0_y, pad0_x, flip_kernel, gain, total);
            count_launch();
            return check_launch("upfirdn2d(small c)");
        }
        if (aligned && up == 1 && down == 1 && kh <= 4 && kw <= 4 && out_h >= 8)
            return launch_blur_rows<T, V>(x, y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, pad0_y, pad0_x,
                                          flip_kernel, gain, ep, st);
        if (aligned)
            return launch_upfirdn<T, V>(x, y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down,
                                        pad0_y, pad0_x, flip_kernel, gain, ep, st);
        return launch_upfirdn<T, 1>(x, y, kernel, n, in_h, in_w, c, out_h, out_w, kh, kw, up, down,
                                    pad0_y, pad0_x, flip_kernel, gain, ep, st);
    });
}
}  // namespace b200gan
