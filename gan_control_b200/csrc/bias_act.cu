// Elementwise epilogue ops: demod row-scale + noise + bias + scaled leaky-ReLU (forward, and the
// backward expressed through the saved OUTPUT), plus the channel reductions the backward needs.
// Replaces gan_model.py:25-41 (FusedLeakyReLU / fused_leaky_relu: add, leaky_relu, mul = 3 passes),
// gm.py:340-345 (noise: 2 passes) with one pass, 16-byte vectorised on the NHWC channel axis.
#include "common.cuh"

namespace b200gan {

// x viewed as [n][hw][c][inner]; NHWC: inner = 1; NCHW: hw = 1, inner = H*W.
struct EwShape {
    int64_t n, hw, c, inner;
};

template <typename T, int VEC, bool BWD>
__global__ void __launch_bounds__(256) bias_act_kernel(
    const T* __restrict__ a,        // fwd: x        bwd: gy
    const T* __restrict__ yref,     // fwd: unused   bwd: saved output y
    T* __restrict__ out, const float* __restrict__ bias, const float* __restrict__ rowscale,
    const T* __restrict__ noise, const float* __restrict__ noise_w, EwShape s, float slope, float gain,
    int64_t total_vec) {
    const float nw = (noise != nullptr && noise_w != nullptr) ? *noise_w : 0.f;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total_vec;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = idx * VEC;
        // VEC > 1 only when inner == 1 and c % VEC == 0, so the pack shares (n, pixel) and
        // covers channels ch0 .. ch0+VEC-1
        const int64_t ch0 = (e0 / s.inner) % s.c;
        const int64_t sample = e0 / (s.hw * s.c * s.inner);
        const int64_t pix = ((e0 / (s.c * s.inner)) % s.hw) * s.inner + (e0 % s.inner);
        Pack<T, VEC> va = *reinterpret_cast<const Pack<T, VEC>*>(a + e0);
        Pack<T, VEC> vy;
        if (BWD) vy = *reinterpret_cast<const Pack<T, VEC>*>(yref + e0);
        float nz = 0.f;
        if (!BWD && noise != nullptr) nz = nw * io<T>::ld(noise + sample * (s.hw * s.inner) + pix);
        Pack<T, VEC> vo;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int64_t ch = ch0 + j;
            float rs = rowscale ? rowscale[sample * s.c + ch] : 1.f;
            float r;
            if (!BWD) {
                float v = io<T>::ld(&va.v[j]) * rs + nz + (bias ? bias[ch] : 0.f);
                r = gain * (v > 0.f ? v : v * slope);
            } else {
                float yv = io<T>::ld(&vy.v[j]);
                r = io<T>::ld(&va.v[j]) * gain * (yv > 0.f ? 1.f : slope) * rs;
            }
            io<T>::st(&vo.v[j], r);
        }
        *reinterpret_cast<Pack<T, VEC>*>(out + e0) = vo;
    }
}

template <typename T, bool BWD>
static int launch_bias_act(const void* a, const void* yref, void* out, const float* bias,
                           const float* rowscale, const void* noise, const float* noise_w, EwShape s,
                           float slope, float gain, cudaStream_t st) {
    int64_t total = s.n * s.hw * s.c * s.inner;
    if (total == 0) return 0;
    constexpr int V = 16 / sizeof(T);
    bool vec = s.inner == 1 && s.c % V == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)out % 16 == 0 &&
               (!BWD || (uintptr_t)yref % 16 == 0);
    int64_t work = vec ? total / V : total;
    int64_t blocks = cdiv(work, 256);
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (vec)
        bias_act_kernel<T, V, BWD><<<(unsigned)blocks, 256, 0, st>>>(
            (const T*)a, (const T*)yref, (T*)out, bias, rowscale, (const T*)noise, noise_w, s, slope, gain, work);
    else
        bias_act_kernel<T, 1, BWD><<<(unsigned)blocks, 256, 0, st>>>(
            (const T*)a, (const T*)yref, (T*)out, bias, rowscale, (const T*)noise, noise_w, s, slope, gain, work);
    count_launch();
    return check_launch("bias_act");
}

// ---- reductions over pixels: out_c[c] = sum_{n,hw} a*b ; out_nc[n][c] = sum_hw a*b ----------
// block = (CT channels) x (RT pixel rows); each block handles one sample and a slab of pixels,
// coalesced along channels, then a shared-memory tree over the pixel dimension, one atomic
// per (block, channel).
template <typename T>
__global__ void __launch_bounds__(256) reduce_nhwc_kernel(const T* __restrict__ a,
                                                          const T* __restrict__ b,
                                                          const T* __restrict__ pixw,
                                                          float* __restrict__ out_c,
                                                          float* __restrict__ out_nc, int64_t hw,
                                                          int64_t c, int64_t pix_per_block) {
    constexpr int CT = 32, RT = 8;
    __shared__ float red[RT][CT + 1];
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int64_t sample = blockIdx.z;
    const int64_t p0 = blockIdx.y * pix_per_block;
    const int64_t p1 = min(hw, p0 + pix_per_block);
    for (int64_t cb = blockIdx.x * CT; cb < c; cb += (int64_t)gridDim.x * CT) {
        const int64_t ch = cb + tx;
        float acc = 0.f;
        if (ch < c) {
            const T* pa = a + (sample * hw) * c + ch;
            const T* pb = b ? b + (sample * hw) * c + ch : nullptr;
            for (int64_t p = p0 + ty; p < p1; p += RT) {
                float v = io<T>::ld(pa + p * c);
                if (pb) v *= io<T>::ld(pb + p * c);
                if (pixw) v *= io<T>::ld(pixw + sample * hw + p);
                acc += v;
            }
        }
        red[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && ch < c) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < RT; ++r) t += red[r][tx];
            if (out_c) atomicAdd(out_c + ch, t);
            if (out_nc) atomicAdd(out_nc + sample * c + ch, t);
        }
        __syncthreads();
    }
}

// ---- fused backward of the StyledConv / ConvLayer epilogue -----------------------------------------
// forward was  y = gain*lrelu(z*d[n][c] + nw*noise[n][p] + bias[c]).  One pass over (gy, y) produces
//   gconv = gy * gain * act'(y) * d          (the gradient entering the convolution's backward)
//   gb[c]    += gz                            gz = gy * gain * act'(y)
//   gd[n][c] += gz * z = gz * (u - nw*noise - bias) / d,   u = pre-activation recovered from y
//   gnw      += gz * noise
// so the convolution output z never has to be stored (SURVEY.md hard part H4).
template <typename T>
__global__ void __launch_bounds__(256) epilogue_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                           T* __restrict__ gconv, const float* __restrict__ rowscale,
                                                           const T* __restrict__ noise, const float* __restrict__ noise_w,
                                                           const float* __restrict__ bias, float* __restrict__ gd,
                                                           float* __restrict__ gb, float* __restrict__ gnw, int64_t hw,
                                                           int64_t c, int64_t pix_per_block, float slope, float gain) {
    constexpr int CT = 32, RT = 8;
    __shared__ float red_b[RT][CT + 1], red_d[RT][CT + 1], red_n[RT][CT + 1];
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int64_t sample = blockIdx.z;
    const int64_t p0 = blockIdx.y * pix_per_block, p1 = min(hw, p0 + pix_per_block);
    const float nw = (noise && noise_w) ? *noise_w : 0.f;
    const float inv_gain = 1.f / gain, inv_gs = 1.f / (gain * slope);
    for (int64_t cb = blockIdx.x * CT; cb < c; cb += (int64_t)gridDim.x * CT) {
        const int64_t ch = cb + tx;
        float sb = 0.f, sd = 0.f, sn = 0.f;
        if (ch < c) {
            const float d = rowscale ? rowscale[sample * c + ch] : 1.f;
            const float bv = bias ? bias[ch] : 0.f;
            const float inv_d = 1.f / d;
            for (int64_t p = p0 + ty; p < p1; p += RT) {
                const int64_t e = (sample * hw + p) * c + ch;
                const float yv = io<T>::ld(y + e);
                const float gz = io<T>::ld(gy + e) * gain * (yv > 0.f ? 1.f : slope);
                io<T>::st(gconv + e, gz * d);
                const float nz = noise ? io<T>::ld(noise + sample * hw + p) : 0.f;
                const float u = yv > 0.f ? yv * inv_gain : yv * inv_gs;
                sb += gz;
                sd += gz * (u - nw * nz - bv) * inv_d;
                sn += gz * nz;
            }
        }
        red_b[ty][tx] = sb; red_d[ty][tx] = sd; red_n[ty][tx] = sn;
        __syncthreads();
        if (ty == 0 && ch < c) {
            float tb = 0.f, td = 0.f, tn = 0.f;
#pragma unroll
            for (int r = 0; r < RT; ++r) { tb += red_b[r][tx]; td += red_d[r][tx]; tn += red_n[r][tx]; }
            if (gb) atomicAdd(gb + ch, tb);
            if (gd) atomicAdd(gd + sample * c + ch, td);
            if (gnw && noise) atomicAdd(gnw, tn);
        }
        __syncthreads();
    }
}

// 16-byte vectorised variant: thread = (channel vector, pixel lane); per-thread partial sums, one
// shared-memory reduction per block, one global atomic per (block, channel).
// WANT_GD: the demodulation gradient sum gz*z is asked for (a per-(sample, channel) scale on the OUTPUT; the weight-modulated
// layers fold it into their weights and the discriminator has none, so the common instance skips the pre-activation recovery)
template <typename T, int VEC, bool WANT_GD>
__global__ void __launch_bounds__(256, 4) epilogue_bwd_vec_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                                  T* __restrict__ gconv, const float* __restrict__ rowscale,
                                                                  const T* __restrict__ noise, const float* __restrict__ noise_w,
                                                                  const float* __restrict__ bias, float* __restrict__ gd,
                                                                  float* __restrict__ gb, float* __restrict__ gnw, int64_t hw,
                                                                  int c, int64_t pix_per_block, float slope, float gain) {
    extern __shared__ float sred[];                   // [2][c] + 1
    float* red_b = sred;
    float* red_d = sred + c;
    float* red_n = sred + 2 * c;
    for (int i = threadIdx.x; i < 2 * c + 1; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
    const int nv = c / VEC;
    const int v = threadIdx.x % nv, pl = threadIdx.x / nv, npl = blockDim.x / nv;
    const int64_t sample = blockIdx.y;
    const int64_t p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
    const float nw = (noise && noise_w) ? *noise_w : 0.f;
    const float inv_gain = 1.f / gain, inv_gs = 1.f / (gain * slope), gain_neg = gain * slope;
    // Per-thread state is kept small (<= 64 registers, 4 blocks per SM) and two pixels' loads are issued before
    // any use: the first version (76 registers, one vector pair in flight) was latency-bound -- ncu: 8.3
    // long-scoreboard stalls per issue, 34 % warps active, 51 % of HBM (profiles/r01_ncu_kernels.md).
    // sd accumulates gz*(u - nw*noise); the bias term and 1/d are applied once at the end:
    //   sum gz*z = (sum gz*(u - nw*noise) - bias*sum gz) / d
    float dv[VEC], sb[VEC], sd[VEC];
    float sn = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        dv[j] = rowscale ? rowscale[sample * c + v * VEC + j] : 1.f;
        sb[j] = 0.f;
        sd[j] = 0.f;
    }
    if (pl < npl) {
        auto one = [&](const Pack<T, VEC>& yv, const Pack<T, VEC>& gv, float nz, int64_t e) {
            Pack<T, VEC> out;
            float gsum = 0.f;
            const float nzw = nw * nz;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float yy = io<T>::ld(&yv.v[j]);
                const float gz = io<T>::ld(&gv.v[j]) * (yy > 0.f ? gain : gain_neg);
                io<T>::st(&out.v[j], gz * dv[j]);
                sb[j] += gz;
                if (WANT_GD) {
                    const float u = yy > 0.f ? yy * inv_gain : yy * inv_gs;
                    sd[j] = fmaf(gz, u - nzw, sd[j]);
                }
                gsum += gz;
            }
            sn = fmaf(gsum, nz, sn);
            *reinterpret_cast<Pack<T, VEC>*>(gconv + e) = out;
        };
        const T* yb = y + sample * hw * c + (int64_t)v * VEC;
        const T* gb_ = gy + sample * hw * c + (int64_t)v * VEC;
        const T* nb = noise ? noise + sample * hw : nullptr;
        const int64_t ebase = sample * hw * c + (int64_t)v * VEC;
        int64_t p = p0 + pl;
        for (; p + npl < p1; p += 2 * (int64_t)npl) {
            const Pack<T, VEC> y0 = *reinterpret_cast<const Pack<T, VEC>*>(yb + p * c);
            const Pack<T, VEC> g0 = *reinterpret_cast<const Pack<T, VEC>*>(gb_ + p * c);
            const Pack<T, VEC> y1 = *reinterpret_cast<const Pack<T, VEC>*>(yb + (p + npl) * c);
            const Pack<T, VEC> g1 = *reinterpret_cast<const Pack<T, VEC>*>(gb_ + (p + npl) * c);
            const float n0 = nb ? io<T>::ld(nb + p) : 0.f;
            const float n1 = nb ? io<T>::ld(nb + p + npl) : 0.f;
            one(y0, g0, n0, ebase + p * c);
            one(y1, g1, n1, ebase + (p + npl) * c);
        }
        for (; p < p1; p += npl) {
            const Pack<T, VEC> y0 = *reinterpret_cast<const Pack<T, VEC>*>(yb + p * c);
            const Pack<T, VEC> g0 = *reinterpret_cast<const Pack<T, VEC>*>(gb_ + p * c);
            one(y0, g0, nb ? io<T>::ld(nb + p) : 0.f, ebase + p * c);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float bvj = bias ? bias[v * VEC + j] : 0.f;
            atomicAdd(&red_b[v * VEC + j], sb[j]);
            if (WANT_GD) atomicAdd(&red_d[v * VEC + j], (sd[j] - bvj * sb[j]) / dv[j]);
        }
        atomicAdd(red_n, sn);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        if (gb) atomicAdd(gb + i, red_b[i]);
        if (gd) atomicAdd(gd + sample * c + i, red_d[i]);
    }
    if (threadIdx.x == 0 && gnw && noise) atomicAdd(gnw, red_n[0]);
}

}  // namespace b200gan

extern "C" int b200gan_epilogue_bwd(const void* gy, const void* y, void* gconv, const float* rowscale, const void* noise,
                                    const float* noise_w, const float* bias, float* gd, float* gb, float* gnw, int dtype,
                                    int64_t n, int64_t hw, int64_t c, float slope, float gain, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n >= 0 && hw >= 1 && c >= 1, "epilogue_bwd: bad shape");
    if (n == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        constexpr int V = 16 / sizeof(T);
        const bool vec = c % V == 0 && c / V <= 256 && c <= 4096 && n <= 65535 &&
                         (((uintptr_t)gy | (uintptr_t)y | (uintptr_t)gconv) % 16 == 0);
        if (vec) {
            int64_t slabs = cdiv((int64_t)sm_count() * 8, n);
            int64_t ppb = cdiv(hw, slabs < 1 ? 1 : slabs);
            if (ppb < 128) ppb = 128;
            slabs = cdiv(hw, ppb);
            const size_t smem = (2 * (size_t)c + 1) * sizeof(float);
            if (gd != nullptr)
                epilogue_bwd_vec_kernel<T, V, true><<<dim3((unsigned)slabs, (unsigned)n), 256, smem, (cudaStream_t)stream>>>(
                    (const T*)gy, (const T*)y, (T*)gconv, rowscale, (const T*)noise, noise_w, bias, gd, gb, gnw, hw, (int)c, ppb,
                    slope, gain);
            else
                epilogue_bwd_vec_kernel<T, V, false><<<dim3((unsigned)slabs, (unsigned)n), 256, smem, (cudaStream_t)stream>>>(
                    (const T*)gy, (const T*)y, (T*)gconv, rowscale, (const T*)noise, noise_w, bias, gd, gb, gnw, hw, (int)c, ppb,
                    slope, gain);
            count_launch();
            return check_launch("epilogue_bwd");
        }
        int64_t cblocks = cdiv(c, 32);
        if (cblocks > 64) cblocks = 64;
        int64_t want = cdiv((int64_t)sm_count() * 8, cblocks * n);
        int64_t slabs = want < 1 ? 1 : want;
        int64_t ppb = cdiv(hw, slabs);
        if (ppb < 64) ppb = 64;
        slabs = cdiv(hw, ppb);
        B200_REQUIRE(n <= 65535 && slabs <= 65535, "epilogue_bwd: grid too large");
        dim3 grid((unsigned)cblocks, (unsigned)slabs, (unsigned)n);
        epilogue_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)gy, (const T*)y, (T*)gconv, rowscale,
                                                                       (const T*)noise, noise_w, bias, gd, gb, gnw, hw, c,
                                                                       ppb, slope, gain);
        count_launch();
        return check_launch("epilogue_bwd");
    });
}

extern "C" int b200gan_bias_act_fwd(const void* x, void* y, const float* bias, const float* rowscale,
                                    const void* noise, const float* noise_w, int dtype, int64_t n,
                                    int64_t hw, int64_t c, int64_t inner, float slope, float gain,
                                    void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n >= 0 && hw >= 1 && c >= 1 && inner >= 1, "bias_act_fwd: bad shape");
    EwShape s{n, hw, c, inner};
    return B200_DISPATCH(dtype, [&] {
        return launch_bias_act<T, false>(x, nullptr, y, bias, rowscale, noise, noise_w, s, slope, gain,
                                         (cudaStream_t)stream);
    });
}

extern "C" int b200gan_bias_act_bwd(const void* gy, const void* y, void* gx, const float* rowscale,
                                    int dtype, int64_t n, int64_t hw, int64_t c, int64_t inner,
                                    float slope, float gain, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n >= 0 && hw >= 1 && c >= 1 && inner >= 1, "bias_act_bwd: bad shape");
    EwShape s{n, hw, c, inner};
    return B200_DISPATCH(dtype, [&] {
        return launch_bias_act<T, true>(gy, y, gx, nullptr, rowscale, nullptr, nullptr, s, slope, gain,
                                        (cudaStream_t)stream);
    });
}

extern "C" int b200gan_reduce_nhwc(const void* a, const void* b, const void* pixw, float* out_c, float* out_nc, int dtype,
                                   int64_t n, int64_t hw, int64_t c, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n >= 0 && hw >= 1 && c >= 1, "reduce_nhwc: bad shape");
    if (n == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        int64_t cblocks = cdiv(c, 32);
        if (cblocks > 64) cblocks = 64;
        // enough pixel slabs to fill the machine, each at least 64 pixels
        int64_t want = cdiv((int64_t)sm_count() * 8, cblocks * n);
        int64_t slabs = want < 1 ? 1 : want;
        int64_t ppb = cdiv(hw, slabs);
        if (ppb < 64) ppb = 64;
        slabs = cdiv(hw, ppb);
        B200_REQUIRE(n <= 65535 && slabs <= 65535, "reduce_nhwc: grid too large");
        dim3 grid((unsigned)cblocks, (unsigned)slabs, (unsigned)n);
        reduce_nhwc_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (const T*)pixw,
                                                                      out_c, out_nc, hw, c, ppb);
        count_launch();
        return check_launch("reduce_nhwc");
    });
}
