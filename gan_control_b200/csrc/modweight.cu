// Weight (de)modulation of ModulatedConv2d in one kernel each way (gan_model.py:284-289):
//     w_eff[b][o][i][t] = scale * W[o][i][t] * s[b][i] * d[b][o],      d[b][o] = rsqrt(sum_{i,t} (scale*W*s)^2 + 1e-8)
// written straight into the K-major operand layout of the convolution engines ([b][tap][o][i], optionally with the taps
// reversed = the gather form of the transposed convolution, gm.py:301-306) and, for the data gradient, into the adjoint
// layout ([b][KK-1-tap][i][o]).  The reference builds this with five broadcast / reduction passes over a
// (B,OC,IC,k,k) tensor; round 1 of this repo did the same with ~12 small PyTorch kernels per layer and ~25 more in the
// backward pass (profiles/r01_launches_step_final.md: 2400 tiny kernels per step).
//
// Backward (first order), given g[b][tap][o][i] = dL/dw_eff in the same K-major layout (fp32, the conv weight gradient):
//     gd[b][o]   = sum_{i,t} g * w0 * s                 (w0 = scale*W)         e[b][o] = -gd * d^3   (demodulation)
//     gs[b][i]   = sum_{o,t} g * w0 * d  +  sum_o e[b][o] * s[b][i] * q[o][i]   (q = sum_t w0^2)
//     gW[o][i][t]= scale * sum_b ( g * s * d  +  e[b][o] * s[b][i]^2 * w0[o][i][t] )
#include "common.cuh"

namespace b200gan {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (ceil(OC / 32), B), 256 threads = 8 warps.  KK = kh*kw <= 9.  A block owns 32 output channels of one sample:
// (1) each warp reduces the demodulation sums of 4 of them; (2) 32x32 (o, i) tiles are written to wk coalesced over i
// and, through a shared-memory transpose, to the adjoint operand coalesced over o (a direct store of the adjoint is a
// 2-byte scatter with stride OC: one 32-byte sector per element).
template <typename T>
__global__ void __launch_bounds__(256) modweight_fwd_kernel(const float* __restrict__ W, const float* __restrict__ s,
                                                            float* __restrict__ d_out, T* __restrict__ wk,
                                                            T* __restrict__ wkT, int B, int OC, int IC, int KK,
                                                            float scale, int demod, int flip, float eps) {
    const int o0 = blockIdx.x * 32, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* sb = s + (int64_t)b * IC;
    __shared__ float dsm[32];
    __shared__ float tile[32][33];
    for (int r = warp; r < 32; r += 8) {
        const int o = o0 + r;
        float dv = 1.f;
        if (demod && o < OC) {
            const float* Wo = W + (int64_t)o * IC * KK;
            float acc = 0.f;
            for (int i = lane; i < IC; i += 32) {
                float q = 0.f;
                for (int t = 0; t < KK; ++t) {
                    const float w0 = Wo[i * KK + t] * scale;
                    q = fmaf(w0, w0, q);
                }
                const float si = sb[i];
                acc = fmaf(si * si, q, acc);
            }
            acc = warp_sum(acc);
            dv = rsqrtf(acc + eps);
        }
        if (lane == 0) {
            dsm[r] = dv;
            if (d_out && o < OC) d_out[(int64_t)b * OC + o] = dv;
        }
    }
    __syncthreads();
    // the (tap, 32-channel input tile) pairs are split over gridDim.z: a shared (B = 1) 512 x 512 weight has only 16
    // output-channel tiles, far too few blocks for 148 SMs
    const int itiles = (IC + 31) / 32;
    for (int work = blockIdx.z; work < KK * itiles; work += gridDim.z) {
        const int t = work / itiles, i0 = (work - t * itiles) * 32;
        const int tt = flip ? KK - 1 - t : t;
        {
            const int i = i0 + lane;
            const float si = i < IC ? scale * sb[i] : 0.f;
#pragma unroll
            for (int r = warp; r < 32; r += 8) {
                const int o = o0 + r;
                float v = 0.f;
                if (o < OC && i < IC) {
                    v = W[((int64_t)o * IC + i) * KK + t] * si * dsm[r];
                    io<T>::st(wk + (((int64_t)b * KK + tt) * OC + o) * IC + i, v);
                }
                tile[r][lane] = v;
            }
            if (wkT) {
                __syncthreads();
#pragma unroll
                for (int r = warp; r < 32; r += 8) {         // r: input channel within the tile, lane: output channel
                    const int ii = i0 + r, o = o0 + lane;
                    if (ii < IC && o < OC)
                        io<T>::st(wkT + (((int64_t)b * KK + (KK - 1 - tt)) * IC + ii) * OC + o, tile[lane][r]);
                }
                __syncthreads();
            }
        }
    }
}

// grid (ceil(OC / 4), B), 128 threads: each warp owns output channels o = 4*blockIdx.x + warp.
// gs must be zero-initialised (atomic accumulation over the OC blocks); e is written.
__global__ void __launch_bounds__(128) modweight_bwd_s_kernel(const float* __restrict__ g, const float* __restrict__ W,
                                                              const float* __restrict__ s, const float* __restrict__ d,
                                                              float* __restrict__ gs, float* __restrict__ e_out, int B, int OC,
                                                              int IC, int KK, float scale, int demod, int flip) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 4 + warp, b = blockIdx.y;
    if (o >= OC) return;
    const float* Wo = W + (int64_t)o * IC * KK;
    const float* sb = s + (int64_t)b * IC;
    const float dv = d ? d[(int64_t)b * OC + o] : 1.f;
    // sweep 1: gd = sum_{i,t} g * w0 * s
    float gd = 0.f;
    if (demod) {
        for (int i = lane; i < IC; i += 32) {
            float a = 0.f;
            for (int t = 0; t < KK; ++t) {
                const int tt = flip ? KK - 1 - t : t;
                a = fmaf(g[(((int64_t)b * KK + tt) * OC + o) * IC + i], Wo[i * KK + t], a);
            }
            gd = fmaf(a * scale, sb[i], gd);
        }
        gd = warp_sum(gd);
    }
    const float e = demod ? -gd * dv * dv * dv : 0.f;
    if (lane == 0 && e_out) e_out[(int64_t)b * OC + o] = e;
    // sweep 2: this channel's contribution to gs[b][:]
    for (int i = lane; i < IC; i += 32) {
        float a = 0.f, q = 0.f;
        for (int t = 0; t < KK; ++t) {
            const int tt = flip ? KK - 1 - t : t;
            const float w0 = Wo[i * KK + t] * scale;
            a = fmaf(g[(((int64_t)b * KK + tt) * OC + o) * IC + i], w0, a);
            q = fmaf(w0, w0, q);
        }
        atomicAdd(gs + (int64_t)b * IC + i, a * dv + e * sb[i] * q);
    }
}

// grid (OC, ceil(IC / 128)), 128 threads: thread (o, i) sums over the batch.  gW[o][i][t] is WRITTEN (param layout).
__global__ void __launch_bounds__(128) modweight_bwd_w_kernel(const float* __restrict__ g, const float* __restrict__ W,
                                                              const float* __restrict__ s, const float* __restrict__ d,
                                                              const float* __restrict__ e, float* __restrict__ gW, int B, int OC,
                                                              int IC, int KK, float scale, int flip) {
    const int o = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= IC) return;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    float demod_sum = 0.f;                 // sum_b e[b][o] * s[b][i]^2
    for (int b = 0; b < B; ++b) {
        const float si = s[(int64_t)b * IC + i];
        const float f = si * (d ? d[(int64_t)b * OC + o] : 1.f);
        if (e) demod_sum = fmaf(e[(int64_t)b * OC + o], si * si, demod_sum);
#pragma unroll
        for (int t = 0; t < 9; ++t)
            if (t < KK) {
                const int tt = flip ? KK - 1 - t : t;
                acc[t] = fmaf(g[(((int64_t)b * KK + tt) * OC + o) * IC + i], f, acc[t]);
            }
    }
    const float* Wp = W + ((int64_t)o * IC + i) * KK;
    float* out = gW + ((int64_t)o * IC + i) * KK;
#pragma unroll
    for (int t = 0; t < 9; ++t)
        if (t < KK) out[t] = scale * (acc[t] + demod_sum * scale * Wp[t]);
}

}  // namespace b200gan

extern "C" int b200gan_modweight_fwd(const float* weight, const float* s, float* d, void* wk, void* wk_adjoint, int dtype,
                                     int b, int oc, int ic, int kh, int kw, float scale, int demodulate, int flip,
                                     void* stream) {
    using namespace b200gan;
    B200_REQUIRE(b >= 0 && oc >= 1 && ic >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 9, "modweight_fwd: bad shape (k*k <= 9)");
    B200_REQUIRE(b <= 65535, "modweight_fwd: batch too large");
    if (b == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        const int otiles = (oc + 31) / 32, work = kh * kw * ((ic + 31) / 32);
        int z = (int)cdiv((int64_t)3 * sm_count(), (int64_t)otiles * b);
        if (z > work) z = work;
        if (z < 1) z = 1;
        if (demodulate && z > 4) z = 4;          // every z slice recomputes the demodulation sums of its 32 channels
        modweight_fwd_kernel<T><<<dim3(otiles, b, z), 256, 0, (cudaStream_t)stream>>>(weight, s, d, (T*)wk, (T*)wk_adjoint, b, oc,
                                                                                     ic, kh * kw, scale, demodulate, flip, 1e-8f);
        count_launch();
        return check_launch("modweight_fwd");
    });
}

extern "C" int b200gan_modweight_bwd(const float* g, const float* weight, const float* s, const float* d, float* gs, float* e,
                                     float* gweight, int b, int oc, int ic, int kh, int kw, float scale, int demodulate,
                                     int flip, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(b >= 0 && oc >= 1 && ic >= 1 && kh * kw <= 9, "modweight_bwd: bad shape (k*k <= 9)");
    B200_REQUIRE(b <= 65535, "modweight_bwd: batch too large");
    B200_REQUIRE(!demodulate || (d != nullptr && e != nullptr), "modweight_bwd: demodulation needs d and the scratch row e");
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (gs) {
        modweight_bwd_s_kernel<<<dim3((oc + 3) / 4, b), 128, 0, st>>>(g, weight, s, d, gs, demodulate ? e : nullptr, b, oc, ic,
                                                                     kh * kw, scale, demodulate, flip);
        count_launch();
        if (int rc = check_launch("modweight_bwd_s")) return rc;
    }
    if (gweight) {
        B200_REQUIRE(!demodulate || gs != nullptr, "modweight_bwd: the weight gradient needs the style pass (e) first");
        modweight_bwd_w_kernel<<<dim3(oc, (ic + 127) / 128), 128, 0, st>>>(g, weight, s, d, demodulate ? e : nullptr, gweight, b,
                                                                           oc, ic, kh * kw, scale, flip);
        count_launch();
        if (int rc = check_launch("modweight_bwd_w")) return rc;
    }
    return 0;
}
