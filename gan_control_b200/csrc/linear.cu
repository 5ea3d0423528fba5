// Dense layers: EqualLinear forward (gan_model.py:189-197) and the small fp32 GEMMs its backward
// passes need.  These are skinny (M = batch <= 64) and weight-read bound; one tiled CUDA-core
// kernel with split-K covers them.  The whole-mapping-network persistent kernel lives in
// mapping.cu.
#include "common.cuh"

namespace b200gan {

constexpr int LM = 32, LN = 64, LK = 16;

template <typename TA, typename TC>
__global__ void __launch_bounds__(256) gemm_kernel(const TA* __restrict__ a, const float* __restrict__ b,
                                                   TC* __restrict__ c, int m, int n, int k, int lda,
                                                   int ldb, int ldc, int trans_a, int trans_b,
                                                   float alpha, float beta, const float* __restrict__ bias,
                                                   float bias_mul, int act, int k_chunk, int use_atomic) {
    __shared__ float As[LK][LM + 1];
    __shared__ float Bs[LK][LN + 1];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * LM, n0 = blockIdx.x * LN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(k, k_begin + k_chunk);
    const int tn = t & 31, tm = t >> 5;   // thread: rows tm*4..+3, cols tn*2..+1
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = k_begin; k0 < k_end; k0 += LK) {
        // A tile: LM x LK = 512 elements, 2 per thread
        for (int e = t; e < LM * LK; e += 256) {
            int mi, ki;
            if (trans_a) { mi = e % LM; ki = e / LM; } else { ki = e % LK; mi = e / LK; }
            int gm = m0 + mi, gk = k0 + ki;
            float v = 0.f;
            if (gm < m && gk < k_end) v = io<TA>::ld(trans_a ? a + (int64_t)gk * lda + gm : a + (int64_t)gm * lda + gk);
            As[ki][mi] = v;
        }
        // B tile: LK x LN = 1024 elements, 4 per thread
        for (int e = t; e < LK * LN; e += 256) {
            int ni, ki;
            if (trans_b) { ki = e % LK; ni = e / LK; } else { ni = e % LN; ki = e / LN; }
            int gn = n0 + ni, gk = k0 + ki;
            float v = 0.f;
            if (gn < n && gk < k_end) v = trans_b ? b[(int64_t)gn * ldb + gk] : b[(int64_t)gk * ldb + gn];
            Bs[ki][ni] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LK; ++kk) {
            float b0 = Bs[kk][tn * 2], b1 = Bs[kk][tn * 2 + 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float av = As[kk][tm * 4 + i];
                acc[i][0] = fmaf(av, b0, acc[i][0]);
                acc[i][1] = fmaf(av, b1, acc[i][1]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + tm * 4 + i;
        if (gm >= m) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int gn = n0 + tn * 2 + j;
            if (gn >= n) continue;
            TC* dst = c + (int64_t)gm * ldc + gn;
            float v = alpha * acc[i][j];
            if (use_atomic) {
                if constexpr (sizeof(TC) == 4) atomicAdd(dst, v);   // split-K: fp32 output only
            } else {
                if (beta != 0.f) v += beta * io<TC>::ld(dst);
                if (bias) v += bias[gn] * bias_mul;
                if (act) v = 1.4142135623730951f * (v > 0.f ? v : 0.2f * v);
                io<TC>::st(dst, v);
            }
        }
    }
}

// Skinny forward of EqualLinear (M = batch <= 64): one warp per output neuron streams its weight row
// once (coalesced), the activations come from L1/L2; bias + leaky-ReLU fused.  A 16x512x512 layer is
// 64 CTAs x ~2 us instead of 8 CTAs looping over K.
template <typename T>
__global__ void __launch_bounds__(256) linear_skinny_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, T* __restrict__ y, int m,
                                                            int n, int k, float scale, float bias_mul, int act) {
    constexpr int MB = 8;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < n; j += nwarps) {
        const float* wrow = w + (int64_t)j * k;
        const float bj = bias ? bias[j] * bias_mul : 0.f;
        for (int m0 = 0; m0 < m; m0 += MB) {
            float acc[MB];
#pragma unroll
            for (int i = 0; i < MB; ++i) acc[i] = 0.f;
            if (sizeof(T) == 4 && (k & 127) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0) {
                // fp32 activations: 16-byte loads, 4 k's per lane per step
                const float* xf = reinterpret_cast<const float*>(x);
                for (int kk = lane * 4; kk < k; kk += 128) {
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + kk);
#pragma unroll
                    for (int i = 0; i < MB; ++i)
                        if (m0 + i < m) {
                            const float4 xv = *reinterpret_cast<const float4*>(xf + (int64_t)(m0 + i) * k + kk);
                            acc[i] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[i]))));
                        }
                }
            } else {
                for (int kk = lane; kk < k; kk += 32) {
                    const float wv = wrow[kk];
#pragma unroll
                    for (int i = 0; i < MB; ++i)
                        if (m0 + i < m) acc[i] = fmaf(wv, io<T>::ld(x + (int64_t)(m0 + i) * k + kk), acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < MB; ++i)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            if (lane < MB && m0 + lane < m) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < MB; ++i)
                    if (i == lane) v = acc[i];
                v = v * scale + bj;
                if (act) v = 1.4142135623730951f * (v > 0.f ? v : 0.2f * v);
                io<T>::st(y + (int64_t)(m0 + lane) * n + j, v);
            }
        }
    }
}

template <typename TA, typename TC>
static int launch_gemm(const TA* a, const float* b, TC* c, int m, int n, int k, int lda, int ldb, int ldc,
                       int trans_a, int trans_b, float alpha, float beta, const float* bias, float bias_mul,
                       int act, cudaStream_t st) {
    if (m == 0 || n == 0) return 0;
    int tiles = (int)(cdiv(m, LM) * cdiv(n, LN));
    int splits = 1;
    const bool can_split = sizeof(TC) == 4 && bias == nullptr && act == 0 && (beta == 0.f || beta == 1.f);
    if (can_split && k >= 256 && tiles < sm_count() / 2) {
        splits = (int)cdiv(sm_count(), tiles);
        int max_splits = k / 64;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int k_chunk = (int)(cdiv(cdiv(k, splits), LK) * LK);
    splits = (int)cdiv(k, k_chunk);
    if (splits > 1 && beta == 0.f) {
        if (ldc == n) {
            cudaMemsetAsync(c, 0, sizeof(float) * (size_t)m * n, st);
        } else {
            splits = 1;
            k_chunk = k;
        }
    }
    dim3 grid((unsigned)cdiv(n, LN), (unsigned)cdiv(m, LM), (unsigned)splits);
    gemm_kernel<TA, TC><<<grid, 256, 0, st>>>(a, b, c, m, n, k, lda, ldb, ldc, trans_a, trans_b, alpha, beta,
                                              bias, bias_mul, act, k_chunk, splits > 1 ? 1 : 0);
    count_launch();
    return check_launch("gemm");
}

// ---- Adam (+ EMA) -------------------------------------------------------------------------

}  // namespace b200gan

extern "C" int b200gan_linear_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int m,
                                  int n, int k, float scale, float bias_mul, int act, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(m >= 0 && n >= 1 && k >= 1, "linear_fwd: bad shape");
    return B200_DISPATCH(dtype, [&] {
        if (m <= 64) {
            if (m == 0) return 0;
            int blocks = (int)cdiv(n, 8);
            int cap = sm_count() * 4;
            if (blocks > cap) blocks = cap;
            linear_skinny_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, w, bias, (T*)y, m, n, k, scale,
                                                                             bias_mul, act);
            count_launch();
            return check_launch("linear_skinny");
        }
        return launch_gemm<T, T>((const T*)x, w, (T*)y, m, n, k, k, k, n, 0, 1, scale, 0.f, bias, bias_mul, act,
                                 (cudaStream_t)stream);
    });
}

extern "C" int b200gan_gemm_f32(const float* a, const float* b, float* c, int m, int n, int k, int lda, int ldb,
                                int ldc, int trans_a, int trans_b, float alpha, float beta, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm_f32: bad shape");
    return launch_gemm<float, float>(a, b, c, m, n, k, lda, ldb, ldc, trans_a, trans_b, alpha, beta, nullptr, 0.f,
                                     0, (cudaStream_t)stream);
}
