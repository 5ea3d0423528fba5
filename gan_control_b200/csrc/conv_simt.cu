// General gather-form convolution and its weight gradient on CUDA cores (fp32 accumulate).
//
// This is the universal path: any (up, down, pad, kernel size, channel count, per-sample or
// shared weights), fp32 or bf16 storage.  It is the exact-fp32 "parity" engine and the fallback
// for shapes the tcgen05 implicit-GEMM kernels (conv_umma.cu) do not take (channel counts that
// are not multiples of 16, tiny images).  Both engines implement the same contract:
//
//   y[b][oy][ox][o] = sum_{ky,kx,i} z[b][oy*down+ky-pad0][ox*down+kx-pad0][i] * w[wb][ky][kx][o][i]
//
// (include/b200gan.h).  Reference computations covered: F.conv2d / F.conv_transpose2d with
// groups=batch in ModulatedConv2d.forward (gan_model.py:295-329), EqualConv2d (gm.py:152-160).
#include "conv.cuh"

namespace b200gan {

constexpr int BM = 64, BN = 64, BK = 16;

// map output pixel + tap -> input pixel; returns false when the tap reads a zero
__device__ __forceinline__ bool tap_src(const ConvGeom& g, int oy, int ox, int ky, int kx, int& iy, int& ix) {
    int zy = oy * g.down + ky - g.pad0, zx = ox * g.down + kx - g.pad0;
    if (zy < 0 || zx < 0) return false;
    if (g.up > 1) {
        if (zy % g.up != 0 || zx % g.up != 0) return false;
        zy /= g.up;
        zx /= g.up;
    }
    iy = zy;
    ix = zx;
    return iy < g.in_h && ix < g.in_w;
}

// element offsets of logical (b, row, col, channel) in the physical NHWC tensor (see ConvGeom::pack_*)
__device__ __forceinline__ int64_t in_offset(const ConvGeom& g, int b, int iy, int ix, int k) {
    if (!g.pack_in) return (((int64_t)b * g.in_h + iy) * g.in_w + ix) * g.ic + k;
    const int cp = g.ic >> 2, q = k / cp, c = k - q * cp;
    return (((int64_t)b * (2 * g.in_h) + 2 * iy + (q >> 1)) * (2 * g.in_w) + 2 * ix + (q & 1)) * cp + c;
}
// output: offset and (physical) pixel index / channel for the epilogue's side inputs
__device__ __forceinline__ int64_t out_offset(const ConvGeom& g, int b, int oy, int ox, int o, int64_t& ppix, int& pch) {
    if (!g.pack_out) {
        ppix = ((int64_t)b * g.out_h + oy) * g.out_w + ox;
        pch = o;
        return ppix * g.oc + o;
    }
    const int cq = g.oc >> 2, q = o / cq;
    pch = o - q * cq;
    ppix = ((int64_t)b * (2 * g.out_h) + 2 * oy + (q >> 1)) * (2 * g.out_w) + 2 * ox + (q & 1);
    return ppix * cq + pch;
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, int valid, bool vec, float out[4]) {
    if (vec && valid >= 4) {
        Pack<T, 4> v = *reinterpret_cast<const Pack<T, 4>*>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = io<T>::ld(&v.v[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = j < valid ? io<T>::ld(p + j) : 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) conv_fwd_simt_kernel(
    const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, ConvGeom g, ConvEp ep, int vec_in, int vec_out) {
    const float* __restrict__ bias = ep.bias;
    const float* __restrict__ rowscale = ep.rowscale;
    const T* __restrict__ noise = (const T*)ep.noise;
    const float* __restrict__ noise_w = ep.noise_w;
    const T* __restrict__ addend = (const T*)ep.addend;
    const T* __restrict__ gate = (const T*)ep.gate;
    const float slope = ep.slope, gain = ep.gain;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int npix = g.out_h * g.out_w;
    const int m0 = blockIdx.x * BM;      // first pixel of this tile (within image)
    const int n0 = blockIdx.y * BN;      // first output channel
    const int b = blockIdx.z;
    const int wb = g.w_per_sample ? b : 0;
    // loader roles: 64 rows x 4 quads of 4 channels
    const int lrow = t >> 2, lq = (t & 3) * 4;
    const int lpix = m0 + lrow;
    const int loy = lpix / g.out_w, lox = lpix % g.out_w;
    // compute roles
    const int tn = t & 15, tm = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = g.kh * g.kw;
    for (int tap = 0; tap < taps; ++tap) {
        const int ky = tap / g.kw, kx = tap % g.kw;
        int iy = 0, ix = 0;
        const bool a_ok = lpix < npix && tap_src(g, loy, lox, ky, kx, iy, ix);
        const int oc_l = n0 + lrow;
        const T* b_ptr = w + (((int64_t)wb * taps + tap) * g.oc + oc_l) * g.ic;
        for (int k0 = 0; k0 < g.ic; k0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
            const int kk = k0 + lq;
            const int valid = g.ic - kk;
            if (a_ok && valid > 0) load4<T>(x + in_offset(g, b, iy, ix, kk), valid, vec_in, av);
            if (oc_l < g.oc && valid > 0) load4<T>(b_ptr + kk, valid, vec_in, bv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                As[lq + j][lrow] = av[j];
                Bs[lq + j][lrow] = bv[j];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // epilogue
    const float nw = (noise != nullptr && noise_w != nullptr) ? *noise_w : 0.f;
    const bool has_ep = bias || rowscale || noise || addend || gate || slope != 1.f || gain != 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pix = m0 + tm * 4 + i;
        if (pix >= npix) continue;
        int64_t ppix;
        int pch0;
        const int oc_phys = g.pack_out ? g.oc >> 2 : g.oc;
        const int o0 = n0 + tn * 4 < g.oc ? n0 + tn * 4 : 0;                  // 4 consecutive channels share a phase
        const int64_t doff = out_offset(g, b, pix / g.out_w, pix % g.out_w, o0, ppix, pch0);
        T* dst = y + doff;
        const float nz = noise ? nw * io<T>::ld(noise + ppix) : 0.f;
        Pack<T, 4> o;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o_ch = n0 + tn * 4 + j;
            float v = acc[i][j];
            if (has_ep && o_ch < g.oc) {
                if (addend) v += io<T>::ld(addend + doff + j);
                if (rowscale) v *= rowscale[(int64_t)b * oc_phys + pch0 + j];
                if (gate) {
                    v *= io<T>::ld(gate + doff + j) > 0.f ? gain : gain * slope;
                } else {
                    v += nz + (bias ? bias[pch0 + j] : 0.f);
                    v = gain * (v > 0.f ? v : v * slope);
                }
            }
            io<T>::st(&o.v[j], v);
        }
        if (vec_out && n0 + tn * 4 + 3 < g.oc) {
            *reinterpret_cast<Pack<T, 4>*>(dst) = o;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + tn * 4 + j < g.oc) dst[j] = o.v[j];
        }
    }
}

// gw[wb][tap][o][i] += sum over this CTA's pixel chunk of gy[pix][o] * z[pix@tap][i]
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* __restrict__ x,
                                                              const T* __restrict__ gy,
                                                              float* __restrict__ gw, ConvGeom g,
                                                              int n_tiles, int chunk, int chunks_per_img,
                                                              int vec_ic, int vec_oc) {
    __shared__ float As[BK][BM + 4];   // [pixel][oc]
    __shared__ float Bs[BK][BN + 4];   // [pixel][ic]
    const int t = threadIdx.x;
    const int tile = blockIdx.x;
    const int m0 = (tile / n_tiles) * BM;   // oc tile
    const int n0 = (tile % n_tiles) * BN;   // ic tile
    const int tap = blockIdx.y;
    const int ky = tap / g.kw, kx = tap % g.kw;
    const int b = blockIdx.z / chunks_per_img;
    const int npix = g.out_h * g.out_w;
    const int p_begin = (blockIdx.z % chunks_per_img) * chunk;
    const int p_end = min(npix, p_begin + chunk);
    const int wb = g.w_per_sample ? b : 0;
    // loader: 16 pixels x 16 quads
    const int lp = t >> 4, lq = (t & 15) * 4;
    const int tn = t & 15, tm = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int p0 = p_begin; p0 < p_end; p0 += BK) {
        const int pix = p0 + lp;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < p_end) {
            const int oy = pix / g.out_w, ox = pix % g.out_w;
            const int vo = g.oc - (m0 + lq);
            int64_t ppix;
            int pch;
            if (vo > 0) load4<T>(gy + out_offset(g, b, oy, ox, m0 + lq, ppix, pch), vo, vec_oc, av);
            int iy, ix;
            const int vi = g.ic - (n0 + lq);
            if (vi > 0 && tap_src(g, oy, ox, ky, kx, iy, ix))
                load4<T>(x + in_offset(g, b, iy, ix, n0 + lq), vi, vec_ic, bv);
        }
        *reinterpret_cast<float4*>(&As[lp][lq]) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4*>(&Bs[lp][lq]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int taps = g.kh * g.kw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o_ch = m0 + tm * 4 + i;
        if (o_ch >= g.oc) continue;
        float* dst = gw + (((int64_t)wb * taps + tap) * g.oc + o_ch) * g.ic + n0 + tn * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n0 + tn * 4 + j < g.ic) atomicAdd(dst + j, acc[i][j]);
    }
}

static int check_geom(const ConvGeom& g, const char* who) {
    B200_REQUIRE(g.b >= 0 && g.in_h >= 1 && g.in_w >= 1 && g.ic >= 1 && g.out_h >= 1 && g.out_w >= 1 && g.oc >= 1,
                 "%s: bad shape", who);
    B200_REQUIRE(g.kh >= 1 && g.kw >= 1 && g.kh <= 16 && g.kw <= 16, "%s: bad kernel size", who);
    B200_REQUIRE(g.up >= 1 && g.down >= 1, "%s: up/down must be >= 1", who);
    B200_REQUIRE(g.b <= 65535, "%s: batch too large", who);
    if (g.pack_in || g.pack_out) {
        B200_REQUIRE(g.up == 1 && g.down == 1, "%s: packed views need up = down = 1", who);
        B200_REQUIRE(!g.pack_in || g.ic % 16 == 0, "%s: pack_in needs ic %% 16 == 0", who);
        B200_REQUIRE(!g.pack_out || g.oc % 16 == 0, "%s: pack_out needs oc %% 16 == 0", who);
    }
    return 0;
}

int conv_fwd_simt(const void* x, const void* w, void* y, int dtype, const ConvGeom& g, const ConvEp& ep, cudaStream_t st) {
    if (int e = check_geom(g, "conv_fwd")) return e;
    if (g.b == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        const int npix = g.out_h * g.out_w;
        dim3 grid((unsigned)cdiv(npix, BM), (unsigned)cdiv(g.oc, BN), (unsigned)g.b);
        const size_t a4 = 4 * sizeof(T);
        int vec_in = g.ic % 4 == 0 && (uintptr_t)x % a4 == 0 && (uintptr_t)w % a4 == 0;
        int vec_out = g.oc % 4 == 0 && (uintptr_t)y % a4 == 0;
        conv_fwd_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, (T*)y, g, ep, vec_in, vec_out);
        count_launch();
        return check_launch("conv_fwd_simt");
    });
}

int conv_wgrad_simt(const void* x, const void* gy, float* gw, int dtype, const ConvGeom& g, cudaStream_t st) {
    if (int e = check_geom(g, "conv_wgrad")) return e;
    if (g.b == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        const int npix = g.out_h * g.out_w;
        const int m_tiles = (int)cdiv(g.oc, BM), n_tiles = (int)cdiv(g.ic, BN);
        const int taps = g.kh * g.kw;
        // split the pixel (K) dimension so that the grid covers the machine a few times over
        int64_t base = (int64_t)m_tiles * n_tiles * taps * g.b;
        int64_t want = cdiv((int64_t)sm_count() * 6, base);
        int chunk = (int)cdiv(npix, want < 1 ? 1 : want);
        chunk = (int)(cdiv(chunk, BK) * BK);
        if (chunk < 4 * BK) chunk = 4 * BK;
        int chunks_per_img = (int)cdiv(npix, chunk);
        B200_REQUIRE((int64_t)g.b * chunks_per_img <= 65535 && taps <= 65535, "conv_wgrad: grid too large");
        dim3 grid((unsigned)(m_tiles * n_tiles), (unsigned)taps, (unsigned)(g.b * chunks_per_img));
        const size_t a4 = 4 * sizeof(T);
        int vec_ic = g.ic % 4 == 0 && (uintptr_t)x % a4 == 0;
        int vec_oc = g.oc % 4 == 0 && (uintptr_t)gy % a4 == 0;
        conv_wgrad_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)gy, gw, g, n_tiles, chunk,
                                                        chunks_per_img, vec_ic, vec_oc);
        count_launch();
        return check_launch("conv_wgrad_simt");
    });
}

}  // namespace b200gan
