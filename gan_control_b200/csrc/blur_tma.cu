// Blur (upfirdn2d with up = down = 1, <= 4x4 taps) as a TMA-fed shared-memory stencil, bf16 NHWC.
//
// The register-only kernel (upfirdn2d.cu::blur_rows_kernel) is latency-bound: ncu shows 64 % of the
// stall samples on long_scoreboard at 26 % of HBM bandwidth (profiles/r01_ncu_kernels.md) -- every thread
// waits for its own loads.  Here the loads are decoupled from the threads: an elected thread streams
// (16+kh-1) x (TW+kw-1)-pixel input tiles through a double-buffered TMA pipeline (out-of-bounds =
// the zero padding), 128 threads run the separable FIR out of shared memory with rolling row
// accumulators (swizzle-aware, conflict-free 16-byte reads) and write 16-byte vectors.
#include <cstring>

#include "common.cuh"
#include "fir_epilogue.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

constexpr int kBlurThreads = 128;
constexpr int kBlurRows = 8;           // output rows per tile (26 KB stages -> 4 CTAs per SM)
constexpr int kBlurStages = 2;

struct BlurParams {
    int n, in_h, in_w, c, out_h, out_w, kh, kw, pad0_y, pad0_x, flip;
    float gain;
    int tiles_x, tiles_y, chunks, total_tiles;
    FastDiv div_chunks, div_tx, div_ty;      // tile decode without integer division (every thread decodes every tile)
    const float* taps;
    __nv_bfloat16* y;
    FirEpilogue ep;
};

// CH = channels per tile: 64 (128-byte pixel rows, 16 output columns) or 32 (64-byte rows, 32 columns); either way
// 128 threads = (output column, 16-byte channel group).  The tile is stored UNswizzled: thread t's vector of pixel
// (row r, column tx + kx) sits at  t*16 + r*PITCH + kx*ROWB, so a warp reads 512 contiguous bytes (conflict-free)
// and every address in the unrolled row loop is an immediate offset from one per-thread base.  ncu on the first
// version of this kernel (swizzled tile, run-time strides): issue-bound, 72 % issue-slot utilisation at 41 % of HBM,
// 250 warp instructions per 8-channel output of which 64 were FMAs (profiles/r01_ncu_kernels.md).  Here the
// address arithmetic is gone, the FMAs are packed (fma.rn.f32x2), and the rolling accumulators are (re)started by
// the first tap of a row instead of being zeroed.
template <int CH>
__global__ void __launch_bounds__(kBlurThreads) blur_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                const __grid_constant__ BlurParams p) {
    constexpr int NCG = CH / 8, TW = kBlurThreads / NCG, BOXW = TW + 3, BOXH = kBlurRows + 3, ROWB = CH * 2;
    constexpr int PITCH = BOXW * ROWB;
    constexpr int STAGE = (PITCH * BOXH + 1023) & ~1023;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t full[kBlurStages];
    __shared__ float taps[16], tap_v[4], tap_h[4];
    __shared__ int separable;
    const int tid = threadIdx.x;
    if (tid < 16) {
        const int ky = tid / 4, kx = tid % 4;
        float v = 0.f;
        if (ky < p.kh && kx < p.kw) {
            const int sy = p.flip ? p.kh - 1 - ky : ky, sx = p.flip ? p.kw - 1 - kx : kx;
            v = p.taps[sy * p.kw + sx] * p.gain;
        }
        taps[tid] = v;
    }
    if (tid == 0) {
        for (int s = 0; s < kBlurStages; ++s) mbar_init(full + s, 1);
        fence_barrier_init();
        prefetch_tensormap(&map_x);
    }
    __syncthreads();
    if (tid == 0) {
        const float f00 = taps[0];
        int sep = fabsf(f00) > 1e-20f;
        for (int ky = 0; ky < 4 && sep; ++ky)
            for (int kx = 0; kx < 4; ++kx)
                if (fabsf(taps[ky * 4] * taps[kx] / f00 - taps[ky * 4 + kx]) > 1e-6f * fabsf(f00)) sep = 0;
        separable = sep;
        for (int k = 0; k < 4; ++k) {
            tap_v[k] = sep ? taps[k * 4] / f00 : 0.f;
            tap_h[k] = sep ? taps[k] : 0.f;
        }
    }
    __syncthreads();
    const bool sep = separable != 0;
    const int cg = tid % NCG, tx = tid / NCG;
    float2 th[4], tv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        th[k] = make_float2(tap_h[k], tap_h[k]);
        tv[k] = make_float2(tap_v[k], tap_v[k]);
    }

    auto issue = [&](int tile, int stage) {
        uint32_t t, uchunk, ubx, uby, ub;
        p.div_chunks.divmod((uint32_t)tile, t, uchunk);
        p.div_tx.divmod(t, ub, ubx);
        p.div_ty.divmod(ub, t, uby);
        const int chunk = (int)uchunk, bx = (int)ubx, by = (int)uby, b = (int)t;
        mbar_arrive_expect_tx(full + stage, (uint32_t)(PITCH * BOXH));
        tma_load_4d(smem + stage * STAGE, &map_x, full + stage, chunk * CH, bx * TW - p.pad0_x, by * kBlurRows - p.pad0_y, b);
    };

    int it = 0;
    if (tid == 0 && (int)blockIdx.x < p.total_tiles) issue(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < p.total_tiles) issue(next, stage ^ 1);      // stage^1 was released by the barrier below
        mbar_wait(full + stage, (it >> 1) & 1);
        uint32_t t, uchunk, ubx, uby, ub;
        p.div_chunks.divmod((uint32_t)tile, t, uchunk);
        p.div_tx.divmod(t, ub, ubx);
        p.div_ty.divmod(ub, t, uby);
        const int chunk = (int)uchunk, bx = (int)ubx, by = (int)uby, b = (int)t;
        const uint8_t* tbase = smem + stage * STAGE + tid * 16;
        const int ox = bx * TW + tx, oy0 = by * kBlurRows;
        float2 acc[4][4];                                   // rolling output rows x channel pairs
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][j] = make_float2(0.f, 0.f);
        __nv_bfloat16* yb = p.y + (((int64_t)b * p.out_h) * p.out_w + ox) * p.c + chunk * CH + cg * 8;
        // taps are zero-padded to 4x4, so every tile reads kBlurRows + 3 input rows.  Input row r feeds output rows
        // r-3..r; the tap-0 contribution ASSIGNS the accumulator slot (so slots need no zeroing and contributions to
        // rows outside the tile are harmless), the tap-3 contribution completes row r-3.  The row loop is rolled in
        // groups of 4 so that the slot indices stay compile-time and the body stays inside the instruction cache.
#pragma unroll 1
        for (int r0 = 0; r0 < BOXH; r0 += 4) {
            const uint8_t* rb = tbase + r0 * PITCH;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = r0 + rr;
                if (r >= BOXH) break;
                float2 in[4][4];
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rb + rr * PITCH + kx * ROWB);
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        in[kx][j] = make_float2(__uint_as_float(w4[j] << 16), __uint_as_float(w4[j] & 0xffff0000u));
                }
                if (sep) {
                    float2 h[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h[j] = __fmul2_rn(th[0], in[0][j]);
                        h[j] = __ffma2_rn(th[1], in[1][j], h[j]);
                        h[j] = __ffma2_rn(th[2], in[2][j], h[j]);
                        h[j] = __ffma2_rn(th[3], in[3][j], h[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[rr & 3][j] = __fmul2_rn(tv[0], h[j]);
                        acc[(rr - 1) & 3][j] = __ffma2_rn(tv[1], h[j], acc[(rr - 1) & 3][j]);
                        acc[(rr - 2) & 3][j] = __ffma2_rn(tv[2], h[j], acc[(rr - 2) & 3][j]);
                        acc[(rr - 3) & 3][j] = __ffma2_rn(tv[3], h[j], acc[(rr - 3) & 3][j]);
                    }
                } else {
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx) {
                            const float f = taps[ky * 4 + kx];
                            const float2 f2 = make_float2(f, f);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (ky == 0 && kx == 0) acc[rr & 3][j] = __fmul2_rn(f2, in[kx][j]);
                                else acc[(rr - ky) & 3][j] = __ffma2_rn(f2, in[kx][j], acc[(rr - ky) & 3][j]);
                            }
                        }
                }
                if (r >= 3) {                                  // output row r - 3 is complete
                    const int orow = r - 3;
                    if (oy0 + orow < p.out_h && ox < p.out_w) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { o[2 * j] = acc[(rr - 3) & 3][j].x; o[2 * j + 1] = acc[(rr - 3) & 3][j].y; }
                        if (p.ep.enabled)
                            fir_epilogue<__nv_bfloat16, 8>(p.ep, o, b, ((int64_t)b * p.out_h + oy0 + orow) * p.out_w + ox, p.c,
                                                           chunk * CH + cg * 8);
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
                            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        *reinterpret_cast<uint4*>(yb + (int64_t)(oy0 + orow) * p.out_w * p.c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        }
        __syncthreads();          // everyone is done reading `stage`: it may be refilled two iterations on
    }
}

bool blur_tma_eligible(int dtype, int c, int kh, int kw, int up, int down, int out_h, int out_w, const void* x, const void* y) {
    if (dtype != B200GAN_BF16 || up != 1 || down != 1 || kh > 4 || kw > 4) return false;
    if (c % 32 != 0) return false;
    if (out_h < 8 || out_w < 16) return false;
    if (((uintptr_t)x | (uintptr_t)y) % 16 != 0) return false;
    return tensor_map_encoder() != nullptr;
}

template <int CH>
static int blur_tma_launch(const CUtensorMap& map_x, const BlurParams& p, cudaStream_t st) {
    constexpr int NCG = CH / 8, TW = kBlurThreads / NCG, BOXW = TW + 3, BOXH = kBlurRows + 3;
    constexpr int STAGE = (BOXW * CH * 2 * BOXH + 1023) & ~1023;
    const size_t smem = 1024 + (size_t)kBlurStages * STAGE;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(blur_tma_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_dev = cur_dev;
    }
    int per_sm = (int)((220 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = sm_count() * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    blur_tma_kernel<CH><<<grid, kBlurThreads, smem, st>>>(map_x, p);
    count_launch();
    return check_launch("blur_tma");
}

int blur_tma(const void* x, void* y, const float* taps, int n, int in_h, int in_w, int c, int out_h, int out_w, int kh,
             int kw, int pad0_y, int pad0_x, int flip, float gain, const FirEpilogue& ep, cudaStream_t st) {
    BlurParams p;
    memset(&p, 0, sizeof(p));
    p.ep = ep;
    p.n = n; p.in_h = in_h; p.in_w = in_w; p.c = c; p.out_h = out_h; p.out_w = out_w; p.kh = kh; p.kw = kw;
    p.pad0_y = pad0_y; p.pad0_x = pad0_x; p.flip = flip; p.gain = gain; p.taps = taps; p.y = (__nv_bfloat16*)y;
    const int ch = c % 64 == 0 ? 64 : 32;
    const int tw = kBlurThreads / (ch / 8);
    p.tiles_x = (out_w + tw - 1) / tw;
    p.tiles_y = (out_h + kBlurRows - 1) / kBlurRows;
    p.chunks = c / ch;
    p.total_tiles = n * p.tiles_y * p.tiles_x * p.chunks;
    p.div_chunks = make_fastdiv((uint32_t)p.chunks);
    p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
    p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
    if (p.total_tiles == 0) return 0;
    CUtensorMap map_x;
    uint64_t dims[4] = {(uint64_t)c, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)c * 2, (uint64_t)in_w * c * 2, (uint64_t)in_h * in_w * c * 2};
    uint32_t box[4] = {(uint32_t)ch, (uint32_t)(tw + 3), (uint32_t)(kBlurRows + 3), 1};   // taps are zero-padded to 4x4
    uint32_t es[4] = {1, 1, 1, 1};
    if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, 0)) return e;      // no swizzle (see kernel)
    return ch == 64 ? blur_tma_launch<64>(map_x, p, st) : blur_tma_launch<32>(map_x, p, st);
}

}  // namespace b200gan
