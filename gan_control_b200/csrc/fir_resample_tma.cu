// upfirdn2d with down = 2 (decimating FIR) and with up = 2 (interpolating FIR), <= 4x4 taps, bf16 NHWC, as
// TMA-fed shared-memory stencils.  Call sites: ResBlock.skip = Blur -> 1x1 stride-2 conv (gm.py:907-909), which only
// ever reads the blur at even positions -> one decimating FIR that writes a quarter of the pixels; and its adjoint
// (zero-insert x2 + reversed taps).  The generic register kernel (upfirdn2d.cu) gathers 16 taps per
// output straight from global memory and is latency-bound (measured 1.0 ms / 1.45 ms at 16x1024^2x32 channels, 21 % /
// 14 % of HBM).  Here, as in blur_tma.cu, an elected thread streams input tiles through a TMA pipeline
// (out-of-bounds = the zero padding), 128 threads = (output column, 16-byte channel group) run the FIR out of an
// UNswizzled tile (a warp reads contiguous bytes, every address an immediate offset), packed fp32 FMAs.
#include <cstring>

#include "common.cuh"
#include "fir_epilogue.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

constexpr int kFirThreads = 128;

struct FirResampleParams {
    int n, in_h, in_w, c, out_h, out_w, kh, kw, pad0_y, pad0_x, flip;
    float gain;
    int tiles_x, tiles_y, chunks, total_tiles;
    FastDiv div_chunks, div_tx, div_ty;      // tile decode without integer division
    const float* taps;
    __nv_bfloat16* y;
};

__device__ __forceinline__ void unpack8(const uint8_t* src, float2 (&v)[4]) {
    const uint4 q = *reinterpret_cast<const uint4*>(src);
    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = make_float2(__uint_as_float(w4[j] << 16), __uint_as_float(w4[j] & 0xffff0000u));
}

__device__ __forceinline__ void store8(__nv_bfloat16* dst, const float2 (&v)[4]) {
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j].x, v[j].y);
        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// taps -> shared 4x4 (zero-padded, flipped when asked, gain folded in)
__device__ __forceinline__ void load_taps(float* taps, const FirResampleParams& p) {
    if (threadIdx.x < 16) {
        const int ky = threadIdx.x / 4, kx = threadIdx.x % 4;
        float v = 0.f;
        if (ky < p.kh && kx < p.kw) {
            const int sy = p.flip ? p.kh - 1 - ky : ky, sx = p.flip ? p.kw - 1 - kx : kx;
            v = p.taps[sy * p.kw + sx] * p.gain;
        }
        taps[threadIdx.x] = v;
    }
}

__device__ __forceinline__ void decode_tile(const FirResampleParams& p, int tile, int& chunk, int& bx, int& by, int& b) {
    uint32_t t, uchunk, ubx, uby, ub;
    p.div_chunks.divmod((uint32_t)tile, t, uchunk);
    p.div_tx.divmod(t, ub, ubx);
    p.div_ty.divmod(ub, t, uby);
    chunk = (int)uchunk; bx = (int)ubx; by = (int)uby; b = (int)t;
}

// ------------------------------------------------------------------------------------------------------------------
// down = 2:  y[oy][ox] = sum_{ky,kx} f[ky][kx] * x[2*oy + ky - pad0][2*ox + kx - pad0]
// Tile = R output rows x TW output columns; input box (2R+2) x (2TW+2).  Input row 2m+a (a = 0,1) feeds output row m
// through tap row a and output row m-1 through tap row a+2: two rolling accumulators, (re)started by tap row 0.
// ------------------------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(kFirThreads) fir_down2_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ FirResampleParams p) {
    constexpr int NCG = CH / 8, TW = kFirThreads / NCG, R = 4, BOXW = 2 * TW + 2, BOXH = 2 * R + 2, ROWB = CH * 2;
    constexpr int PITCH = BOXW * ROWB, STAGE = (PITCH * BOXH + 1023) & ~1023, S = 2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t full[S];
    __shared__ float taps[16];
    const int tid = threadIdx.x;
    load_taps(taps, p);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        fence_barrier_init();
        prefetch_tensormap(&map_x);
    }
    __syncthreads();
    float2 f[4][4];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k / 4][k % 4] = make_float2(taps[k], taps[k]);
    const int cg = tid % NCG, tx = tid / NCG;

    auto issue = [&](int tile, int stage) {
        int chunk, bx, by, b;
        decode_tile(p, tile, chunk, bx, by, b);
        mbar_arrive_expect_tx(full + stage, (uint32_t)(PITCH * BOXH));
        tma_load_4d(smem + stage * STAGE, &map_x, full + stage, chunk * CH, 2 * bx * TW - p.pad0_x, 2 * by * R - p.pad0_y, b);
    };
    if (tid == 0)
        for (int s = 0; s < S - 1; ++s)
            if ((int)(blockIdx.x + s * gridDim.x) < p.total_tiles) issue(blockIdx.x + s * gridDim.x, s);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int stage = it % S;
        const int next = tile + (S - 1) * gridDim.x;
        if (tid == 0 && next < p.total_tiles) issue(next, (it + S - 1) % S);     // released by the barrier of iteration it-1
        mbar_wait(full + stage, (it / S) & 1);
        int chunk, bx, by, b;
        decode_tile(p, tile, chunk, bx, by, b);
        const uint8_t* tb = smem + stage * STAGE + 2 * tx * ROWB + cg * 16;
        const int ox = bx * TW + tx, oy0 = by * R;
        __nv_bfloat16* yb = p.y + (((int64_t)b * p.out_h) * p.out_w + ox) * p.c + chunk * CH + cg * 8;
        float2 acc[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m <= R; ++m) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                float2 in[4][4];
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) unpack8(tb + (2 * m + a) * PITCH + kx * ROWB, in[kx]);
#pragma unroll
                for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (m < R) {
                            if (a == 0 && kx == 0) acc[m & 1][j] = __fmul2_rn(f[0][0], in[0][j]);
                            else acc[m & 1][j] = __ffma2_rn(f[a][kx], in[kx][j], acc[m & 1][j]);
                        }
                        if (m >= 1) acc[(m - 1) & 1][j] = __ffma2_rn(f[a + 2][kx], in[kx][j], acc[(m - 1) & 1][j]);
                    }
            }
            if (m >= 1) {                                       // output row m-1 is complete
                const int oy = oy0 + m - 1;
                if (oy < p.out_h && ox < p.out_w) store8(yb + (int64_t)oy * p.out_w * p.c, acc[(m - 1) & 1]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// up = 2:  y[oy][ox] = sum_{ky,kx} f[ky][kx] * z[oy + ky - pad0][ox + kx - pad0],  z[2i][2j] = x[i][j], 0 elsewhere.
// Only taps with (oy + ky - pad0) even land on a sample: 2 x 2 of the 4 x 4 per output pixel.  Tile = 8 output rows x
// TW output columns (origin even), input box 6 x (TW/2 + 2).  The column parity (which tap columns a thread uses) is
// per thread, the row parity pattern is uniform (Q = pad0 & 1) and compiled both ways.
// ------------------------------------------------------------------------------------------------------------------
template <int CH, int Q>
__device__ __forceinline__ void fir_up2_rows(const uint8_t* tb, const float* taps, int kx0, __nv_bfloat16* yb, int oy0,
                                             bool col_ok, const FirResampleParams& p) {
    constexpr int NCG = CH / 8, TW = kFirThreads / NCG, BOXW = TW / 2 + 2, ROWB = CH * 2, PITCH = BOXW * ROWB;
    float2 fa[4], fb[4];                                        // taps of this thread's two columns, per tap row
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        fa[ky] = make_float2(taps[ky * 4 + kx0], taps[ky * 4 + kx0]);
        fb[ky] = make_float2(taps[ky * 4 + kx0 + 2], taps[ky * 4 + kx0 + 2]);
    }
    // output row o uses tap rows ky0 = Q ^ (o & 1) and ky0 + 2 on local input rows lr0 = (o + ky0 + Q) / 2 and lr0 + 1
    float2 lo[2][4], hi[2][4];
    constexpr int first = (0 + Q + Q) / 2;
    unpack8(tb + first * PITCH, lo[0]);
    unpack8(tb + first * PITCH + ROWB, lo[1]);
    unpack8(tb + (first + 1) * PITCH, hi[0]);
    unpack8(tb + (first + 1) * PITCH + ROWB, hi[1]);
    int cur = first;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int ky0 = Q ^ (o & 1);
        const int lr0 = (o + ky0 + Q) / 2;
        if (lr0 != cur) {                                       // compile-time after unrolling: slide the two-row window
#pragma unroll
            for (int j = 0; j < 4; ++j) { lo[0][j] = hi[0][j]; lo[1][j] = hi[1][j]; }
            unpack8(tb + (lr0 + 1) * PITCH, hi[0]);
            unpack8(tb + (lr0 + 1) * PITCH + ROWB, hi[1]);
            cur = lr0;
        }
        float2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j] = __fmul2_rn(fa[ky0], lo[0][j]);
            acc[j] = __ffma2_rn(fb[ky0], lo[1][j], acc[j]);
            acc[j] = __ffma2_rn(fa[ky0 + 2], hi[0][j], acc[j]);
            acc[j] = __ffma2_rn(fb[ky0 + 2], hi[1][j], acc[j]);
        }
        if (col_ok && oy0 + o < p.out_h) store8(yb + (int64_t)(oy0 + o) * p.out_w * p.c, acc);
    }
}

template <int CH>
__global__ void __launch_bounds__(kFirThreads) fir_up2_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ FirResampleParams p) {
    constexpr int NCG = CH / 8, TW = kFirThreads / NCG, R = 8, BOXW = TW / 2 + 2, BOXH = 6, ROWB = CH * 2;
    constexpr int PITCH = BOXW * ROWB, STAGE = (PITCH * BOXH + 1023) & ~1023, S = 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t full[S];
    __shared__ float taps[16];
    const int tid = threadIdx.x;
    load_taps(taps, p);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        fence_barrier_init();
        prefetch_tensormap(&map_x);
    }
    __syncthreads();
    const int cg = tid % NCG, tx = tid / NCG;
    // column geometry of this thread (tile origins are even, so it does not depend on the tile)
    const int kx0 = (p.pad0_x - tx) & 1;                                     // (tx + kx0 - pad0_x) is even
    const int lc0 = floordiv(tx + kx0 - p.pad0_x, 2) - floordiv(-p.pad0_x, 2);
    const int q = p.pad0_y & 1;

    auto issue = [&](int tile, int stage) {
        int chunk, bx, by, b;
        decode_tile(p, tile, chunk, bx, by, b);
        mbar_arrive_expect_tx(full + stage, (uint32_t)(PITCH * BOXH));
        tma_load_4d(smem + stage * STAGE, &map_x, full + stage, chunk * CH, floordiv(bx * TW - p.pad0_x, 2),
                    floordiv(by * R - p.pad0_y, 2), b);
    };
    if (tid == 0)
        for (int s = 0; s < S - 1; ++s)
            if ((int)(blockIdx.x + s * gridDim.x) < p.total_tiles) issue(blockIdx.x + s * gridDim.x, s);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int stage = it % S;
        const int next = tile + (S - 1) * gridDim.x;
        if (tid == 0 && next < p.total_tiles) issue(next, (it + S - 1) % S);
        mbar_wait(full + stage, (it / S) & 1);
        int chunk, bx, by, b;
        decode_tile(p, tile, chunk, bx, by, b);
        const uint8_t* tb = smem + stage * STAGE + lc0 * ROWB + cg * 16;
        const int ox = bx * TW + tx, oy0 = by * R;
        __nv_bfloat16* yb = p.y + (((int64_t)b * p.out_h) * p.out_w + ox) * p.c + chunk * CH + cg * 8;
        if (q) fir_up2_rows<CH, 1>(tb, taps, kx0, yb, oy0, ox < p.out_w, p);
        else   fir_up2_rows<CH, 0>(tb, taps, kx0, yb, oy0, ox < p.out_w, p);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
bool fir_resample_tma_eligible(int dtype, int c, int kh, int kw, int up, int down, int out_h, int out_w, const void* x,
                               const void* y) {
    if (dtype != B200GAN_BF16 || kh > 4 || kw > 4) return false;
    if (!((up == 1 && down == 2) || (up == 2 && down == 1))) return false;
    if (c % 32 != 0) return false;
    if (out_h < 4 || out_w < 16) return false;
    if (((uintptr_t)x | (uintptr_t)y) % 16 != 0) return false;
    return tensor_map_encoder() != nullptr;
}

template <int ID, typename Kernel>
static int fir_launch(Kernel kernel, size_t stage_bytes, int stages, const CUtensorMap& map_x, const FirResampleParams& p,
                      cudaStream_t st, const char* what) {
    const size_t smem = 1024 + (size_t)stages * stage_bytes;
    // once per device and kernel (not a stream operation; kept out of CUDA-graph capture)
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_dev = cur_dev;
    }
    int per_sm = (int)((220 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = sm_count() * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    kernel<<<grid, kFirThreads, smem, st>>>(map_x, p);
    count_launch();
    return check_launch(what);
}

int fir_resample_tma(const void* x, void* y, const float* taps, int n, int in_h, int in_w, int c, int out_h, int out_w,
                     int kh, int kw, int up, int down, int pad0_y, int pad0_x, int flip, float gain, cudaStream_t st) {
    FirResampleParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.in_h = in_h; p.in_w = in_w; p.c = c; p.out_h = out_h; p.out_w = out_w; p.kh = kh; p.kw = kw;
    p.pad0_y = pad0_y; p.pad0_x = pad0_x; p.flip = flip; p.gain = gain; p.taps = taps; p.y = (__nv_bfloat16*)y;
    const int ch = c % 64 == 0 ? 64 : 32;
    const int tw = kFirThreads / (ch / 8);
    const int rows = down == 2 ? 4 : 8;
    p.tiles_x = (out_w + tw - 1) / tw;
    p.tiles_y = (out_h + rows - 1) / rows;
    p.chunks = c / ch;
    p.total_tiles = n * p.tiles_y * p.tiles_x * p.chunks;
    if (p.total_tiles == 0) return 0;
    p.div_chunks = make_fastdiv((uint32_t)p.chunks);
    p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
    p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
    const int box_w = down == 2 ? 2 * tw + 2 : tw / 2 + 2;
    const int box_h = down == 2 ? 2 * rows + 2 : 6;
    CUtensorMap map_x;
    uint64_t dims[4] = {(uint64_t)c, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)c * 2, (uint64_t)in_w * c * 2, (uint64_t)in_h * in_w * c * 2};
    uint32_t box[4] = {(uint32_t)ch, (uint32_t)box_w, (uint32_t)box_h, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, 0)) return e;
    const size_t stage = ((size_t)box_w * ch * 2 * box_h + 1023) & ~(size_t)1023;
    if (down == 2)
        return ch == 64 ? fir_launch<0>(fir_down2_tma_kernel<64>, stage, 2, map_x, p, st, "fir_down2_tma")
                        : fir_launch<1>(fir_down2_tma_kernel<32>, stage, 2, map_x, p, st, "fir_down2_tma");
    return ch == 64 ? fir_launch<2>(fir_up2_tma_kernel<64>, stage, 4, map_x, p, st, "fir_up2_tma")
                    : fir_launch<3>(fir_up2_tma_kernel<32>, stage, 4, map_x, p, st, "fir_up2_tma");
}

}  // namespace b200gan
