// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// One persistent, warp-specialised kernel serves every dense convolution on the StyleGAN2 path
// (include/b200gan.h "convolution family"):
//     up=1, down=1 : stride-1 3x3 / 1x1        (ModulatedConv2d plain, EqualConv2d, all s1 dgrads)
//     up=1, down=2 : stride-2 conv              (D downsampling convs; dgrad of the G up-conv)
//     up=2, down=1 : transposed stride-2 conv   (G up-conv gm.py:304; dgrad of the D s2 convs),
//                    decomposed into its 4 output phases, each a small stride-1 conv on the
//                    input grid -- no zero-insertion, no wasted MMA work.
// GEMM view per tile:  D[128 pixels][BN out-channels] += A[128 pixels][KC in-ch] * B[BN][KC]^T
// for every (tap, channel chunk).  A tiles are fetched straight from the NHWC activation by a 4-D
// TMA box {KC, TW, TH, TN} whose start coordinate carries the tap offset (out-of-bounds = zero
// padding for free; element strides give the stride-2 gather); B tiles come from the K-major
// weight tensor [wb][tap][OC][IC] (per-sample weights when wb = batch).  Accumulators live in
// TMEM, double-buffered, so the epilogue of tile i (demod scale, noise, bias, leaky-ReLU, bf16
// store) overlaps the MMAs of tile i+1.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (thread t <-> TMEM lane t <-> pixel t of the tile).
#include <cstdlib>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

constexpr int kMaxTaps = 9;
constexpr int kThreads = 256;
constexpr int kTileM = 128;

struct FwdPhase {
    int ntaps;
    int dy[kMaxTaps], dx[kMaxTaps], wtap[kMaxTaps];
    int py, px;        // output offset of the phase
    int ph, pw;        // phase-grid extent
    int tiles_h, tiles_w, tiles_n;
    int tile_begin;
    FastDiv div_tw, div_th;      // tile index decode without integer division (three warp roles decode every tile)
};

struct FwdParams {
    int B, OH, OW, OC, IC;
    int n_phases;
    FwdPhase phase[4];
    int TW, TH, TN, rows;
    int es, os;
    int BN, n_oc_tiles;
    int kchunks, kc, row_bytes, layout;
    int sub;                         // 64-channel sub-tiles per pipeline stage (2: K = 128 per barrier round, see conv_fwd_umma)
    int taps_total, w_per_sample;
    int total_tiles;
    FastDiv div_oc;
    int stages, a_stage_bytes, b_stage_bytes, tx_bytes;
    int tmem_cols;
    const float* bias;
    const float* rowscale;
    const __nv_bfloat16* noise;
    const float* noise_w;
    float slope, gain;
    int has_ep;
    int f16;                         // 16-bit storage is IEEE half
    const __nv_bfloat16* addend;     // output-shaped side inputs (include/b200gan.h b200gan_conv_epilogue)
    const __nv_bfloat16* gate;
    __nv_bfloat16* y;
};

struct TileCoord {
    int phase, n0, h0, w0, ocb;
};

__device__ __forceinline__ TileCoord decode_tile(const FwdParams& p, int tile) {
    TileCoord t;
    int ph = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < p.n_phases && tile >= p.phase[i].tile_begin) ph = i;
    const FwdPhase& P = p.phase[ph];
    // a short-K tile (one phase of a transposed convolution: 8 .. 32 MMAs) costs less tensor time than six integer
    // divisions in each of the three roles that decode it: multiply-high + shift instead (exact below 2^31)
    uint32_t r = (uint32_t)(tile - P.tile_begin), q, rem;
    t.phase = ph;
    p.div_oc.divmod(r, q, rem);
    t.ocb = (int)rem;
    P.div_tw.divmod(q, r, rem);
    t.w0 = (int)rem * p.TW;
    P.div_th.divmod(r, q, rem);
    t.h0 = (int)rem * p.TH;
    t.n0 = (int)q * p.TN;
    return t;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + p.stages * p.a_stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(b_buf + p.stages * p.b_stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull + a, 1);
            mbar_init(tempty + a, 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Role loops are executed by the WHOLE warp with warp-uniform control flow; only the TMA / MMA /
    // commit instructions sit under elect_one().  (Running the loop under `if (lane == 0)` makes every
    // descriptor a per-thread value that has to be moved to uniform registers around each UTCHMMA:
    // measured ~200 clk per MMA, profiles/r01_conv_issue_bound.md.)
    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0, par = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            const FwdPhase& P = p.phase[t.phase];
            const int wz0 = (p.w_per_sample ? t.n0 : 0) * p.taps_total;
            for (int tp = 0; tp < P.ntaps; ++tp) {
                const int cx = t.w0 * p.es + P.dx[tp], cy = t.h0 * p.es + P.dy[tp];
                for (int c = 0; c < p.kchunks; c += p.sub) {
                    mbar_wait(empty + stage, par ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + stage, (uint32_t)p.tx_bytes);
                        for (int j = 0; j < p.sub; ++j) {
                            tma_load_4d(a_buf + stage * p.a_stage_bytes + j * (kTileM * p.row_bytes), &map_x, full + stage,
                                        (c + j) * p.kc, cx, cy, t.n0);
                            tma_load_3d(b_buf + stage * p.b_stage_bytes + j * (p.BN * p.row_bytes), &map_w, full + stage,
                                        (c + j) * p.kc, t.ocb * p.BN, wz0 + P.wtap[tp]);
                        }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = instr_desc_bf16(kTileM, p.BN, 0, 0, p.f16);
        const uint32_t hi = desc_hi(8u * (uint32_t)p.row_bytes, (uint32_t)p.layout);
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), 16), b_lo0 = desc_lo(smem_u32(b_buf), 16);
        const uint32_t a_inc = (uint32_t)p.a_stage_bytes >> 4, b_inc = (uint32_t)p.b_stage_bytes >> 4;
        const int nstages = p.stages, kchunks = p.kchunks, bn = p.BN, total = p.total_tiles;
        const bool k4 = p.kc == 64;
        const uint32_t a_sub = (uint32_t)(kTileM * p.row_bytes) >> 4, b_sub = (uint32_t)(p.BN * p.row_bytes) >> 4;
        int stage = 0, par = 0, it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const TileCoord t = decode_tile(p, tile);
            const int acc = it & 1, acc_par = (it >> 1) & 1;
            mbar_wait(tempty + acc, acc_par ^ 1);
            tc_fence_after();
            const int ksteps = p.phase[t.phase].ntaps * (kchunks / p.sub);
            if (ksteps == 0) {           // phase without taps: the epilogue writes zeros
                if (elect_one()) mbar_arrive(tfull + acc);
                __syncwarp();
                continue;
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * bn);
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(full + stage, par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)stage * a_inc, b_lo = b_lo0 + (uint32_t)stage * b_inc;
                    // +32 B (= 2 descriptor units) per K = 16 step inside the swizzled row
                    if (ks == 0) mma_issue<false>(d_tmem, a_lo, hi, b_lo, hi, idesc);
                    else         mma_issue<true>(d_tmem, a_lo, hi, b_lo, hi, idesc);
                    mma_issue<true>(d_tmem, a_lo + 2, hi, b_lo + 2, hi, idesc);
                    if (k4) {
                        mma_issue<true>(d_tmem, a_lo + 4, hi, b_lo + 4, hi, idesc);
                        mma_issue<true>(d_tmem, a_lo + 6, hi, b_lo + 6, hi, idesc);
                    }
                    if (p.sub == 2) {           // second 64-channel sub-tile of the stage (always 128-byte rows: k4)
                        const uint32_t a2 = a_lo + a_sub, b2 = b_lo + b_sub;
                        mma_issue<true>(d_tmem, a2, hi, b2, hi, idesc);
                        mma_issue<true>(d_tmem, a2 + 2, hi, b2 + 2, hi, idesc);
                        mma_issue<true>(d_tmem, a2 + 4, hi, b2 + 4, hi, idesc);
                        mma_issue<true>(d_tmem, a2 + 6, hi, b2 + 6, hi, idesc);
                    }
                    mma_commit(empty + stage);
                    if (ks == ksteps - 1) mma_commit(tfull + acc);
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; par ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp - 4;
        const int r = q * 32 + lane;
        const int w_l = r % p.TW, h_l = (r / p.TW) % p.TH, n_l = r / (p.TW * p.TH);
        const float nw = (p.noise != nullptr && p.noise_w != nullptr) ? *p.noise_w : 0.f;
        const float g = p.gain, gs = p.gain * p.slope;
        const bool fast_act = g > 0.f && p.slope >= 0.f && p.slope <= 1.f;   // gain*lrelu(u) = max(g*u, g*slope*u)
        // side inputs (bias / rowscale as float4) need 16-byte aligned rows
        const bool vec_side = ((((uintptr_t)p.bias | (uintptr_t)p.rowscale) & 15) == 0) && (p.OC % 4 == 0);
        // this thread's output pixel of a tile; the raw noise bits of the NEXT tile are requested while the current
        // one is drained (the load's DRAM latency used to sit between the accumulator wait and the first store)
        auto pixel_of = [&](int tile, int& n, int& oy, int& ox, int& ocb, bool& zero_tile) -> bool {
            const TileCoord t = decode_tile(p, tile);
            const FwdPhase& P = p.phase[t.phase];
            n = t.n0 + n_l;
            const int j = t.h0 + h_l, i = t.w0 + w_l;
            oy = j * p.os + P.py;
            ox = i * p.os + P.px;
            ocb = t.ocb;
            zero_tile = P.ntaps == 0;
            return r < p.rows && n < p.B && j < P.ph && i < P.pw && oy < p.OH && ox < p.OW;
        };
        uint32_t nraw = 0u;
        auto fetch_noise = [&](int tile) {
            if (p.noise == nullptr || tile >= p.total_tiles) return;
            int n, oy, ox, ocb;
            bool zt;
            if (pixel_of(tile, n, oy, ox, ocb, zt))
                nraw = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.noise) + ((int64_t)n * p.OH + oy) * p.OW + ox);
        };
        int it = 0;
        fetch_noise(blockIdx.x);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            int n, oy, ox, ocb;
            bool zero_tile;
            const bool valid = pixel_of(tile, n, oy, ox, ocb, zero_tile);
            const int acc = it & 1, acc_par = (it >> 1) & 1;
            mbar_wait(tfull + acc, acc_par);
            tc_fence_after();
            const int64_t pix = ((int64_t)n * p.OH + oy) * p.OW + ox;
            __nv_bfloat16* dst = p.y + pix * p.OC + ocb * p.BN;
            const float nz = (valid && p.noise) ? nw * lo16(nraw, p.f16) : 0.f;
            fetch_noise(tile + gridDim.x);
            const float* rs = p.rowscale ? p.rowscale + (int64_t)(valid ? n : 0) * p.OC + ocb * p.BN : nullptr;
            const float* bs = p.bias ? p.bias + ocb * p.BN : nullptr;
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(q * 32) << 16);
            const int64_t side0 = pix * p.OC + ocb * p.BN;           // this pixel's row of the output-shaped side tensors
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                // side loads first: their latency overlaps the TMEM read of the chunk
                Side16 s_add, s_gate;
                if (valid && p.addend) s_add = side_load16(p.addend + side0 + c0);
                if (valid && p.gate) s_gate = side_load16(p.gate + side0 + c0);
                float v[16];
                if (!zero_tile) {
                    tmem_ld_x16(taddr + (uint32_t)c0, v);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.f;
                }
                if (valid) {
                    if (p.gate) {
                        // backward mode: (acc + addend) * rowscale * gain * act'(gate)
                        float rr[16];
                        if (vec_side) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 r4 = rs ? __ldg(reinterpret_cast<const float4*>(rs + c0) + e) : make_float4(1.f, 1.f, 1.f, 1.f);
                                rr[4 * e] = r4.x; rr[4 * e + 1] = r4.y; rr[4 * e + 2] = r4.z; rr[4 * e + 3] = r4.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) rr[e] = rs ? rs[c0 + e] : 1.f;
                        }
                        side_apply16(v, p.addend ? &s_add : nullptr, &s_gate, rr, g, gs, p.f16);
                    } else if (p.has_ep) {
                        if (p.addend) {
                            float one[16];
                            side_apply16(v, &s_add, nullptr, one, 1.f, 1.f, p.f16);
                        }
                        float rr[16], bb[16];
                        if (vec_side && rs == nullptr && fast_act) {
                            // no demodulation row (it is folded into the weights on the weight-modulated layers): bias +
                            // noise + leaky-ReLU on packed fp32 pairs (see conv_umma_halo.cu epilogue_math_store16)
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 b4 = bs ? __ldg(reinterpret_cast<const float4*>(bs + c0) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                                bb[4 * e] = b4.x; bb[4 * e + 1] = b4.y; bb[4 * e + 2] = b4.z; bb[4 * e + 3] = b4.w;
                            }
                            const F2 nz2 = f2_pack(nz, nz), g2 = f2_pack(g, g), gs2 = f2_pack(gs, gs);
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                const F2 u2 = f2_add(f2_pack(v[e], v[e + 1]), f2_add(f2_pack(bb[e], bb[e + 1]), nz2));
                                float t0, t1, s0, s1;
                                f2_unpack(f2_mul(u2, g2), t0, t1);
                                f2_unpack(f2_mul(u2, gs2), s0, s1);
                                v[e] = fmaxf(t0, s0);
                                v[e + 1] = fmaxf(t1, s1);
                            }
                        } else {
                        if (vec_side) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 r4 = rs ? __ldg(reinterpret_cast<const float4*>(rs + c0) + e) : make_float4(1.f, 1.f, 1.f, 1.f);
                                const float4 b4 = bs ? __ldg(reinterpret_cast<const float4*>(bs + c0) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                                rr[4 * e] = r4.x; rr[4 * e + 1] = r4.y; rr[4 * e + 2] = r4.z; rr[4 * e + 3] = r4.w;
                                bb[4 * e] = b4.x; bb[4 * e + 1] = b4.y; bb[4 * e + 2] = b4.z; bb[4 * e + 3] = b4.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                rr[e] = rs ? rs[c0 + e] : 1.f;
                                bb[e] = bs ? bs[c0 + e] : 0.f;
                            }
                        }
                        if (fast_act) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float u = fmaf(v[e], rr[e], nz + bb[e]);
                                v[e] = fmaxf(u * g, u * gs);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float u = fmaf(v[e], rr[e], nz + bb[e]);
                                v[e] = g * (u > 0.f ? u : u * p.slope);
                            }
                        }
                        }
                    }
                    uint32_t pk[8];
                    pack16(v, pk, p.f16);
                    uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
                    d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    return fn;
}

int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return B200GAN_ENOSUP;
    }
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                     (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box,
                     (const cuuint32_t*)elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return B200GAN_EINVAL;
    }
    return 0;
}

static int floordiv_h(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

// Does the tcgen05 engine take this convolution?  (Everything else goes to conv_simt.cu.)
bool conv_fwd_umma_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y) {
    if (dtype != B200GAN_BF16 && dtype != B200GAN_F16) return false;
    if (g.kh * g.kw > kMaxTaps) return false;
    if (!((g.up == 1 && (g.down == 1 || g.down == 2)) || (g.up == 2 && g.down == 1))) return false;
    if (g.ic < 32 || g.ic % 8 != 0) return false;
    if (g.ic < 64 && g.ic != 32) return false;
    if (g.oc < 16 || g.oc % 16 != 0) return false;
    if (g.oc > 256 && g.oc % 128 != 0) return false;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) % 16 != 0) return false;
    if ((int64_t)g.b * g.out_h * g.out_w < 64) return false;          // too small to be worth a launch of this shape
    return tensor_map_encoder() != nullptr;
}

int conv_fwd_umma(const void* x, const void* w, void* y, const ConvGeom& g, const ConvEp& ep, cudaStream_t st) {
    const float* bias = ep.bias;
    const float* rowscale = ep.rowscale;
    const void* noise = ep.noise;
    const float* noise_w = ep.noise_w;
    const float slope = ep.slope, gain = ep.gain;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.B = g.b; p.OH = g.out_h; p.OW = g.out_w; p.OC = g.oc; p.IC = g.ic;
    p.es = g.down; p.os = g.up;
    p.taps_total = g.kh * g.kw;
    p.w_per_sample = g.w_per_sample;
    // ---- phases ----
    int max_ph = 0, max_pw = 0;
    if (g.up == 1) {
        p.n_phases = 1;
        FwdPhase& P = p.phase[0];
        for (int ky = 0; ky < g.kh; ++ky)
            for (int kx = 0; kx < g.kw; ++kx) {
                P.dy[P.ntaps] = ky - g.pad0;
                P.dx[P.ntaps] = kx - g.pad0;
                P.wtap[P.ntaps++] = ky * g.kw + kx;
            }
        P.ph = g.out_h; P.pw = g.out_w;
        max_ph = P.ph; max_pw = P.pw;
    } else {
        p.n_phases = 0;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                FwdPhase& P = p.phase[p.n_phases];
                P.py = py; P.px = px;
                P.ph = (g.out_h - py + 1) / 2;
                P.pw = (g.out_w - px + 1) / 2;
                if (P.ph <= 0 || P.pw <= 0) continue;
                for (int ky = 0; ky < g.kh; ++ky) {
                    if (((py + ky - g.pad0) % 2 + 2) % 2 != 0) continue;
                    for (int kx = 0; kx < g.kw; ++kx) {
                        if (((px + kx - g.pad0) % 2 + 2) % 2 != 0) continue;
                        P.dy[P.ntaps] = floordiv_h(py + ky - g.pad0, 2);
                        P.dx[P.ntaps] = floordiv_h(px + kx - g.pad0, 2);
                        P.wtap[P.ntaps++] = ky * g.kw + kx;
                    }
                }
                max_ph = P.ph > max_ph ? P.ph : max_ph;
                max_pw = P.pw > max_pw ? P.pw : max_pw;
                ++p.n_phases;
            }
    }
    // ---- tile geometry ----
    p.TW = max_pw < 16 ? max_pw : 16;
    p.TH = max_ph < kTileM / p.TW ? max_ph : kTileM / p.TW;
    p.TN = g.w_per_sample ? 1 : kTileM / (p.TW * p.TH);
    if (p.TN < 1) p.TN = 1;
    if (p.TN > g.b) p.TN = g.b;
    p.rows = p.TW * p.TH * p.TN;
    p.BN = g.oc <= 256 ? g.oc : (g.oc % 256 == 0 ? 256 : 128);
    // Low-resolution layers (4^2 .. 32^2) have only a handful of pixel tiles, and each tile walks the whole K = taps*IC
    // loop serially (512 -> 512: 288 MMAs): with BN = 256 a 8^2 layer occupies 16 of the 148 SMs for 288 x 128 clk.
    // Narrower N tiles spread the same work over more SMs (the A tile is re-read from L2 per N tile; an MMA costs
    // 32 + N/4 clk for N <= 128, scripts/umma_pacing.cu, so the total tensor time grows only mildly).
    {
        int m_tiles = 0;
        for (int i = 0; i < p.n_phases; ++i) {
            const FwdPhase& P = p.phase[i];
            m_tiles += ((P.ph + p.TH - 1) / p.TH) * ((P.pw + p.TW - 1) / p.TW) * ((g.b + p.TN - 1) / p.TN);
        }
        while (p.BN > 32 && p.BN % 32 == 0 && m_tiles * (g.oc / p.BN) < sm_count() && g.oc % (p.BN / 2) == 0) p.BN /= 2;
    }
    p.n_oc_tiles = g.oc / p.BN;
    p.row_bytes = g.ic >= 64 ? 128 : 64;
    p.layout = p.row_bytes == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.kc = p.row_bytes / 2;
    p.kchunks = (g.ic + p.kc - 1) / p.kc;
    int tiles = 0;
    for (int i = 0; i < p.n_phases; ++i) {
        FwdPhase& P = p.phase[i];
        P.tiles_h = (P.ph + p.TH - 1) / p.TH;
        P.tiles_w = (P.pw + p.TW - 1) / p.TW;
        P.tiles_n = (g.b + p.TN - 1) / p.TN;
        P.tile_begin = tiles;
        P.div_tw = make_fastdiv((uint32_t)P.tiles_w);
        P.div_th = make_fastdiv((uint32_t)P.tiles_h);
        tiles += P.tiles_h * P.tiles_w * P.tiles_n * p.n_oc_tiles;
    }
    p.total_tiles = tiles;
    p.div_oc = make_fastdiv((uint32_t)p.n_oc_tiles);
    // K = 128 per pipeline stage where the stage is short: at N <= 128 the four MMAs of a 64-channel stage (<= 256 clk)
    // take no longer than the issuing warp's barrier round (wait, fence, elect, commit, ring advance), so the general
    // engine ran issue-bound there (128 -> 128 @256^2: 50 % of the tensor peak, N = 64 transposed-convolution phases
    // 25 %; profiles/r02_launches_step.md).  Two sub-tiles per stage halve the rounds per tile.
    p.sub = (g.ic % 128 == 0 && p.row_bytes == 128 && p.BN <= 128) ? 2 : 1;
    if (const char* e = getenv("B200GAN_UMMA_SUB")) {               // timing experiments: 1 = one sub-tile per stage
        if (atoi(e) == 1) p.sub = 1;
    }
    p.a_stage_bytes = p.sub * kTileM * p.row_bytes;
    p.b_stage_bytes = p.sub * p.BN * p.row_bytes;
    p.tx_bytes = p.sub * (p.rows * p.row_bytes + p.BN * p.row_bytes);
    const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 12) p.stages = 12;
    if (p.stages < 2) p.stages = 2;
    p.tmem_cols = next_pow2(2 * p.BN);
    p.bias = bias; p.rowscale = rowscale; p.noise = (const __nv_bfloat16*)noise; p.noise_w = noise_w;
    p.slope = slope; p.gain = gain;
    p.has_ep = ep_active(ep) ? 1 : 0;
    p.f16 = g.f16;
    p.addend = (const __nv_bfloat16*)ep.addend;
    p.gate = (const __nv_bfloat16*)ep.gate;
    p.y = (__nv_bfloat16*)y;

    // ---- tensor maps ----
    CUtensorMap map_x, map_w;
    {
        uint64_t dims[4] = {(uint64_t)g.ic, (uint64_t)g.in_w, (uint64_t)g.in_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.ic * 2, (uint64_t)g.in_w * g.ic * 2, (uint64_t)g.in_h * g.in_w * g.ic * 2};
        uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)(p.TW * p.es), (uint32_t)(p.TH * p.es), (uint32_t)p.TN};
        uint32_t es[4] = {1, (uint32_t)p.es, (uint32_t)p.es, 1};
        if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, p.row_bytes)) return e;
    }
    {
        const int wb = g.w_per_sample ? g.b : 1;
        uint64_t dims[3] = {(uint64_t)g.ic, (uint64_t)g.oc, (uint64_t)wb * p.taps_total};
        uint64_t strides[2] = {(uint64_t)g.ic * 2, (uint64_t)g.oc * g.ic * 2};
        uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)p.BN, 1};
        uint32_t es[3] = {1, 1, 1};
        if (int e = encode_bf16_map(&map_w, w, 3, dims, strides, box, es, p.row_bytes)) return e;
    }
    const size_t smem = 1024 + (size_t)p.stages * stage_bytes + (2 * p.stages + 4) * sizeof(uint64_t) + 64;
    // once per device (not a stream operation; kept out of CUDA-graph capture)
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(conv_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_dev = cur_dev;
    }
    int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    conv_fwd_umma_kernel<<<grid, kThreads, smem, st>>>(map_x, map_w, p);
    count_launch();
    return check_launch("conv_fwd_umma");
}

}  // namespace b200gan
