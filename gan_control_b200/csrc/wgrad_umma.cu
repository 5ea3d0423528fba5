// tcgen05 / TMEM weight-gradient kernel of the convolution family (bf16 operands, fp32 accumulate).
//
//   gw[wb][tap][o][i] += sum_{b,oy,ox} gy[b][oy][ox][o] * z[b][oy*down+ky-pad0][ox*down+kx-pad0][i]
//
// GEMM view: the reduction (K) dimension is PIXELS, so both operands are "MN-major" straight out of
// the NHWC tensors: a TMA box of pixels x 64 (or 32) channels lands in shared memory exactly as one
// swizzle atom column of a tcgen05 MN-major operand.
//   M side (A operand) : 128 rows = input channels of the x tile, gathered tap by tap.  For IC >= 128
//                        one tap x 128 channels; for IC = 64 two taps x 64; for IC = 32 four taps x 32:
//                        small-channel (1024^2 / 512^2) layers still fill the 128 MMA rows.
//   N side (B operand) : up to 128 output channels of the gy tile (loaded ONCE per pixel tile and
//                        shared by all taps).
//   D[(tap,i)][o] accumulates in TMEM for every tap group at once (<= 512 columns), over a split-K
//   range of pixel tiles; the epilogue adds it into gw with fp32 atomics (lanes = consecutive i,
//   coalesced).
// Geometry: stride-2 convs read x through TMA element strides; transposed convs (up = 2) are split
// into their 4 output-parity phases and read gy through element strides.  Out-of-bounds = zeros.
#include <cstdlib>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

constexpr int kWgMaxTaps = 9;
constexpr int kWgThreads = 256;

struct WgGroup {
    int natoms;              // valid atoms in this M tile
    int tap[4];              // index into the phase tap list
    int choff[4];            // channel offset inside the icb block
};

struct WgPhase {
    int py, px, ph, pw;      // gy sampling offset and iteration-grid extent
    int ntaps;
    int dy[kWgMaxTaps], dx[kWgMaxTaps], wtap[kWgMaxTaps];
    int ngroups, nchunks;    // tap groups, and chunks of <= gu groups (one chunk per work unit)
    WgGroup group[kWgMaxTaps];
    int tiles_h, tiles_w, tiles_n, ktiles, splits, kt_per_split;
    int unit_begin, units;
};

struct WgParams {
    int B, IC, OC, taps_total, per_sample;
    int n_phases;
    WgPhase phase[4];
    int TW, TH, TN, rows;
    int sa, sb;              // element strides of gy / x
    int atom_m, rowb_m, layout_m, slab_m, natoms_m;     // M side: channels per atom, row bytes, slab bytes
    int atom_n, rowb_n, layout_n, slab_n, nslabs_n;     // N side
    int N, n_ocb, n_icb, gu;
    int total_units;
    int sa_stages, sb_stages, a_slot_bytes, b_slot_bytes;
    int tmem_cols;
    int issuers;             // MMA-issuing warps (1..3): accumulator group gl of a unit is issued by warp 1 + gl % issuers
    int f16;                 // operands are IEEE half instead of bfloat16
    float* gw;
};

struct WgUnit {
    int phase, wb, icb, ocb, g0, ng, kt0, kt1;
};

__device__ __forceinline__ WgUnit wg_decode(const WgParams& p, int unit) {
    WgUnit u;
    int ph = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < p.n_phases && unit >= p.phase[i].unit_begin) ph = i;
    const WgPhase& P = p.phase[ph];
    int r = unit - P.unit_begin;
    u.phase = ph;
    const int split = r % P.splits;
    r /= P.splits;
    const int chunk = r % P.nchunks;
    r /= P.nchunks;
    u.ocb = r % p.n_ocb;
    r /= p.n_ocb;
    u.icb = r % p.n_icb;
    u.wb = r / p.n_icb;
    u.g0 = chunk * p.gu;
    u.ng = min(p.gu, P.ngroups - u.g0);
    u.kt0 = split * P.kt_per_split;
    u.kt1 = min(P.ktiles, u.kt0 + P.kt_per_split);
    return u;
}

__device__ __forceinline__ void wg_ktile(const WgParams& p, const WgPhase& P, const WgUnit& u, int kt, int& n0, int& h0, int& w0) {
    w0 = (kt % P.tiles_w) * p.TW;
    kt /= P.tiles_w;
    h0 = (kt % P.tiles_h) * p.TH;
    n0 = p.per_sample ? u.wb : (kt / P.tiles_h) * p.TN;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_gy,
                       const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* a_buf = smem;                                             // M-side ring (x tap tiles)
    uint8_t* b_buf = a_buf + p.sa_stages * p.a_slot_bytes;             // N-side ring (gy tiles)
    uint64_t* afull = reinterpret_cast<uint64_t*>(b_buf + p.sb_stages * p.b_slot_bytes);
    uint64_t* aempty = afull + p.sa_stages;
    uint64_t* bfull = aempty + p.sa_stages;
    uint64_t* bempty = bfull + p.sb_stages;
    uint64_t* tfull = bempty + p.sb_stages;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_gy);
    }
    if (warp == 1 && lane == 0) {
        // Every issuing warp waits on EVERY full barrier and arrives on EVERY empty one, owner of the stage or not: a parity
        // wait cannot tell "one phase early" from "done" nor "two phases late" from "pending", so a warp that skipped the
        // stages of the other warps' groups could pass a wait before the previous fill had landed, or be lapped by the
        // producer (the first multi-issuer version did skip them and an epilogue wait timed out once; profiles/r02_summary.md).
        const uint32_t n_iss = (uint32_t)p.issuers;
        for (int s = 0; s < p.sa_stages; ++s) { mbar_init(afull + s, 1); mbar_init(aempty + s, n_iss); }
        for (int s = 0; s < p.sb_stages; ++s) { mbar_init(bfull + s, 1); mbar_init(bempty + s, n_iss); }
        mbar_init(tfull, n_iss);
        mbar_init(tempty, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (whole warp, elected lane issues) =================
        int as = 0, apar = 0, bs = 0, bpar = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            const WgUnit u = wg_decode(p, unit);
            const WgPhase& P = p.phase[u.phase];
            for (int kt = u.kt0; kt < u.kt1; ++kt) {
                int n0, h0, w0;
                wg_ktile(p, P, u, kt, n0, h0, w0);
                mbar_wait(bempty + bs, bpar ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(bfull + bs, (uint32_t)(p.nslabs_n * p.slab_n));
                    for (int s = 0; s < p.nslabs_n; ++s)
                        tma_load_4d(b_buf + bs * p.b_slot_bytes + s * p.slab_n, &map_gy, bfull + bs,
                                    u.ocb * p.N + s * p.atom_n, w0 * p.sa + P.px, h0 * p.sa + P.py, n0);
                }
                __syncwarp();
                if (++bs == p.sb_stages) { bs = 0; bpar ^= 1; }
                for (int g = u.g0; g < u.g0 + u.ng; ++g) {
                    const WgGroup& G = P.group[g];
                    mbar_wait(aempty + as, apar ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(afull + as, (uint32_t)(G.natoms * p.slab_m));
                        for (int a = 0; a < G.natoms; ++a) {
                            const int tp = G.tap[a];
                            tma_load_4d(a_buf + as * p.a_slot_bytes + a * p.slab_m, &map_x, afull + as,
                                        u.icb * 128 + G.choff[a], w0 * p.sb + P.dx[tp], h0 * p.sb + P.dy[tp], n0);
                        }
                    }
                    __syncwarp();
                    if (++as == p.sa_stages) { as = 0; apar ^= 1; }
                }
            }
        }
    } else if (warp >= 1 && warp <= p.issuers) {
        // ================= MMA issuers (whole warps, elected lane issues) =================
        // The accumulator groups of a unit are independent: group gl is issued by warp 1 + gl % issuers.  All issuers walk
        // both rings in lockstep with the producer (see the barrier counts above).
        const int mine = warp - 1, n_iss = p.issuers;
        const uint32_t idesc = instr_desc_bf16(128, p.N, 1, 1, p.f16);        // both operands MN-major
        const int ksteps = p.rows / 16;
        const uint32_t a_hi = desc_hi(8u * (uint32_t)p.rowb_m, (uint32_t)p.layout_m);
        const uint32_t b_hi = desc_hi(8u * (uint32_t)p.rowb_n, (uint32_t)p.layout_n);
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), (uint32_t)p.slab_m), b_lo0 = desc_lo(smem_u32(b_buf), (uint32_t)p.slab_n);
        const uint32_t a_inc = (uint32_t)p.a_slot_bytes >> 4, b_inc = (uint32_t)p.b_slot_bytes >> 4;
        const uint32_t a_kstep = (uint32_t)p.rowb_m, b_kstep = (uint32_t)p.rowb_n;     // 16 rows * row_bytes >> 4
        const int sa_stages = p.sa_stages, sb_stages = p.sb_stages, nn = p.N, total = p.total_units;
        int as = 0, apar = 0, bs = 0, bpar = 0, it = 0, owner = 0;
        for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++it) {
            const WgUnit u = wg_decode(p, unit);
            mbar_wait(tempty, (it & 1) ^ 1);
            tc_fence_after();
            for (int kt = u.kt0; kt < u.kt1; ++kt) {
                mbar_wait(bfull + bs, bpar);
                const uint32_t b_lo = b_lo0 + (uint32_t)bs * b_inc;
                const uint32_t first = kt != u.kt0;
                owner = 0;
                for (int gl = 0; gl < u.ng; ++gl) {
                    mbar_wait(afull + as, apar);
                    tc_fence_after();
                    if (elect_one()) {
                        if (owner == mine) {
                            const uint32_t a_lo = a_lo0 + (uint32_t)as * a_inc;
                            const uint32_t d_tmem = tmem_base + (uint32_t)(gl * nn);
                            mma_issue_dyn(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, first);
                            for (int ks = 1; ks < ksteps; ++ks)
                                mma_issue<true>(d_tmem, a_lo + (uint32_t)ks * a_kstep, a_hi, b_lo + (uint32_t)ks * b_kstep, b_hi, idesc);
                            mma_commit(aempty + as);
                        } else {
                            mbar_arrive(aempty + as);        // seen, not read
                        }
                        if (gl == u.ng - 1) {
                            // a commit with no MMA of this warp pending arrives at once
                            mma_commit(bempty + bs);
                            if (kt == u.kt1 - 1) mma_commit(tfull);
                        }
                    }
                    __syncwarp();
                    if (++owner == n_iss) owner = 0;
                    if (++as == sa_stages) { as = 0; apar ^= 1; }
                }
                if (++bs == sb_stages) { bs = 0; bpar ^= 1; }
            }
            if (u.kt1 <= u.kt0) {
                if (elect_one()) mbar_arrive(tfull);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> fp32 atomics into gw =================
        const int q = warp - 4;
        const int m = q * 32 + lane;
        const int atom = m / p.atom_m, within = m % p.atom_m;
        int it = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++it) {
            const WgUnit u = wg_decode(p, unit);
            const WgPhase& P = p.phase[u.phase];
            mbar_wait(tfull, it & 1);
            tc_fence_after();
            const bool have = u.kt1 > u.kt0;
            for (int gl = 0; gl < u.ng; ++gl) {
                const WgGroup& G = P.group[u.g0 + gl];
                const bool row_ok = have && atom < G.natoms;
                const int tp = row_ok ? G.tap[atom] : 0;
                const int ic = u.icb * 128 + (row_ok ? G.choff[atom] : 0) + within;
                const bool valid = row_ok && ic < p.IC;
                float* dst = p.gw + (((int64_t)u.wb * p.taps_total + P.wtap[tp]) * p.OC + u.ocb * p.N) * p.IC + ic;
                const uint32_t taddr = tmem_base + (uint32_t)(gl * p.N) + ((uint32_t)(q * 32) << 16);
                for (int c0 = 0; c0 < p.N; c0 += 16) {
                    float v[16];
                    if (have) {
                        tmem_ld_x16(taddr + (uint32_t)c0, v);
                    }
                    if (valid) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (u.ocb * p.N + c0 + e < p.OC) atomicAdd(dst + (int64_t)(c0 + e) * p.IC, v[e]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
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
static int wg_floordiv(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
static int wg_pow2_ge(int v, int lo) {
    int p = lo;
    while (p < v) p <<= 1;
    return p;
}

bool conv_wgrad_umma_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy) {
    if (dtype != B200GAN_BF16 && dtype != B200GAN_F16) return false;
    if (g.kh * g.kw > kWgMaxTaps) return false;
    if (!((g.up == 1 && (g.down == 1 || g.down == 2)) || (g.up == 2 && g.down == 1))) return false;
    if (!(g.ic == 32 || g.ic == 64 || g.ic % 128 == 0)) return false;
    if (!(g.oc == 32 || g.oc == 64 || g.oc % 128 == 0)) return false;
    if (((uintptr_t)x | (uintptr_t)gy) % 16 != 0) return false;
    if ((int64_t)g.out_h * g.out_w < 16) return false;
    return tensor_map_encoder() != nullptr;
}

int conv_wgrad_umma(const void* x, const void* gy, float* gw, const ConvGeom& g, cudaStream_t st) {
    WgParams p;
    memset(&p, 0, sizeof(p));
    p.B = g.b; p.IC = g.ic; p.OC = g.oc; p.taps_total = g.kh * g.kw; p.per_sample = g.w_per_sample;
    p.sa = g.up; p.sb = g.down;
    p.gw = gw;
    p.f16 = g.f16;
    // operand geometry
    p.atom_m = g.ic >= 64 ? 64 : 32;
    p.rowb_m = p.atom_m * 2;
    p.layout_m = p.rowb_m == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.natoms_m = 128 / p.atom_m;
    p.n_icb = g.ic >= 128 ? g.ic / 128 : 1;
    p.atom_n = g.oc >= 64 ? 64 : 32;
    p.rowb_n = p.atom_n * 2;
    p.layout_n = p.rowb_n == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.N = g.oc < 128 ? g.oc : 128;
    p.nslabs_n = p.N / p.atom_n;
    p.n_ocb = g.oc / p.N;
    p.gu = 512 / p.N;
    // ---- phases and taps ----
    int max_ph = 0, max_pw = 0;
    const int nph = g.up == 2 ? 4 : 1;
    for (int q = 0; q < nph; ++q) {
        const int py = g.up == 2 ? q / 2 : 0, px = g.up == 2 ? q % 2 : 0;
        WgPhase P;
        memset(&P, 0, sizeof(P));
        P.py = py; P.px = px;
        P.ph = g.up == 2 ? (g.out_h - py + 1) / 2 : g.out_h;
        P.pw = g.up == 2 ? (g.out_w - px + 1) / 2 : g.out_w;
        if (P.ph <= 0 || P.pw <= 0) continue;
        for (int ky = 0; ky < g.kh; ++ky)
            for (int kx = 0; kx < g.kw; ++kx) {
                if (g.up == 2) {
                    if (((py + ky - g.pad0) % 2 + 2) % 2 != 0 || ((px + kx - g.pad0) % 2 + 2) % 2 != 0) continue;
                    P.dy[P.ntaps] = wg_floordiv(py + ky - g.pad0, 2);
                    P.dx[P.ntaps] = wg_floordiv(px + kx - g.pad0, 2);
                } else {
                    P.dy[P.ntaps] = ky - g.pad0;
                    P.dx[P.ntaps] = kx - g.pad0;
                }
                P.wtap[P.ntaps++] = ky * g.kw + kx;
            }
        if (P.ntaps == 0) continue;          // a phase no tap reads contributes nothing
        // tap groups: fill the 128 M rows with (tap, channel-atom) pairs
        if (g.ic >= 128) {
            for (int t = 0; t < P.ntaps; ++t) {
                WgGroup& G = P.group[P.ngroups++];
                G.natoms = 2;
                G.tap[0] = G.tap[1] = t;
                G.choff[0] = 0; G.choff[1] = 64;
            }
        } else {
            const int per = p.natoms_m;      // 2 taps of 64 channels or 4 taps of 32
            for (int t = 0; t < P.ntaps; t += per) {
                WgGroup& G = P.group[P.ngroups++];
                G.natoms = P.ntaps - t < per ? P.ntaps - t : per;
                for (int a = 0; a < G.natoms; ++a) { G.tap[a] = t + a; G.choff[a] = 0; }
            }
        }
        P.nchunks = (P.ngroups + p.gu - 1) / p.gu;
        max_ph = P.ph > max_ph ? P.ph : max_ph;
        max_pw = P.pw > max_pw ? P.pw : max_pw;
        p.phase[p.n_phases++] = P;
    }
    if (p.n_phases == 0) return 0;
    // ---- pixel (K) tiles: power-of-two boxes, every row written by TMA (zeros when out of bounds) ----
    p.TW = wg_pow2_ge(max_pw < 16 ? max_pw : 16, 4);
    if (p.TW > 16) p.TW = 16;
    int th_cap = 128 / p.TW;
    p.TH = wg_pow2_ge(max_ph < th_cap ? max_ph : th_cap, 1);
    if (p.TH > th_cap) p.TH = th_cap;
    p.TN = g.w_per_sample ? 1 : 128 / (p.TW * p.TH);
    if (p.TN < 1) p.TN = 1;
    p.rows = p.TW * p.TH * p.TN;
    if (p.rows % 16 != 0) {
        set_error("conv_wgrad_umma: K tile of %d rows", p.rows);
        return B200GAN_ENOSUP;
    }
    p.slab_m = p.rows * p.rowb_m;
    p.slab_n = p.rows * p.rowb_n;
    p.a_slot_bytes = p.natoms_m * p.slab_m;
    p.b_slot_bytes = p.nslabs_n * p.slab_n;
    p.sb_stages = 2;
    p.sa_stages = (int)((200 * 1024 - p.sb_stages * p.b_slot_bytes) / p.a_slot_bytes);
    if (p.sa_stages > 6) p.sa_stages = 6;
    if (p.sa_stages < 2) p.sa_stages = 2;
    // ---- work units: (wb, icb, ocb, tap chunk, split-K) ----
    const int wbs = g.w_per_sample ? g.b : 1;
    int base_units = 0;
    for (int i = 0; i < p.n_phases; ++i) base_units += wbs * p.n_icb * p.n_ocb * p.phase[i].nchunks;
    const int target = 2 * sm_count();
    int units = 0;
    for (int i = 0; i < p.n_phases; ++i) {
        WgPhase& P = p.phase[i];
        P.tiles_h = (P.ph + p.TH - 1) / p.TH;
        P.tiles_w = (P.pw + p.TW - 1) / p.TW;
        P.tiles_n = g.w_per_sample ? 1 : (g.b + p.TN - 1) / p.TN;
        P.ktiles = P.tiles_h * P.tiles_w * P.tiles_n;
        int want = (target + base_units - 1) / base_units;
        int max_splits = (P.ktiles + 3) / 4;                     // >= 4 pixel tiles per unit
        if (max_splits < 1) max_splits = 1;
        P.splits = want < max_splits ? want : max_splits;
        if (P.splits < 1) P.splits = 1;
        P.kt_per_split = (P.ktiles + P.splits - 1) / P.splits;
        P.splits = (P.ktiles + P.kt_per_split - 1) / P.kt_per_split;
        P.unit_begin = units;
        P.units = wbs * p.n_icb * p.n_ocb * P.nchunks * P.splits;
        units += P.units;
    }
    p.total_units = units;
    int max_groups = 0;
    for (int i = 0; i < p.n_phases; ++i) {
        int m = p.phase[i].ngroups < p.gu ? p.phase[i].ngroups : p.gu;
        max_groups = m > max_groups ? m : max_groups;
    }
    p.tmem_cols = wg_pow2_ge(max_groups * p.N, 32);
    // One issuer ships.  B200GAN_WGRAD_GENERAL_ISSUERS=2|3 spreads the accumulator groups over that many warps (-9..-13 % on
    // the 128..512-channel layers); it has passed scripts/stress_wgrad_general.py but not a full round of use.
    p.issuers = 1;
    if (const char* e = getenv("B200GAN_WGRAD_GENERAL_ISSUERS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 3) p.issuers = v < max_groups ? v : (max_groups < 1 ? 1 : max_groups);
    }

    CUtensorMap map_x, map_gy;
    {
        uint64_t dims[4] = {(uint64_t)g.ic, (uint64_t)g.in_w, (uint64_t)g.in_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.ic * 2, (uint64_t)g.in_w * g.ic * 2, (uint64_t)g.in_h * g.in_w * g.ic * 2};
        uint32_t box[4] = {(uint32_t)p.atom_m, (uint32_t)(p.TW * p.sb), (uint32_t)(p.TH * p.sb), (uint32_t)p.TN};
        uint32_t es[4] = {1, (uint32_t)p.sb, (uint32_t)p.sb, 1};
        if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, p.rowb_m)) return e;
    }
    {
        uint64_t dims[4] = {(uint64_t)g.oc, (uint64_t)g.out_w, (uint64_t)g.out_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.oc * 2, (uint64_t)g.out_w * g.oc * 2, (uint64_t)g.out_h * g.out_w * g.oc * 2};
        uint32_t box[4] = {(uint32_t)p.atom_n, (uint32_t)(p.TW * p.sa), (uint32_t)(p.TH * p.sa), (uint32_t)p.TN};
        uint32_t es[4] = {1, (uint32_t)p.sa, (uint32_t)p.sa, 1};
        if (int e = encode_bf16_map(&map_gy, gy, 4, dims, strides, box, es, p.rowb_n)) return e;
    }
    // the M operand always spans 128 rows: unused atoms read (and ignore) whatever follows in the ring,
    // so keep one extra slot of slack behind the rings
    const size_t smem = 1024 + (size_t)p.sa_stages * p.a_slot_bytes + (size_t)p.sb_stages * p.b_slot_bytes +
                        (2 * p.sa_stages + 2 * p.sb_stages + 2) * sizeof(uint64_t) + 16;
    // once per device (not a stream operation; kept out of CUDA-graph capture)
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(conv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_dev = cur_dev;
    }
    int grid = p.total_units < sm_count() ? p.total_units : sm_count();
    conv_wgrad_umma_kernel<<<grid, kWgThreads, smem, st>>>(map_x, map_gy, p);
    count_launch();
    return check_launch("conv_wgrad_umma");
}

}  // namespace b200gan
