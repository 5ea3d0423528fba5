// tcgen05 weight gradient, halo-reuse variant for the <= 64-channel stride-1 3x3 / 1x1 layers (the 1024^2 and
// 512^2 convolutions): like conv_umma_halo.cu, the x region of a 16x8-pixel tile is fetched ONCE
// ((16+k-1) x 16 pixels) and every tap operand is a descriptor into that buffer.
//   K (reduction) = the 128 pixels of the tile; an 8-pixel tile row is one MN-major swizzle atom group, the
//                   next row is one buffer row pitch (16 pixels) further: SBO = 16 * row_bytes.
//   M            = (kx, input channel): the kx-shifted views of one ky row are exactly row_bytes apart, so they
//                   are consecutive M atoms with LBO = row_bytes (overlapping atoms, plain address arithmetic);
//                   IC = 32: one MMA group per ky (kx = 0,1,2 + one ignored atom), IC = 64: (kx 0,1) and (kx 2).
//   N            = output channels of the dense gy tile.
// All taps accumulate in TMEM (<= 6 groups x 64 columns); one fp32 atomic pass per work unit.
#include <cstdlib>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

constexpr int kWhThreads = 256;
constexpr int kWhTW = 8, kWhTH = 16;

struct WhParams {
    int B, H, W, IC, OC, k, pad0, per_sample;
    int tiles_h, tiles_w, tiles_per_img;
    FastDiv div_tpi, div_tw;                              // pixel tile -> (sample, tile row, tile column) without division
    int units, units_per_wb, kt_per_unit, kt_total;       // split-K over pixel tiles
    int rowb_m, layout_m, rowb_n, layout_n;
    int PW, RH, a_slot_bytes, b_slot_bytes, stages;
    int ngroups, atoms_per_group;                         // MMA groups per tile, kx taps packed per group
    int issuers;                                          // MMA-issuing warps (1..3): group g is issued by warp 1 + g % issuers
    int tmem_cols;
    // space-to-depth views (ConvGeom::pack_*): the packed operand is read one row phase py at a time through a 5-D
    // map (its (px, c) pair row is contiguous), IC / OC above are then the channels of that HALF of the view and the
    // result lands at channel offset ic_off / oc_off of the full (OC_total, IC_total) weight gradient
    int x_packed, gy_packed, x_phase, gy_phase;
    int IC_total, OC_total, ic_off, oc_off;
    float* gw;
    int f16;                 // operands are IEEE half instead of bfloat16
};

// KDIM (kernel size 3 / 1) and APG (kx taps per MMA group) are compile-time so that the issuing warps' group loop is
// straight-line code with immediate descriptor offsets (the same change took the forward kernel from 0.82 to 0.66 ms).
template <int KDIM, int APG>
__global__ void __launch_bounds__(kWhThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_gy,
                       const __grid_constant__ WhParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* a_buf = smem;                                            // x halo tiles
    uint8_t* b_buf = a_buf + p.stages * p.a_slot_bytes;               // gy tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(b_buf + p.stages * p.b_slot_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_gy);
    }
    if (warp == 1 && lane == 0) {
        // every issuing warp commits to a stage's `empty` and to `tfull`: the barriers complete when all of them have
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, (uint32_t)p.issuers); }
        mbar_init(tfull, (uint32_t)p.issuers);
        mbar_init(tempty, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (weight batch wb, pixel-tile range).  Shared weights: tiles run over the whole batch.
    auto unit_range = [&](int unit, int& wb, int& kt0, int& kt1) {
        wb = unit / p.units_per_wb;
        const int s = unit % p.units_per_wb;
        kt0 = s * p.kt_per_unit;
        kt1 = min(p.kt_total, kt0 + p.kt_per_unit);
    };
    auto tile_coord = [&](int wb, int kt, int& n, int& h0, int& w0) {
        uint32_t q, t, th, tw;
        p.div_tpi.divmod((uint32_t)kt, q, t);
        p.div_tw.divmod(t, th, tw);
        n = p.per_sample ? wb : (int)q;
        h0 = (int)th * kWhTH;
        w0 = (int)tw * kWhTW;
    };

    if (warp == 0) {
        int stage = 0, par = 0;
        for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
            int wb, kt0, kt1;
            unit_range(unit, wb, kt0, kt1);
            for (int kt = kt0; kt < kt1; ++kt) {
                int n, h0, w0;
                tile_coord(wb, kt, n, h0, w0);
                mbar_wait(empty + stage, par ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full + stage, (uint32_t)(p.a_slot_bytes + p.b_slot_bytes));
                    if (p.x_packed)
                        tma_load_5d(a_buf + stage * p.a_slot_bytes, &map_x, full + stage, 0, w0 - p.pad0, p.x_phase, h0 - p.pad0, n);
                    else
                        tma_load_4d(a_buf + stage * p.a_slot_bytes, &map_x, full + stage, 0, w0 - p.pad0, h0 - p.pad0, n);
                    if (p.gy_packed)
                        tma_load_5d(b_buf + stage * p.b_slot_bytes, &map_gy, full + stage, 0, w0, p.gy_phase, h0, n);
                    else
                        tma_load_4d(b_buf + stage * p.b_slot_bytes, &map_gy, full + stage, 0, w0, h0, n);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; par ^= 1; }
            }
        }
    } else if (warp >= 1 && warp <= p.issuers) {
        // The MMA groups of a tile (one per ky for 32 channels, two per ky for 64) write DISJOINT accumulators, so they are
        // issued by up to three warps in parallel, each walking every tile: with ONE issuing warp the kernel ran at
        // 1830 clk per tile for 960 clk of MMA time (24 instructions of 40 clk) -- issue-bound like the forward kernel
        // before it got several issuers (profiles/r02_headline.md: tensor pipe 18 %, issue slots 7 %).
        const int mine = warp - 1, n_iss = p.issuers;
        const uint32_t idesc = instr_desc_bf16(128, p.OC, 1, 1, p.f16);
        const uint32_t a_hi = desc_hi((uint32_t)(p.PW * p.rowb_m), (uint32_t)p.layout_m);     // next tile row of pixels
        const uint32_t b_hi = desc_hi(8u * (uint32_t)p.rowb_n, (uint32_t)p.layout_n);
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), (uint32_t)p.rowb_m);                  // LBO = one pixel = next kx
        const uint32_t b_lo0 = desc_lo(smem_u32(b_buf), 16);
        const uint32_t a_inc = (uint32_t)p.a_slot_bytes >> 4, b_inc = (uint32_t)p.b_slot_bytes >> 4;
        constexpr uint32_t pw = KDIM == 1 ? 8u : 16u;          // buffer row pitch in pixels (WhParams::PW)
        const uint32_t row_units = (uint32_t)p.rowb_m >> 4;
        const uint32_t a_kstep = 2u * pw * row_units;          // K = 16 pixels = two tile rows
        const uint32_t b_kstep = (uint32_t)p.rowb_n;           // 16 rows * row_bytes >> 4
        const int nstages = p.stages, oc = p.OC;
        int stage = 0, par = 0, it = 0;
        for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++it) {
            int wb, kt0, kt1;
            unit_range(unit, wb, kt0, kt1);
            mbar_wait(tempty, (it & 1) ^ 1);
            tc_fence_after();
            for (int kt = kt0; kt < kt1; ++kt) {
                mbar_wait(full + stage, par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_st = a_lo0 + (uint32_t)stage * a_inc, b_st = b_lo0 + (uint32_t)stage * b_inc;
                    const uint32_t acc0 = kt != kt0;
#pragma unroll
                    for (int ky = 0; ky < KDIM; ++ky)
#pragma unroll
                        for (int kx0 = 0; kx0 < KDIM; kx0 += APG) {
                            constexpr int GPK = (KDIM + APG - 1) / APG;            // groups per ky
                            const int g = ky * GPK + kx0 / APG;
                            if (g % n_iss != mine) continue;
                            const uint32_t a_g = a_st + ((uint32_t)ky * pw + (uint32_t)kx0) * row_units;
                            const uint32_t d_tmem = tmem_base + (uint32_t)(g * oc);
                            mma_issue_dyn(d_tmem, a_g, a_hi, b_st, b_hi, idesc, acc0);
#pragma unroll
                            for (int ks = 1; ks < 8; ++ks)
                                mma_issue<true>(d_tmem, a_g + (uint32_t)ks * a_kstep, a_hi, b_st + (uint32_t)ks * b_kstep, b_hi, idesc);
                        }
                    mma_commit(empty + stage);
                    if (kt == kt1 - 1) mma_commit(tfull);
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; par ^= 1; }
            }
            if (kt1 <= kt0) {
                if (elect_one()) mbar_arrive(tfull);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int m = q * 32 + lane;
        const int atom = m / p.IC, ic = m % p.IC;          // M row = (kx - kx0, input channel)
        const int taps = p.k * p.k;
        int it = 0;
        for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++it) {
            int wb, kt0, kt1;
            unit_range(unit, wb, kt0, kt1);
            mbar_wait(tfull, it & 1);
            tc_fence_after();
            const bool have = kt1 > kt0;
            int g = 0;
            for (int ky = 0; ky < p.k; ++ky)
                for (int kx0 = 0; kx0 < p.k; kx0 += p.atoms_per_group, ++g) {
                    const int kx = kx0 + atom;
                    const bool valid = have && atom < p.atoms_per_group && kx < p.k;
                    float* dst = p.gw + (((int64_t)wb * taps + ky * p.k + (valid ? kx : 0)) * p.OC_total + p.oc_off) * p.IC_total +
                                 p.ic_off + ic;
                    const uint32_t taddr = tmem_base + (uint32_t)(g * p.OC) + ((uint32_t)(q * 32) << 16);
                    for (int c0 = 0; c0 < p.OC; c0 += 16) {
                        float v[16];
                        if (have) tmem_ld_x16(taddr + (uint32_t)c0, v);
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) atomicAdd(dst + (int64_t)(c0 + e) * p.IC_total, v[e]);
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

bool conv_wgrad_halo_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy) {
    if (dtype != B200GAN_BF16 && dtype != B200GAN_F16) return false;
    if (g.up != 1 || g.down != 1 || g.kh != g.kw || (g.kh != 1 && g.kh != 3)) return false;
    // a packed operand is processed as its two row-phase halves (see WhParams)
    const int ic = g.pack_in ? g.ic / 2 : g.ic, oc = g.pack_out ? g.oc / 2 : g.oc;
    if ((g.pack_in || g.pack_out) && g.kh != 3) return false;
    if ((g.pack_in && g.ic % 4 != 0) || (g.pack_out && g.oc % 4 != 0)) return false;
    if (!(ic == 32 || ic == 64) || !(oc == 32 || oc == 64)) return false;
    if (g.out_h < kWhTH || g.out_w < kWhTW) return false;
    if (g.out_h != g.in_h + 2 * g.pad0 - g.kh + 1 || g.out_w != g.in_w + 2 * g.pad0 - g.kw + 1) return false;
    if (((uintptr_t)x | (uintptr_t)gy) % 16 != 0) return false;
    return tensor_map_encoder() != nullptr;
}

int conv_wgrad_halo(const void* x, const void* gy, float* gw, const ConvGeom& g, cudaStream_t st) {
    WhParams p;
    memset(&p, 0, sizeof(p));
    const int ic = g.pack_in ? g.ic / 2 : g.ic, oc = g.pack_out ? g.oc / 2 : g.oc;      // per launch (row-phase half)
    p.B = g.b; p.H = g.in_h; p.W = g.in_w; p.IC = ic; p.OC = oc; p.k = g.kh; p.pad0 = g.pad0;
    p.IC_total = g.ic; p.OC_total = g.oc; p.x_packed = g.pack_in; p.gy_packed = g.pack_out;
    p.per_sample = g.w_per_sample; p.gw = gw; p.f16 = g.f16;
    p.tiles_h = (g.out_h + kWhTH - 1) / kWhTH;
    p.tiles_w = (g.out_w + kWhTW - 1) / kWhTW;
    p.tiles_per_img = p.tiles_h * p.tiles_w;
    p.div_tpi = make_fastdiv((uint32_t)p.tiles_per_img);
    p.div_tw = make_fastdiv((uint32_t)p.tiles_w);
    const int wbs = g.w_per_sample ? g.b : 1;
    p.kt_total = g.w_per_sample ? p.tiles_per_img : p.tiles_per_img * g.b;
    int splits = (3 * sm_count() + wbs - 1) / wbs;
    int max_splits = (p.kt_total + 7) / 8;                 // >= 8 pixel tiles per unit
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.kt_per_unit = (p.kt_total + splits - 1) / splits;
    p.units_per_wb = (p.kt_total + p.kt_per_unit - 1) / p.kt_per_unit;
    p.units = wbs * p.units_per_wb;
    p.rowb_m = ic * 2; p.layout_m = p.rowb_m == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.rowb_n = oc * 2; p.layout_n = p.rowb_n == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.PW = g.kw == 1 ? 8 : 16;
    p.RH = kWhTH + g.kh - 1;
    p.a_slot_bytes = p.RH * p.PW * p.rowb_m;
    // the last ignored atom of a group reads up to (k-1 + 1) pixels past the tile row and one K step may run
    // past the last buffer row by nothing (K rows = 16 tile rows <= RH): keep one spare KiB of slack per slot
    p.a_slot_bytes = ((p.a_slot_bytes + 1023) & ~1023);
    p.b_slot_bytes = 128 * p.rowb_n;
    p.atoms_per_group = g.kh == 1 ? 1 : (128 / ic >= 3 ? 3 : 2);
    p.ngroups = g.kh == 1 ? 1 : g.kh * ((g.kw + p.atoms_per_group - 1) / p.atoms_per_group);
    p.issuers = p.ngroups < 3 ? p.ngroups : 3;
    if (const char* e = getenv("B200GAN_WGRAD_ISSUERS")) {          // timing experiments: 1 = the single-issuer kernel
        const int v = atoi(e);
        if (v >= 1 && v <= 3 && v <= p.ngroups) p.issuers = v;
    }
    int cols = 32;
    while (cols < p.ngroups * p.OC) cols <<= 1;
    if (cols > 512) return B200GAN_ENOSUP;
    p.tmem_cols = cols;
    p.stages = (int)((190 * 1024) / (p.a_slot_bytes + p.b_slot_bytes));
    if (p.stages > 4) p.stages = 4;
    if (p.stages < 2) return B200GAN_ENOSUP;
    if (p.a_slot_bytes != p.RH * p.PW * p.rowb_m) {
        set_error("conv_wgrad_halo: slot/box size mismatch");
        return B200GAN_ENOSUP;
    }
    // plain operand: (b, h, w, c); packed operand: physical (b, 2h, 2w, c/4) seen as (b, h, py, w, [px, c/4])
    auto make_map = [&](CUtensorMap* m, const void* ptr, bool packed, int ch, int w, int h, int box_w, int box_h,
                        int row_bytes) -> int {
        if (packed) {
            uint64_t dims[5] = {(uint64_t)ch, (uint64_t)w, 2, (uint64_t)h, (uint64_t)g.b};
            uint64_t strides[4] = {(uint64_t)ch * 2, (uint64_t)w * ch * 2, 2 * (uint64_t)w * ch * 2,
                                   (uint64_t)h * 2 * w * ch * 2};
            uint32_t box[5] = {(uint32_t)ch, (uint32_t)box_w, 1, (uint32_t)box_h, 1};
            uint32_t es[5] = {1, 1, 1, 1, 1};
            return encode_bf16_map(m, ptr, 5, dims, strides, box, es, row_bytes);
        }
        uint64_t dims[4] = {(uint64_t)ch, (uint64_t)w, (uint64_t)h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)ch * 2, (uint64_t)w * ch * 2, (uint64_t)h * w * ch * 2};
        uint32_t box[4] = {(uint32_t)ch, (uint32_t)box_w, (uint32_t)box_h, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        return encode_bf16_map(m, ptr, 4, dims, strides, box, es, row_bytes);
    };
    CUtensorMap map_x, map_gy;
    if (int e = make_map(&map_x, x, g.pack_in, ic, g.in_w, g.in_h, p.PW, p.RH, p.rowb_m)) return e;
    if (int e = make_map(&map_gy, gy, g.pack_out, oc, g.out_w, g.out_h, kWhTW, kWhTH, p.rowb_n)) return e;
    const size_t smem = 1024 + (size_t)p.stages * (p.a_slot_bytes + p.b_slot_bytes) + 2048 + (2 * p.stages + 2) * sizeof(uint64_t) + 64;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(conv_wgrad_halo_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv_wgrad_halo_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv_wgrad_halo_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_dev = cur_dev;
    }
    int grid = p.units < sm_count() ? p.units : sm_count();
    for (int xp = 0; xp < (g.pack_in ? 2 : 1); ++xp)
        for (int yp = 0; yp < (g.pack_out ? 2 : 1); ++yp) {
            p.x_phase = xp; p.gy_phase = yp;
            p.ic_off = xp * ic; p.oc_off = yp * oc;
            if (g.kh == 1)                    conv_wgrad_halo_kernel<1, 1><<<grid, kWhThreads, smem, st>>>(map_x, map_gy, p);
            else if (p.atoms_per_group == 3)  conv_wgrad_halo_kernel<3, 3><<<grid, kWhThreads, smem, st>>>(map_x, map_gy, p);
            else                              conv_wgrad_halo_kernel<3, 2><<<grid, kWhThreads, smem, st>>>(map_x, map_gy, p);
            count_launch();
        }
    return check_launch("conv_wgrad_halo");
}

}  // namespace b200gan
