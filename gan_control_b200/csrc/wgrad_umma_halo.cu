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
    int units, units_per_wb, kt_per_unit, kt_total;       // split-K over pixel tiles
    int rowb_m, layout_m, rowb_n, layout_n;
    int PW, RH, a_slot_bytes, b_slot_bytes, stages;
    int ngroups, atoms_per_group;                         // MMA groups per tile, kx taps packed per group
    int tmem_cols;
    float* gw;
};

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
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
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
        n = p.per_sample ? wb : kt / p.tiles_per_img;
        const int t = kt % p.tiles_per_img;
        h0 = (t / p.tiles_w) * kWhTH;
        w0 = (t % p.tiles_w) * kWhTW;
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
                    tma_load_4d(a_buf + stage * p.a_slot_bytes, &map_x, full + stage, 0, w0 - p.pad0, h0 - p.pad0, n);
                    tma_load_4d(b_buf + stage * p.b_slot_bytes, &map_gy, full + stage, 0, w0, h0, n);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; par ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc_bf16(128, p.OC, 1, 1);
        const uint32_t a_hi = desc_hi((uint32_t)(p.PW * p.rowb_m), (uint32_t)p.layout_m);     // next tile row of pixels
        const uint32_t b_hi = desc_hi(8u * (uint32_t)p.rowb_n, (uint32_t)p.layout_n);
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), (uint32_t)p.rowb_m);                  // LBO = one pixel = next kx
        const uint32_t b_lo0 = desc_lo(smem_u32(b_buf), 16);
        const uint32_t a_inc = (uint32_t)p.a_slot_bytes >> 4, b_inc = (uint32_t)p.b_slot_bytes >> 4;
        const uint32_t row_units = (uint32_t)p.rowb_m >> 4, pw = (uint32_t)p.PW;
        const uint32_t a_kstep = 2u * pw * row_units;          // K = 16 pixels = two tile rows
        const uint32_t b_kstep = (uint32_t)p.rowb_n;           // 16 rows * row_bytes >> 4
        const int nstages = p.stages, ngroups = p.ngroups, apg = p.atoms_per_group, kdim = p.k, oc = p.OC;
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
                    int g = 0;
                    for (int ky = 0; ky < kdim; ++ky)
                        for (int kx0 = 0; kx0 < kdim; kx0 += apg, ++g) {
                            const uint32_t a_g = a_st + ((uint32_t)ky * pw + (uint32_t)kx0) * row_units;
                            const uint32_t d_tmem = tmem_base + (uint32_t)(g * oc);
                            mma_issue_dyn(d_tmem, a_g, a_hi, b_st, b_hi, idesc, acc0);
#pragma unroll
                            for (int ks = 1; ks < 8; ++ks)
                                mma_issue<true>(d_tmem, a_g + (uint32_t)ks * a_kstep, a_hi, b_st + (uint32_t)ks * b_kstep, b_hi, idesc);
                        }
                    (void)ngroups;
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
                    float* dst = p.gw + (((int64_t)wb * taps + ky * p.k + (valid ? kx : 0)) * p.OC) * p.IC + ic;
                    const uint32_t taddr = tmem_base + (uint32_t)(g * p.OC) + ((uint32_t)(q * 32) << 16);
                    for (int c0 = 0; c0 < p.OC; c0 += 16) {
                        float v[16];
                        if (have) tmem_ld_x16(taddr + (uint32_t)c0, v);
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) atomicAdd(dst + (int64_t)(c0 + e) * p.IC, v[e]);
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
    if (dtype != B200GAN_BF16) return false;
    if (g.up != 1 || g.down != 1 || g.kh != g.kw || (g.kh != 1 && g.kh != 3)) return false;
    if (!(g.ic == 32 || g.ic == 64) || !(g.oc == 32 || g.oc == 64)) return false;
    if (g.out_h < kWhTH || g.out_w < kWhTW) return false;
    if (g.out_h != g.in_h + 2 * g.pad0 - g.kh + 1 || g.out_w != g.in_w + 2 * g.pad0 - g.kw + 1) return false;
    if (((uintptr_t)x | (uintptr_t)gy) % 16 != 0) return false;
    return tensor_map_encoder() != nullptr;
}

int conv_wgrad_halo(const void* x, const void* gy, float* gw, const ConvGeom& g, cudaStream_t st) {
    WhParams p;
    memset(&p, 0, sizeof(p));
    p.B = g.b; p.H = g.in_h; p.W = g.in_w; p.IC = g.ic; p.OC = g.oc; p.k = g.kh; p.pad0 = g.pad0;
    p.per_sample = g.w_per_sample; p.gw = gw;
    p.tiles_h = (g.out_h + kWhTH - 1) / kWhTH;
    p.tiles_w = (g.out_w + kWhTW - 1) / kWhTW;
    p.tiles_per_img = p.tiles_h * p.tiles_w;
    const int wbs = g.w_per_sample ? g.b : 1;
    p.kt_total = g.w_per_sample ? p.tiles_per_img : p.tiles_per_img * g.b;
    int splits = (3 * sm_count() + wbs - 1) / wbs;
    int max_splits = (p.kt_total + 7) / 8;                 // >= 8 pixel tiles per unit
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.kt_per_unit = (p.kt_total + splits - 1) / splits;
    p.units_per_wb = (p.kt_total + p.kt_per_unit - 1) / p.kt_per_unit;
    p.units = wbs * p.units_per_wb;
    p.rowb_m = g.ic * 2; p.layout_m = p.rowb_m == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.rowb_n = g.oc * 2; p.layout_n = p.rowb_n == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.PW = g.kw == 1 ? 8 : 16;
    p.RH = kWhTH + g.kh - 1;
    p.a_slot_bytes = p.RH * p.PW * p.rowb_m;
    // the last ignored atom of a group reads up to (k-1 + 1) pixels past the tile row and one K step may run
    // past the last buffer row by nothing (K rows = 16 tile rows <= RH): keep one spare KiB of slack per slot
    p.a_slot_bytes = ((p.a_slot_bytes + 1023) & ~1023);
    p.b_slot_bytes = 128 * p.rowb_n;
    p.atoms_per_group = g.kh == 1 ? 1 : (128 / g.ic >= 3 ? 3 : 2);
    p.ngroups = g.kh == 1 ? 1 : g.kh * ((g.kw + p.atoms_per_group - 1) / p.atoms_per_group);
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
    CUtensorMap map_x, map_gy;
    {
        uint64_t dims[4] = {(uint64_t)g.ic, (uint64_t)g.in_w, (uint64_t)g.in_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.ic * 2, (uint64_t)g.in_w * g.ic * 2, (uint64_t)g.in_h * g.in_w * g.ic * 2};
        uint32_t box[4] = {(uint32_t)g.ic, (uint32_t)p.PW, (uint32_t)p.RH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, p.rowb_m)) return e;
    }
    {
        uint64_t dims[4] = {(uint64_t)g.oc, (uint64_t)g.out_w, (uint64_t)g.out_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.oc * 2, (uint64_t)g.out_w * g.oc * 2, (uint64_t)g.out_h * g.out_w * g.oc * 2};
        uint32_t box[4] = {(uint32_t)g.oc, (uint32_t)kWhTW, (uint32_t)kWhTH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (int e = encode_bf16_map(&map_gy, gy, 4, dims, strides, box, es, p.rowb_n)) return e;
    }
    const size_t smem = 1024 + (size_t)p.stages * (p.a_slot_bytes + p.b_slot_bytes) + 2048 + (2 * p.stages + 2) * sizeof(uint64_t) + 64;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_dev = cur_dev;
    }
    int grid = p.units < sm_count() ? p.units : sm_count();
    conv_wgrad_halo_kernel<<<grid, kWhThreads, smem, st>>>(map_x, map_gy, p);
    count_launch();
    return check_launch("conv_wgrad_halo");
}

}  // namespace b200gan
