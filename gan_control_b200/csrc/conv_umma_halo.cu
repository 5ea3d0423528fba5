// tcgen05 implicit-GEMM convolution, "halo" variant for the small-channel / high-resolution layers
// (3x3 or 1x1, stride 1, IC and OC <= 64: the 1024^2 and 512^2 StyledConv / ResBlock convs and their
// data gradients).  The general kernel (conv_umma.cu) fetches one shifted A tile per tap -- 9x the
// activation traffic from L2 and 9 TMA round trips per tile; at 32..64 channels that, not the tensor
// pipe, bounds it (profiles/: 7 % of bf16 peak, 13 % of HBM at 32->32 @1024^2).  Here:
//   * the input region of a 16x8-pixel output tile, (16+k-1) x 16 pixels, is fetched ONCE per channel
//     chunk; the 9 tap operands are descriptors into that one buffer, shifted by whole pixel rows
//     (ky * 2048 B) and by kx pixels (kx * 128 B), 8-pixel swizzle atoms = one tile row, stride between atoms = the buffer's
//     16-pixel row pitch (measured on B200: the hardware swizzles on absolute address bits, the
//     descriptor's base-offset field must stay 0 -- tests/test_kernels_gpu.py::test_conv_fwd_halo);
//   * all taps of the (per-sample) weights stay resident in shared memory while consecutive tiles use
//     the same sample, so steady-state traffic per tile is the activation halo only.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "conv.cuh"
#include "umma.cuh"

namespace b200gan {

using namespace umma;

#ifndef B200GAN_F32X2
#define B200GAN_F32X2 1      // packed fp32 pairs in the epilogue math (umma.cuh)
#endif

#ifndef B200GAN_HALO_DEBUG
#define B200GAN_HALO_DEBUG 0     // 1: the run-time switches of HaloParams::dbg are compiled in (they sit in the issue-critical loops)
#endif
constexpr bool kHaloDebug = B200GAN_HALO_DEBUG != 0;

constexpr int kHaloThreads = 384;            // warp 0: TMA, 1-3: MMA issuers (2 also allocates TMEM), 4-7 and 8-11: two epilogue groups
constexpr int kHaloIssuers = 3;              // MMA-issuing warps (1, 2, 3), tiles round-robin
constexpr int kHaloAcc = 6;                  // max TMEM accumulators in flight (HaloParams::nacc in use; tiles alternate between the epilogue groups)
constexpr int kHTW = 8, kHTH = 16;

struct HaloParams {
    int B, H, W, OC, IC;                 // stride-1 conv: output extent == (H + 2*pad - k + 1) passed as OH/OW
    int OH, OW, k, pad0;
    int tiles_h, tiles_w, total_tiles;
    FastDiv div_img, div_tw;             // tile -> (sample, tile in image), tile in image -> (tile row, tile column)
    int BN, kchunks, kc, row_bytes, layout;
    int PW, RH;                          // buffer row pitch (pixels) and rows
    int w_per_sample;
    int pack_in, pack_out;               // space-to-depth views (ConvGeom): 5-D activation map / depth-to-space store
    int cpp;                             // pack_in: channel chunks per row phase py
    int CQ;                              // pack_out: physical output channels (= OC / 4)
    int a_stages, a_stage_bytes, w_tile_bytes, w_bytes;
    // Tiles go round-robin over `issuers` MMA warps, `nacc` accumulators and the two epilogue groups.  mbarrier waits
    // only see a phase PARITY, so every waiter must observe each barrier's phases one by one: issuers and the epilogue
    // group count (2) both divide nacc (an accumulator always belongs to the same issuer and the same group), and
    // issuers <= a_stages / kchunks (the previous fill of a stage is then complete when a warp returns to it).
    int issuers, nacc;
    int tmem_cols;
    const float* bias;
    const float* rowscale;
    const __nv_bfloat16* noise;
    const float* noise_w;
    float slope, gain;
    int has_ep;
    int f16;                             // 16-bit storage is IEEE half
    const __nv_bfloat16* addend;         // output-shaped side inputs (include/b200gan.h b200gan_conv_epilogue)
    const __nv_bfloat16* gate;
    int dbg;                             // timing experiments only (build with -DB200GAN_HALO_DEBUG=1, then env B200GAN_HALO_DEBUG):
                                         // 1 no MMA, 2 no TMA loads, 4 no stores, 8 no tcgen05.ld
    __nv_bfloat16* y;
};

// 16 accumulator columns of one pixel -> epilogue -> 32 bytes of bf16.  r / b: this chunk's demodulation scales and
// biases, already in registers.
template <bool ROW, bool SIDE>
__device__ __forceinline__ void epilogue_math_store16(const HaloParams& p, float (&v)[16], __nv_bfloat16* dst,
                                                      const float (&r)[16], const float (&b)[16], float nz) {
    if (p.has_ep && !(SIDE && p.gate != nullptr)) {
        // gain * lrelu(u) = max(g*u, g*slope*u) for gain > 0, 0 <= slope <= 1 (every use on the path): FFMA, FMUL, FMNMX
        const float g = p.gain, gs = p.gain * p.slope;
        if (g > 0.f && p.slope >= 0.f && p.slope <= 1.f) {
            if (B200GAN_F32X2 && p.rowscale == nullptr) {
                // two channels per instruction: (nz + b), (v + .), * g, * gs as packed pairs; only the max stays scalar.
                // Measured (scripts/epilogue_cost.py, bench.py): 32 -> 32 @1024^2 with noise + bias 0.700 -> 0.687 ms,
                // 64 -> 64 @512^2 0.355 -> 0.350 ms.  With a demodulation row (FFMA2 form) it was SLOWER (0.72 -> 0.94 ms),
                // so that (now unused: the demodulation is folded into the weights) case keeps the scalar code.
                const F2 nz2 = f2_pack(nz, nz), g2 = f2_pack(g, g), gs2 = f2_pack(gs, gs);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const F2 u2 = f2_add(f2_pack(v[e], v[e + 1]), f2_add(f2_pack(b[e], b[e + 1]), nz2));
                    float t0, t1, s0, s1;
                    f2_unpack(f2_mul(u2, g2), t0, t1);
                    f2_unpack(f2_mul(u2, gs2), s0, s1);
                    v[e] = fmaxf(t0, s0);
                    v[e + 1] = fmaxf(t1, s1);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float u = ROW ? fmaf(v[e], r[e], nz + b[e]) : v[e] + (nz + b[e]);
                    v[e] = fmaxf(u * g, u * gs);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float u = ROW ? fmaf(v[e], r[e], nz + b[e]) : v[e] + (nz + b[e]);
                v[e] = g * (u > 0.f ? u : u * p.slope);
            }
        }
    }
    uint32_t pk[8];
    pack16(v, pk, p.f16);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
// 16 floats through four 16-byte loads (global via the read-only path, or shared memory: warp-uniform = broadcast)
template <bool GLOBAL>
__device__ __forceinline__ void load16(const float* src, float (&o)[16], float fill) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float4 t = make_float4(fill, fill, fill, fill);
        if (src) t = GLOBAL ? __ldg(reinterpret_cast<const float4*>(src) + e) : *(reinterpret_cast<const float4*>(src) + e);
        o[4 * e] = t.x; o[4 * e + 1] = t.y; o[4 * e + 2] = t.z; o[4 * e + 3] = t.w;
    }
}

// KDIM = 3 or 1 (kernel size), ROWB = 128 or 64 (bytes per pixel row of a channel chunk = swizzle width): both
// compile-time so that the MMA-issuing warp's tap loop is straight-line code with immediate descriptor offsets.
// SIDE: the epilogue reads output-shaped side inputs (addend / gate).  A template parameter because the epilogue warps'
// instruction count is on the critical path of the 32-channel layers (profiles/r02_epilogue_cost.md): the plain forward
// kernel must not carry the side-input code (it cost 0.72 -> 0.82 ms when it was a run-time branch).
template <int KDIM, int ROWB, bool SIDE>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_fwd_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* w_buf = smem;                                             // [tap][chunk][BN][row_bytes]
    uint8_t* a_buf = w_buf + ((p.w_bytes + 1023) & ~1023);
    uint64_t* afull = reinterpret_cast<uint64_t*>(a_buf + p.a_stages * p.a_stage_bytes);
    uint64_t* aempty = afull + p.a_stages;
    uint64_t* wfull = aempty + p.a_stages;
    uint64_t* tfull = wfull + 1;
    uint64_t* tempty = tfull + kHaloAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kHaloAcc);

    __shared__ __align__(16) float s_bias[128];          // physical output channels (<= 128)
    __shared__ __align__(16) float s_rs[2][128];         // demodulation row of the sample each epilogue group is on
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_w);
    }
    {
        const int ch = p.pack_out ? p.CQ : p.OC;
        if ((int)threadIdx.x < ch) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(afull + s, 1); mbar_init(aempty + s, 1); }
        mbar_init(wfull, 1);
        for (int a = 0; a < kHaloAcc; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int taps = KDIM * KDIM;

    if (warp == 0) {
        // ================= TMA producer (whole warp, elected lane issues) =================
        int stage = 0, par = 0, key = -1, it = 0;
        int last_stage[kHaloIssuers], last_par[kHaloIssuers];       // last stage of the most recent tiles (one per MMA warp)
#pragma unroll
        for (int k = 0; k < kHaloIssuers; ++k) { last_stage[k] = -1; last_par[k] = 0; }
        int hist = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            uint32_t n, t, th, tw;
            p.div_img.divmod((uint32_t)tile, n, t);
            p.div_tw.divmod(t, th, tw);
            const int h0 = (int)th * kHTH, w0 = (int)tw * kHTW;
            const int wkey = p.w_per_sample ? (int)n : 0;
            if (wkey != key) {
                // drain: every MMA that reads the resident weights has completed once the last activation stage of
                // each of the most recent tiles (one per MMA-issuing warp: a commit covers its own thread's MMAs)
                // has been released
#pragma unroll
                for (int k = 0; k < kHaloIssuers; ++k)
                    if (last_stage[k] >= 0) mbar_wait(aempty + last_stage[k], last_par[k]);
                if (elect_one()) {
                    mbar_arrive_expect_tx(wfull, (uint32_t)p.w_bytes);
                    for (int tp = 0; tp < taps; ++tp)
                        for (int c = 0; c < p.kchunks; ++c)
                            tma_load_3d(w_buf + (tp * p.kchunks + c) * p.w_tile_bytes, &map_w, wfull, c * p.kc, 0, wkey * taps + tp);
                }
                __syncwarp();
                key = wkey;
            }
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(aempty + stage, par ^ 1);
                if (kHaloDebug && (p.dbg & 2)) {
                    if (elect_one()) mbar_arrive(afull + stage);
                } else if (elect_one()) {
                    mbar_arrive_expect_tx(afull + stage, (uint32_t)p.a_stage_bytes);
                    if (p.pack_in)      // chunk c = (row phase py, channel chunk of the 2*C contiguous (px, c) elements)
                        tma_load_5d(a_buf + stage * p.a_stage_bytes, &map_x, afull + stage, (c % p.cpp) * p.kc, w0 - p.pad0,
                                    c / p.cpp, h0 - p.pad0, (int)n);
                    else
                        tma_load_4d(a_buf + stage * p.a_stage_bytes, &map_x, afull + stage, c * p.kc, w0 - p.pad0, h0 - p.pad0, (int)n);
                }
                __syncwarp();
                // this stage's previous use was released before the refill above, so a recorded (stage, parity) of that
                // use is stale -- waiting for its parity again would alias a LATER phase of the same barrier and hang
#pragma unroll
                for (int k = 0; k < kHaloIssuers; ++k) {
                    if (last_stage[k] == stage) last_stage[k] = -1;
                    if (k == hist) { last_stage[k] = stage; last_par[k] = par; }
                }
                if (++stage == p.a_stages) { stage = 0; par ^= 1; }
            }
            if (++hist == p.issuers) hist = 0;
        }
    } else if (warp >= 1 && warp <= p.issuers) {
        // ================= kHaloIssuers MMA issuers (whole warp, elected lane issues), tiles round-robin =================
        // The per-tile protocol of an issuing warp (tile decode, two mbarrier waits + fences, elect, two commits) costs
        // ~600 clk during which the 8-deep tcgen05 queue drains (profiles/r01_umma_pacing.md: the MMA stream alone took
        // 1408 clk per tile for 720 clk of pipe time).  Several warps issue tiles round-robin, so one warp's protocol
        // overlaps the others' MMAs (measured 32->32 @1024^2: 0.66 ms with one issuer, 0.56 ms with two, 0.48 ms with three).
        // All walk every tile (weight-buffer phases and the stage ring are sequential) but issue only their own.
        const int mine = warp - 1;
        int turn = 0;
        // ncu on the first version (run-time tap loop): this ONE warp bounds the kernel -- 266 dependent, mostly
        // uniform-datapath instructions per tile at ~7 clk each = the whole 1820-clk tile period, tensor pipe 17 %
        // active, the TMA producer and the epilogue warps waiting on it (profiles/r01_conv_issue_bound.md, r01_ncu_kernels.md).  Now the
        // taps are unrolled with immediate descriptor offsets and the tile decode is a multiply-high.
        const uint32_t idesc = instr_desc_bf16(128, p.BN, 0, 0, p.f16);
        constexpr uint32_t PW = KDIM == 1 ? 8 : 16;
        constexpr uint32_t ROW_UNITS = ROWB >> 4;                        // descriptor units (16 B) per pixel
        constexpr int KSTEPS = ROWB / 32;                                // K = 16 elements = 32 B per MMA
        // A: one 8-pixel swizzle atom per tile row, atoms PW pixels apart; the swizzle is a function of the
        // absolute shared-memory address, so tap-shifted start addresses need no base-offset field
        const uint32_t a_hi = desc_hi(PW * ROWB, (uint32_t)p.layout);
        const uint32_t b_hi = desc_hi(8u * ROWB, (uint32_t)p.layout);
        const uint32_t w_lo0 = desc_lo(smem_u32(w_buf), 16), w_inc = (uint32_t)p.w_tile_bytes >> 4;
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), 16), a_inc = (uint32_t)p.a_stage_bytes >> 4;
        const int nstages = p.a_stages, kchunks = p.kchunks, bn = p.BN, total = p.total_tiles;
        const uint32_t w_tap = w_inc * (uint32_t)kchunks;               // tiles are [tap][chunk]
        int stage = 0, par = 0, key = -1, wpar = 0;
        // this warp's tiles are iterations mine, mine + issuers, ...: their accumulator (iteration % nacc) and barrier parity
        // ((iteration / nacc) & 1) advance by `issuers` with a wrap -- no integer division in the issuing warp (issuers | nacc)
        int my_acc = mine, my_par = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int wkey = p.w_per_sample ? (int)p.div_img.quot((uint32_t)tile) : 0;
            if (wkey != key) {
                mbar_wait(wfull, wpar);
                wpar ^= 1;
                key = wkey;
            }
            const bool skip = turn != mine;
            if (++turn == p.issuers) turn = 0;
            if (skip) {                                 // another warp's tile: only advance the stage ring
                for (int c = 0; c < kchunks; ++c)
                    if (++stage == nstages) { stage = 0; par ^= 1; }
                continue;
            }
            const int acc = my_acc, acc_par = my_par;
            my_acc += p.issuers;
            if (my_acc >= p.nacc) { my_acc -= p.nacc; my_par ^= 1; }
            mbar_wait(tempty + acc, acc_par ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * bn);
            for (int c = 0; c < kchunks; ++c) {
                mbar_wait(afull + stage, par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = a_lo0 + (uint32_t)stage * a_inc;
                    const uint32_t w_base = w_lo0 + (uint32_t)c * w_inc;
#pragma unroll
                    for (int tap = 0; tap < taps; ++tap) {
                        if (kHaloDebug && (p.dbg & 1)) break;
                        const uint32_t a_lo = a_base + (uint32_t)((tap / KDIM) * PW + (tap % KDIM)) * ROW_UNITS;
                        const uint32_t w_lo = w_base + (uint32_t)tap * w_tap;
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ++ks) {
                            if (tap == 0 && ks == 0)
                                mma_issue_dyn(d_tmem, a_lo, a_hi, w_lo, b_hi, idesc, (uint32_t)(c != 0));
                            else
                                mma_issue<true>(d_tmem, a_lo + 2 * ks, a_hi, w_lo + 2 * ks, b_hi, idesc);
                        }
                    }
                    mma_commit(aempty + stage);
                    if (c == kchunks - 1) mma_commit(tfull + acc);
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; par ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: two groups of four warps, tiles alternate between them =================
        // ncu (profiles/r01_ncu_kernels.md, 32->32 @1024^2): with ONE group and two accumulators the kernel ran at
        // 1340 clk per tile against 720 clk of HBM time: the epilogue warps were busy ~58 % of it (64 scalar side loads
        // per lane, the noise load's DRAM latency exposed after the accumulator wait) and the MMA warp could not run
        // more than one tile ahead.  Now: four accumulators, two groups, noise fetched BEFORE the wait, vector side loads.
        const int q = warp & 3, group = (warp - 4) >> 2;      // TMEM lane quadrant = warp % 4
        const int r = q * 32 + lane;
        const int w_l = r % kHTW, h_l = r / kHTW;
        const float nw = (p.noise != nullptr && p.noise_w != nullptr) ? *p.noise_w : 0.f;
        // ncu (profiles/r02_ncu_halo_epilogue.md): the per-pixel noise value was loaded before the accumulator wait but
        // CONSUMED there too (nw * noise), so the warp sat on the DRAM latency of that 2-byte load every tile: 25 % of
        // all stall samples of the kernel on one instruction, 0.15 ms of 0.72 ms.  Now the raw bits of the NEXT tile's
        // noise are requested while the current tile is drained and only converted after the next accumulator wait.
        uint32_t nraw0 = 0u, nraw1 = 0u;
        const int ch_phys = p.pack_out ? p.CQ : p.OC, gt = (warp & 3) * 32 + lane;      // thread within the group
        float* my_rs = s_rs[group];
        const float* bs_sm = p.bias ? s_bias : nullptr;
        int rs_n = -1;
        // Where the per-channel side inputs come from (measured, profiles/r02_epilogue_cost.md): at 64 physical output
        // channels the bias and the current sample's demodulation row are staged in shared memory (0.434 -> 0.385 ms for
        // the full StyledConv epilogue at 64 -> 64 @512^2); with <= 32 channels the MMAs are bound by the shared-memory
        // port (N <= 32: 40 clk of operand fetch per instruction) and LDS wavefronts are taken from them: the read-only
        // global path (L1 hits) is faster there (0.57 vs 0.65 ms).  A register-resident bias row spilled (168 registers).
        const bool side_smem = ch_phys > 32;
        auto chunk = [&](uint32_t taddr, int col, __nv_bfloat16* dst, int c0, const float* rs_g, float nz, bool valid) {
            float v[16], r16[16], b16[16];
            // output-shaped side inputs (the gradient another consumer contributes, the producer's saved output): their
            // loads are issued before the TMEM read so that the two latencies overlap
            Side16 s_add, s_gate;
            if (SIDE) {
                const int64_t soff = (dst - p.y) + c0;
                if (valid && p.addend) s_add = side_load16(p.addend + soff);
                if (valid && p.gate) s_gate = side_load16(p.gate + soff);
            }
            if (!(kHaloDebug && (p.dbg & 8))) tmem_ld_x16(taddr + (uint32_t)col, v);
            if (valid && !(kHaloDebug && (p.dbg & 4))) {
                if (side_smem) {
                    load16<false>(p.rowscale ? my_rs + c0 : nullptr, r16, 1.f);
                    load16<false>(bs_sm ? bs_sm + c0 : nullptr, b16, 0.f);
                } else {
                    load16<true>(rs_g ? rs_g + c0 : nullptr, r16, 1.f);
                    load16<true>(p.bias ? p.bias + c0 : nullptr, b16, 0.f);
                }
                if (SIDE && (p.addend || p.gate))
                    side_apply16(v, p.addend ? &s_add : nullptr, p.gate ? &s_gate : nullptr, r16, p.gain, p.gain * p.slope, p.f16);
                epilogue_math_store16<true, SIDE>(p, v, dst + c0, r16, b16, nz);
            }
        };
        // rowscale[n][:] -> shared memory when this group moves on to another sample (all four warps of a group walk
        // the same tiles, so they arrive here in the same iteration: a 128-thread named barrier, id 1 / 2)
        auto stage_rowscale = [&](int n) {
            if (p.rowscale == nullptr || n == rs_n) return;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");      // everyone is done reading the old row
            if (gt < ch_phys) my_rs[gt] = p.rowscale[(int64_t)n * ch_phys + gt];
            asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
            rs_n = n;
        };
        // address of this thread's noise sample(s) of `tile` (null when the tile / the pixel does not exist)
        auto noise_ptr = [&](int tile) -> const __nv_bfloat16* {
            if (p.noise == nullptr || tile >= p.total_tiles) return nullptr;
            uint32_t n, t, th, tw;
            p.div_img.divmod((uint32_t)tile, n, t);
            p.div_tw.divmod(t, th, tw);
            const int oy = (int)th * kHTH + h_l, ox = (int)tw * kHTW + w_l;
            if (oy >= p.OH || ox >= p.OW) return nullptr;
            if (!p.pack_out) return p.noise + ((int64_t)n * p.OH + oy) * p.OW + ox;
            return p.noise + ((int64_t)n * (2 * p.OH) + 2 * oy) * (2 * p.OW) + 2 * ox;
        };
        // two-level prefetch: the line is pulled into L2 four of this group's tiles ahead (no register cost), the value
        // into a register one tile ahead
        auto fetch_noise = [&](int tile) {
            if (const __nv_bfloat16* far = noise_ptr(tile + 6 * (int)gridDim.x)) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(far));
                if (p.pack_out) asm volatile("prefetch.global.L2 [%0];" ::"l"(far + 2 * p.OW));
            }
            const __nv_bfloat16* q = noise_ptr(tile);
            if (q == nullptr) return;
            if (!p.pack_out) {
                nraw0 = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(q));
            } else {
                nraw0 = __ldg(reinterpret_cast<const uint32_t*>(q));
                nraw1 = __ldg(reinterpret_cast<const uint32_t*>(q + 2 * p.OW));
            }
        };
        // the output-shaped side inputs (addend / gate) of a tile two of this group's tiles ahead are pulled into L2: the
        // per-chunk loads in `chunk` then see an L2 hit instead of a DRAM round trip in the middle of the epilogue
        auto prefetch_side = [&](int tile) {
            if (!SIDE || (p.addend == nullptr && p.gate == nullptr) || tile >= p.total_tiles) return;
            uint32_t n, t, th, tw;
            p.div_img.divmod((uint32_t)tile, n, t);
            p.div_tw.divmod(t, th, tw);
            const int oy = (int)th * kHTH + h_l, ox = (int)tw * kHTW + w_l;
            if (oy >= p.OH || ox >= p.OW) return;
            const int nph = p.pack_out ? 4 : 1;
            for (int ph = 0; ph < nph; ++ph) {
                const int64_t off = p.pack_out
                    ? ((((int64_t)n * (2 * p.OH) + 2 * oy + (ph >> 1)) * (2 * p.OW) + 2 * ox + (ph & 1)) * p.CQ)
                    : ((((int64_t)n * p.OH + oy) * p.OW + ox) * p.OC);
                const int bytes = (p.pack_out ? p.CQ : p.OC) * 2;
                for (int bo = 0; bo < bytes; bo += 128) {
                    if (p.addend) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.addend + off) + bo));
                    if (p.gate) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.gate + off) + bo));
                }
            }
        };
        int my_acc = group, my_par = 0;                  // iterations group, group + 2, ...: (iteration % nacc, parity), 2 | nacc
        const int first = blockIdx.x + group * gridDim.x;
        fetch_noise(first);
        prefetch_side(first);
        prefetch_side(first + 2 * (int)gridDim.x);
        for (int tile = first; tile < p.total_tiles; tile += 2 * gridDim.x) {
            uint32_t n, t, th, tw;
            p.div_img.divmod((uint32_t)tile, n, t);
            p.div_tw.divmod(t, th, tw);
            const int oy = (int)th * kHTH + h_l, ox = (int)tw * kHTW + w_l;
            const int acc = my_acc, acc_par = my_par;
            my_acc += 2;
            if (my_acc >= p.nacc) { my_acc -= p.nacc; my_par ^= 1; }
            const bool valid = oy < p.OH && ox < p.OW;
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.BN) + ((uint32_t)(q * 32) << 16);
            if (!p.pack_out) {
                const int64_t pix = ((int64_t)n * p.OH + oy) * p.OW + ox;
                __nv_bfloat16* dst = p.y + pix * p.OC;
                if (side_smem) stage_rowscale((int)n);
                const float* rs_g = p.rowscale ? p.rowscale + (int64_t)n * p.OC : nullptr;
                mbar_wait(tfull + acc, acc_par);
                tc_fence_after();
                const float nz = (valid && p.noise) ? nw * lo16(nraw0, p.f16) : 0.f;
                fetch_noise(tile + 2 * gridDim.x);
                prefetch_side(tile + 4 * (int)gridDim.x);
                for (int c0 = 0; c0 < p.BN; c0 += 16) chunk(taddr, c0, dst, c0, rs_g, nz, valid);
            } else {
                // depth-to-space: accumulator columns [(py*2+px)*CQ + c] of view pixel (oy, ox) are channel c of the
                // physical pixel (2*oy+py, 2*ox+px); bias / rowscale / noise follow the physical tensor
                if (side_smem) stage_rowscale((int)n);
                const float* rs_g = p.rowscale ? p.rowscale + (int64_t)n * p.CQ : nullptr;
                const int64_t pix0 = ((int64_t)n * (2 * p.OH) + 2 * oy) * (2 * p.OW) + 2 * ox;
                mbar_wait(tfull + acc, acc_par);
                tc_fence_after();
                float nz[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid && p.noise) {
                    nz[0] = nw * lo16(nraw0, p.f16); nz[1] = nw * hi16(nraw0, p.f16);
                    nz[2] = nw * lo16(nraw1, p.f16); nz[3] = nw * hi16(nraw1, p.f16);
                }
                fetch_noise(tile + 2 * gridDim.x);
                prefetch_side(tile + 4 * (int)gridDim.x);
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    __nv_bfloat16* dst = p.y + (pix0 + (ph >> 1) * (2 * p.OW) + (ph & 1)) * p.CQ;
                    for (int c0 = 0; c0 < p.CQ; c0 += 16) chunk(taddr, ph * p.CQ + c0, dst, c0, rs_g, nz[ph], valid);
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

// shared-memory plan: resident weights + >= 2 activation stages inside the 227 KB opt-in limit
static bool halo_plan(const ConvGeom& g, HaloParams& p) {
    const int min_row = g.pack_in ? (g.ic / 2) * 2 : g.ic * 2;       // bytes of the contiguous run a chunk is cut from
    p.row_bytes = min_row >= 128 ? 128 : 64;
    p.layout = p.row_bytes == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    p.kc = p.row_bytes / 2;
    p.kchunks = g.ic / p.kc;
    p.cpp = g.pack_in ? (g.ic / 2) / p.kc : 0;
    p.BN = g.oc;
    p.PW = g.kw == 1 ? 8 : 16;
    p.RH = kHTH + g.kh - 1;
    p.a_stage_bytes = p.RH * p.PW * p.row_bytes;
    if (p.a_stage_bytes % 1024 != 0) return false;                   // expect_tx counts exactly one box per stage
    p.w_tile_bytes = p.BN * p.row_bytes;
    p.w_bytes = g.kh * g.kw * p.kchunks * p.w_tile_bytes;
    const int budget = 224 * 1024 - 1024 - ((p.w_bytes + 1023) & ~1023) - 256;
    p.a_stages = budget / p.a_stage_bytes;
    if (p.a_stages > 6) p.a_stages = 6;
    return p.a_stages >= 2;
}

bool conv_fwd_halo_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y) {
    if (dtype != B200GAN_BF16 && dtype != B200GAN_F16) return false;
    if (g.up != 1 || g.down != 1 || g.kh != g.kw || (g.kh != 1 && g.kh != 3)) return false;
    if (g.out_h < kHTH || g.out_w < kHTW) return false;
    if (g.out_h != g.in_h + 2 * g.pad0 - g.kh + 1 || g.out_w != g.in_w + 2 * g.pad0 - g.kw + 1) return false;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) % 16 != 0) return false;
    if (g.pack_in || g.pack_out) {
        // view channels = 4 x physical channels; a chunk must not straddle a row phase, a 16-column epilogue group
        // must not straddle an output phase
        if (g.kh != 3) return false;
        if (g.pack_in && !(g.ic == 64 || g.ic == 128)) return false;
        if (!g.pack_in && !(g.ic == 32 || g.ic == 64)) return false;
        if (g.pack_out && !(g.oc == 64 || g.oc == 128)) return false;
        if (!g.pack_out && !(g.oc == 16 || g.oc == 32 || g.oc == 64)) return false;
    } else {
        if (!(g.ic == 32 || g.ic == 64) || !(g.oc == 16 || g.oc == 32 || g.oc == 64)) return false;
    }
    HaloParams plan;
    memset(&plan, 0, sizeof(plan));
    if (!halo_plan(g, plan)) return false;
    return tensor_map_encoder() != nullptr;
}

int conv_fwd_halo(const void* x, const void* w, void* y, const ConvGeom& g, const ConvEp& ep, cudaStream_t st) {
    const float* bias = ep.bias;
    const float* rowscale = ep.rowscale;
    const void* noise = ep.noise;
    const float* noise_w = ep.noise_w;
    const float slope = ep.slope, gain = ep.gain;
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.B = g.b; p.H = g.in_h; p.W = g.in_w; p.OC = g.oc; p.IC = g.ic; p.OH = g.out_h; p.OW = g.out_w;
    p.k = g.kh; p.pad0 = g.pad0; p.w_per_sample = g.w_per_sample;
    p.pack_in = g.pack_in; p.pack_out = g.pack_out; p.CQ = g.oc / 4;
    p.tiles_h = (g.out_h + kHTH - 1) / kHTH;
    p.tiles_w = (g.out_w + kHTW - 1) / kHTW;
    p.total_tiles = g.b * p.tiles_h * p.tiles_w;
    p.div_img = make_fastdiv((uint32_t)(p.tiles_h * p.tiles_w));
    p.div_tw = make_fastdiv((uint32_t)p.tiles_w);
    if (!halo_plan(g, p)) return B200GAN_ENOSUP;
    // see HaloParams::issuers
    const int ring_tiles = p.a_stages / p.kchunks;
    p.nacc = p.BN <= 64 ? 6 : 4;
    p.issuers = p.BN <= 64 ? kHaloIssuers : 2;
    if (p.issuers > ring_tiles) p.issuers = ring_tiles;
    if (p.issuers < 1) return B200GAN_ENOSUP;
    int cols = 32;
    while (cols < p.nacc * p.BN) cols <<= 1;
    p.tmem_cols = cols;
    p.bias = bias; p.rowscale = rowscale; p.noise = (const __nv_bfloat16*)noise; p.noise_w = noise_w;
    p.slope = slope; p.gain = gain;
    p.has_ep = ep_active(ep) ? 1 : 0;
    p.f16 = g.f16;
    p.addend = (const __nv_bfloat16*)ep.addend;
    p.gate = (const __nv_bfloat16*)ep.gate;
    p.y = (__nv_bfloat16*)y;
    if (const char* e = getenv("B200GAN_HALO_DEBUG")) p.dbg = atoi(e);

    CUtensorMap map_x, map_w;
    if (g.pack_in) {
        // physical (b, 2H, 2W, C) seen as (b, H, py, W, [px, c]): the 2*C elements of a (px, c) pair row are contiguous
        const uint64_t c2 = (uint64_t)g.ic / 2;                          // 2 * physical channels
        uint64_t dims[5] = {c2, (uint64_t)g.in_w, 2, (uint64_t)g.in_h, (uint64_t)g.b};
        uint64_t strides[4] = {c2 * 2, (uint64_t)g.in_w * c2 * 2, 2 * (uint64_t)g.in_w * c2 * 2,
                               (uint64_t)g.in_h * 2 * g.in_w * c2 * 2};
        uint32_t box[5] = {(uint32_t)p.kc, (uint32_t)p.PW, 1, (uint32_t)p.RH, 1};
        uint32_t es[5] = {1, 1, 1, 1, 1};
        if (int e = encode_bf16_map(&map_x, x, 5, dims, strides, box, es, p.row_bytes)) return e;
    } else {
        uint64_t dims[4] = {(uint64_t)g.ic, (uint64_t)g.in_w, (uint64_t)g.in_h, (uint64_t)g.b};
        uint64_t strides[3] = {(uint64_t)g.ic * 2, (uint64_t)g.in_w * g.ic * 2, (uint64_t)g.in_h * g.in_w * g.ic * 2};
        uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.PW, (uint32_t)p.RH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (int e = encode_bf16_map(&map_x, x, 4, dims, strides, box, es, p.row_bytes)) return e;
    }
    {
        const int wb = g.w_per_sample ? g.b : 1;
        uint64_t dims[3] = {(uint64_t)g.ic, (uint64_t)g.oc, (uint64_t)wb * g.kh * g.kw};
        uint64_t strides[2] = {(uint64_t)g.ic * 2, (uint64_t)g.oc * g.ic * 2};
        uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)p.BN, 1};
        uint32_t es[3] = {1, 1, 1};
        if (int e = encode_bf16_map(&map_w, w, 3, dims, strides, box, es, p.row_bytes)) return e;
    }
    const size_t smem = 1024 + ((size_t)(p.w_bytes + 1023) & ~(size_t)1023) + (size_t)p.a_stages * p.a_stage_bytes +
                        (2 * p.a_stages + 1 + 2 * kHaloAcc) * sizeof(uint64_t) + 16;
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        cudaFuncSetAttribute(conv_fwd_halo_kernel<3, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);   // + 1.5 KB static
        cudaFuncSetAttribute(conv_fwd_halo_kernel<3, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        cudaFuncSetAttribute(conv_fwd_halo_kernel<3, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);   // + 1.5 KB static
        cudaFuncSetAttribute(conv_fwd_halo_kernel<3, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        cudaFuncSetAttribute(conv_fwd_halo_kernel<1, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);   // + 1.5 KB static
        cudaFuncSetAttribute(conv_fwd_halo_kernel<1, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        cudaFuncSetAttribute(conv_fwd_halo_kernel<1, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);   // + 1.5 KB static
        cudaFuncSetAttribute(conv_fwd_halo_kernel<1, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        attr_dev = cur_dev;
    }
    int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    const bool side = p.addend != nullptr || p.gate != nullptr;
#define B200_HALO_LAUNCH(KD, RB)                                                                            \
    do {                                                                                                    \
        if (side) conv_fwd_halo_kernel<KD, RB, true><<<grid, kHaloThreads, smem, st>>>(map_x, map_w, p);    \
        else      conv_fwd_halo_kernel<KD, RB, false><<<grid, kHaloThreads, smem, st>>>(map_x, map_w, p);   \
    } while (0)
    if (g.kh == 3 && p.row_bytes == 128) B200_HALO_LAUNCH(3, 128);
    else if (g.kh == 3)                  B200_HALO_LAUNCH(3, 64);
    else if (p.row_bytes == 128)         B200_HALO_LAUNCH(1, 128);
    else                                 B200_HALO_LAUNCH(1, 64);
#undef B200_HALO_LAUNCH
    count_launch();
    return check_launch("conv_fwd_halo");
}

}  // namespace b200gan
