// Blackwell (sm_100a) building blocks used by the tcgen05 implicit-GEMM kernels: mbarrier, TMA
// (cp.async.bulk.tensor), TMEM allocation, tcgen05.mma / commit / ld, and the shared-memory matrix
// descriptors.  Raw PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is resolved at run time)

#include "common.cuh"

namespace b200gan {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug becomes a trapped kernel (an error code at the next sync) instead
// of a hung GPU.  try_wait itself suspends for a hardware time slice, so the bound is seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 26)) {
            printf("b200gan: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- TMEM / tcgen05 -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {          // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, one CTA.  Issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Lean issue path: the single MMA-issuing thread is instruction-bound (profiles/r01: ~200 clk per
// tcgen05.mma when the descriptors are rebuilt every time), so the 64-bit descriptors are kept as a
// constant high word plus a 32-bit low word (start address >> 4 | LBO << 16) that is advanced by adds.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout, uint32_t base_offset = 0) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((base_offset & 7u) << 17) | (layout << 29);
}
template <bool ACC>
__device__ __forceinline__ void mma_issue(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (ACC)
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.eq.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
            : "memory");
}
__device__ __forceinline__ void mma_issue_dyn(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive on an mbarrier once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t <-> lane t of the warp's quadrant)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 16-bit storage: bfloat16 or (f16 != 0) IEEE half ----------------------------------------------------------------
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 16 floats -> 8 packed words; the storage-format test is ONE warp-uniform branch around the eight conversions (inside
// pack16x2 the compiler evaluates both conversions per pair and selects: 16 F2FP + 8 SEL per 16 channels instead of 8 F2FP)
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&pk)[8], int f16) {
    if (f16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
}
__device__ __forceinline__ float lo16(uint32_t w, int f16) {
    return f16 ? __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))) : __uint_as_float(w << 16);
}
__device__ __forceinline__ float hi16(uint32_t w, int f16) {
    return f16 ? __half2float(__ushort_as_half((unsigned short)(w >> 16))) : __uint_as_float(w & 0xffff0000u);
}

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two lanes of fp32 math) ---------------------------
// The epilogue warps share their scheduler with the MMA-issuing / TMA warps, whose issue latency is the kernel's critical
// path on the short-K layers (profiles/r02_epilogue_cost.md: every epilogue instruction costs ~1 clk of kernel time).
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// ---- epilogue side inputs: 16 consecutive bf16 of an output-shaped tensor (addend / gate, include/b200gan.h) ----------
struct Side16 {
    uint4 a, b;
};
__device__ __forceinline__ Side16 side_load16(const __nv_bfloat16* p) {
    Side16 s;
    s.a = __ldg(reinterpret_cast<const uint4*>(p));
    s.b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    return s;
}
__device__ __forceinline__ void side_unpack16(const Side16& s, float (&o)[16], int f16) {
    const uint32_t w[8] = {s.a.x, s.a.y, s.a.z, s.a.w, s.b.x, s.b.y, s.b.z, s.b.w};
    if (f16) {                      // one warp-uniform branch, not a select per element (see pack16)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            o[2 * e] = lo16(w[e], 1);
            o[2 * e + 1] = hi16(w[e], 1);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            o[2 * e] = lo16(w[e], 0);
            o[2 * e + 1] = hi16(w[e], 0);
        }
    }
}
// v = (v + addend) [gate mode: * r * (gate > 0 ? g : gs)]; the forward activation is applied by the caller otherwise
__device__ __forceinline__ void side_apply16(float (&v)[16], const Side16* addend, const Side16* gate, const float (&r)[16],
                                             float g, float gs, int f16) {
    if (addend) {
        float a[16];
        side_unpack16(*addend, a, f16);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += a[e];
    }
    if (gate) {
        float y[16];
        side_unpack16(*gate, y, f16);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] *= r[e] * (y[e] > 0.f ? g : gs);
    }
}

// ---- descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address [0,14),
// LBO [16,30), SBO [32,46) (all >> 4), version = 1 at [46,48), layout type at [61,64).
enum : uint32_t { LAYOUT_SW128 = 2, LAYOUT_SW64 = 4 };
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A/B.
__host__ __device__ inline uint32_t instr_desc_bf16(int m, int n, int a_mn_major, int b_mn_major, int f16 = 0) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    if (!f16) {                         // kind::f16 operand formats: 0 = F16, 1 = BF16
        d |= 1u << 7;                   // a_format = BF16
        d |= 1u << 10;                  // b_format = BF16
    }
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

}  // namespace umma

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder();
// bf16 tensor, `rank` dims (innermost first), byte strides for dims 1..rank-1, box and element strides.
int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes);

}  // namespace b200gan
