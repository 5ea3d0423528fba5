// Pacing of tcgen05.mma (SS mode, bf16, M = 128, K = 16) on B200 as a function of N, of how many TMEM
// accumulators the instruction stream alternates between, and of the operand layout.  Decides the tiling of the
// small-channel convolution kernels (profiles/r01_umma_pacing.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gan_control_b200/build/umma_pacing scripts/umma_pacing.cu
#include <cstdio>
#include <cstdlib>

#include "../gan_control_b200/csrc/umma.cuh"

using namespace b200gan::umma;

// mode 0: K-major A and B, SW128 (rows of 128 B, 8-row atoms 1024 B apart): the forward conv's operands
// mode 1: K-major, SW64 (rows of 64 B): the 32-channel forward conv
// mode 2: MN-major A and B, SW128 (the weight-gradient kernels)
__global__ void __launch_bounds__(128, 1) pacing_kernel(int n, int nacc, int mode, int iters, int a_shift, int sbo_mul, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* a_buf = smem;                    // 64 KB
    uint8_t* b_buf = smem + 64 * 1024;        // 64 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (warp == 0) {
        const int rowb = mode == 1 ? 64 : 128;
        const uint32_t layout = mode == 1 ? LAYOUT_SW64 : LAYOUT_SW128;
        uint32_t idesc, a_hi, b_hi, a_lo0, b_lo0;
        if (mode == 2) {
            idesc = instr_desc_bf16(128, n, 1, 1);
            // MN-major SW128: atom = 64 MN elements (128 B) x 8 K rows; LBO = next atom along MN, SBO = next 8 K rows
            a_hi = desc_hi(2048u, layout);
            b_hi = desc_hi(4096u, layout);
            a_lo0 = desc_lo(smem_u32(a_buf), 1024);
            b_lo0 = desc_lo(smem_u32(b_buf), 1024);
        } else {
            idesc = instr_desc_bf16(128, n, 0, 0);
            a_hi = desc_hi(8u * rowb * sbo_mul, layout);     // sbo_mul = 2: the halo kernels' gapped atoms (one per 16-pixel buffer row)
            b_hi = desc_hi(8u * rowb, layout);
            a_lo0 = desc_lo(smem_u32(a_buf), 16);
            b_lo0 = desc_lo(smem_u32(b_buf), 16);
        }
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
            __syncwarp();
            t0 = clock64();
            if (elect_one()) {
                const uint32_t kmask = mode == 1 ? 1u : 3u, px = (uint32_t)(rowb >> 4), am = (uint32_t)(nacc - 1);
#pragma unroll 8
                for (uint32_t i = 0; i < (uint32_t)iters; ++i) {
                    const uint32_t d = tmem_base + (i & am) * (uint32_t)n;
                    // a_shift: step the A start address like the conv kernels do (k-step inside the row, then pixel rows)
                    const uint32_t ks = (i & kmask) * 2;
                    const uint32_t a_lo = a_lo0 + (a_shift ? ks + ((i >> 2) & 7u) * px : 0u);
                    mma_issue_dyn(d, a_lo, a_hi, b_lo0 + ks, b_hi, idesc, (uint32_t)(i >= (uint32_t)nacc));
                }
                mma_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(pacing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2048;
    printf("tcgen05.mma kind::f16 M=128 K=16 SS: clocks per instruction (one CTA per SM, 148 CTAs, %d instructions)\n", iters);
    printf("%-28s %6s %6s %8s %10s %10s\n", "mode", "N", "nacc", "a_shift", "clk/mma", "floor");
    const char* names[3] = {"K-major SW128", "K-major SW64", "MN-major SW128"};
    for (int mode = 0; mode < 3; ++mode)
        for (int n : {16, 32, 64, 128, 256})
            for (int nacc : {1, 4})
                for (int a_shift : {0, 1, 2, 3}) {           // bit 1: gapped A atoms (SBO x2)
                    if (nacc * n > 512) continue;
                    if (mode == 2 && a_shift) continue;
                    for (int grid : {148}) {
                        pacing_kernel<<<grid, 128, 130 * 1024 + 2048>>>(n, nacc, mode, iters, a_shift & 1, (a_shift & 2) ? 2 : 1, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        long long h[148];
                        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                        printf("%-28s %6d %6d %8d %10.1f %10.1f  (grid %d)\n", names[mode], n, nacc, a_shift, (double)mx / iters,
                               128.0 * n / 256.0, grid);
                        fflush(stdout);
                    }
                }
    return 0;
}
