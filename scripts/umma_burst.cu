// Follow-up to umma_pacing.cu: cost of the per-tile protocol around bursts of tcgen05.mma.  One warp issues
// `burst` MMAs (M=128, N, K=16, SS) into accumulator (t % nacc), then tcgen05.commit -> mbarrier[t % 8]; with
// `dep` it first waits for the commit of burst t - lag (the accumulator / stage reuse dependency of the conv kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gan_control_b200/build/umma_burst scripts/umma_burst.cu
#include <cstdio>
#include <cstdlib>

#include "../gan_control_b200/csrc/umma.cuh"

using namespace b200gan::umma;

// variant bits: 1 = no commit between bursts (only the last), 2 = never overwrite (always accumulate), 4 = two commits per burst,
// 8 = lane 0 instead of elect.sync, 16 = no __syncwarp per tile, 32 = 18 MMAs as straight-line code (immediate offsets)
__global__ void __launch_bounds__(128, 1) burst_kernel(int n, int burst, int tiles, int lag, int whole_warp, int variant, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = smem + 64 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 1024);      // 8 barriers
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 10);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar + i, 1);
        mbar_init(bar + 8, 1 << 19);
        mbar_init(bar + 9, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    if (warp == 0) {
        const uint32_t idesc = instr_desc_bf16(128, n, 0, 0);
        const uint32_t hi = desc_hi(8u * 128u, LAYOUT_SW128);
        const uint32_t a_lo0 = desc_lo(smem_u32(a_buf), 16), b_lo0 = desc_lo(smem_u32(b_buf), 16);
        const int nacc = 512 / n < 8 ? 512 / n : 8;
        __syncwarp();
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            if (lag > 0 && t >= lag) {
                mbar_wait(bar + ((t - lag) & 7), (uint32_t)(((t - lag) >> 3) & 1));
                tc_fence_after();
            }
            if ((variant & 8) ? (threadIdx.x == 0) : elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)((t % nacc) * n);
                const uint32_t first = (variant & 2) ? 1u : 0u;
                mma_issue_dyn(d, a_lo0, hi, b_lo0, hi, idesc, first);
                if (variant & 32) {
#pragma unroll
                    for (int i = 1; i < 18; ++i) mma_issue<true>(d, a_lo0 + 2 * (i & 3) + 4 * (i >> 2), hi, b_lo0 + 2 * (i & 3), hi, idesc);
                } else
#pragma unroll 1
                for (int i = 1; i + 3 <= burst; i += 3) {
                    mma_issue<true>(d, a_lo0 + 2, hi, b_lo0 + 2, hi, idesc);
                    mma_issue<true>(d, a_lo0 + 4, hi, b_lo0 + 4, hi, idesc);
                    mma_issue<true>(d, a_lo0 + 6, hi, b_lo0 + 6, hi, idesc);
                }
                if (!(variant & 1)) mma_commit(bar + (t & 7));
                if (variant & 4) mma_commit(bar + 8);
            }
            if (!(variant & 16)) __syncwarp();
        }
        __syncwarp();
        if (elect_one()) mma_commit(bar + 9);
        __syncwarp();
        mbar_wait(bar + 9, 0);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaFuncSetAttribute(burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int tiles = 512;
    printf("%6s %6s %6s %6s %12s %12s\n", "N", "burst", "lag", "warp", "clk/tile", "clk/mma");
    for (int n : {32, 128})
        for (int burst : {19})
            for (int lag : {0})
                for (int ww : {0, 8, 16, 24, 32, 33, 40, 48, 56, 57}) {      // column "warp" = variant here
                    if ((ww & 1) && (ww & 4)) continue;
                    burst_kernel<<<148, 128, 130 * 1024 + 2048>>>(n, burst, tiles, lag, 1, ww, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[148];
                    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("%6d %6d %6d %6d %12.1f %12.1f\n", n, burst, lag, ww, (double)mx / tiles, (double)mx / tiles / burst);
                    fflush(stdout);
                }
    return 0;
}
