// The mapping network (8 x EqualLinear + fused leaky-ReLU behind a PixelNorm, gan_model.py:633-642;
// or gan-control's block-diagonal split-FC `MultiFcStack`, gm.py:489-502) as ONE persistent
// cooperative kernel.  The reference issues 8..56 tiny GEMM + bias + activation launches (pure launch
// latency); here every CTA stays resident, one warp owns one output neuron at a time
// (weight rows streamed once, coalesced, straight from L2/HBM), and a grid-wide barrier separates
// layers.  All layer outputs are kept (they are the backward pass's saved activations).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200gan {

constexpr int MB = 8;   // batch rows per register block

__global__ void __launch_bounds__(256) mapping_kernel(const float* __restrict__ z, float* __restrict__ acts,
                                                      const b200gan_fc_layer* __restrict__ layers, int n_groups,
                                                      int n_layers, int batch, int z_dim, int row_width,
                                                      int normalize) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int64_t row_stride = row_width;
    const int64_t layer_stride = (int64_t)batch * row_width;

    // stage 0: (per-slice) PixelNorm of z into acts[0]
    for (int item = warp_global; item < batch * n_groups; item += n_warps) {
        const int b = item / n_groups, g = item % n_groups;
        const b200gan_fc_layer L = layers[g];
        const float* src = z + (int64_t)b * z_dim + L.in_off;
        float ss = 0.f;
        for (int k = lane; k < L.in_dim; k += 32) ss += src[k] * src[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float r = normalize ? rsqrtf(ss / (float)L.in_dim + 1e-8f) : 1.f;
        float* dst = acts + (int64_t)b * row_stride + L.in_off;
        for (int k = lane; k < L.in_dim; k += 32) dst[k] = src[k] * r;
    }
    grid.sync();

    for (int l = 0; l < n_layers; ++l) {
        const float* in = acts + (int64_t)l * layer_stride;
        float* out = acts + (int64_t)(l + 1) * layer_stride;
        // enumerate (group, neuron) pairs
        int total = 0;
        for (int g = 0; g < n_groups; ++g) total += layers[l * n_groups + g].out_dim;
        for (int item = warp_global; item < total; item += n_warps) {
            int g = 0, j = item;
            while (j >= layers[l * n_groups + g].out_dim) {
                j -= layers[l * n_groups + g].out_dim;
                ++g;
            }
            const b200gan_fc_layer L = layers[l * n_groups + g];
            const float* wrow = L.w + (int64_t)j * L.in_dim;
            const float bj = L.bias ? L.bias[j] * L.bias_mul : 0.f;
            for (int b0 = 0; b0 < batch; b0 += MB) {
                float acc[MB];
#pragma unroll
                for (int i = 0; i < MB; ++i) acc[i] = 0.f;
                for (int k = lane; k < L.in_dim; k += 32) {
                    const float wv = wrow[k];
#pragma unroll
                    for (int i = 0; i < MB; ++i)
                        if (b0 + i < batch)
                            acc[i] = fmaf(wv, in[(int64_t)(b0 + i) * row_stride + L.in_off + k], acc[i]);
                }
#pragma unroll
                for (int i = 0; i < MB; ++i) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
                }
                if (lane < MB && b0 + lane < batch) {
                    float v = 0.f;
#pragma unroll
                    for (int i = 0; i < MB; ++i)
                        if (i == lane) v = acc[i];
                    v = v * L.scale + bj;
                    v = 1.4142135623730951f * (v > 0.f ? v : 0.2f * v);     // fused_lrelu (gm.py:192)
                    out[(int64_t)(b0 + lane) * row_stride + L.out_off + j] = v;
                }
            }
        }
        grid.sync();
    }
}

}  // namespace b200gan

extern "C" int b200gan_mapping_fwd(const float* z, float* acts, const b200gan_fc_layer* layers, int n_groups,
                                   int n_layers, int batch, int z_dim, int row_width, int normalize, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n_groups >= 1 && n_layers >= 1 && batch >= 0 && z_dim >= 1 && row_width >= 1, "mapping_fwd: bad shape");
    if (batch == 0) return 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mapping_kernel, 256, 0);
    B200_REQUIRE(per_sm >= 1, "mapping_fwd: kernel does not fit on an SM");
    int blocks = sm_count();   // one resident CTA per SM: 8 warps x 148 = 1184 neurons in flight
    void* args[] = {(void*)&z, (void*)&acts, (void*)&layers, (void*)&n_groups, (void*)&n_layers,
                    (void*)&batch, (void*)&z_dim, (void*)&row_width, (void*)&normalize};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)mapping_kernel, dim3(blocks), dim3(256), args, 0,
                                                (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("mapping_fwd: %s", cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}
