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
        const float r = (normalize & 1) ? rsqrtf(ss / (float)L.in_dim + 1e-8f) : 1.f;
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
                const float* xin = in + (int64_t)b0 * row_stride + L.in_off;
                if ((L.in_dim & 127) == 0 && (row_width & 3) == 0 && (L.in_off & 3) == 0 && ((uintptr_t)wrow & 15) == 0 &&
                    ((uintptr_t)in & 15) == 0) {
                    // 16-byte loads, four k per lane and step: a 512-wide layer is 4 dependent load rounds instead of 16
                    // (the kernel is a chain of L2 round trips: 46 us per layer before, profiles/r02_launches_step.md)
                    for (int k = lane * 4; k < L.in_dim; k += 128) {
                        const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (b0 + i < batch) {
                                const float4 xv = *reinterpret_cast<const float4*>(xin + (int64_t)i * row_stride + k);
                                acc[i] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[i]))));
                            }
                    }
                } else {
                    for (int k = lane; k < L.in_dim; k += 32) {
                        const float wv = wrow[k];
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (b0 + i < batch) acc[i] = fmaf(wv, xin[(int64_t)i * row_stride + k], acc[i]);
                    }
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
                    if (!(normalize & 2)) v = 1.4142135623730951f * (v > 0.f ? v : 0.2f * v);     // fused_lrelu (gm.py:192)
                    out[(int64_t)(b0 + lane) * row_stride + L.out_off + j] = v;
                }
            }
        }
        grid.sync();
    }
}


// Backward of the whole mapping network in ONE cooperative kernel (first order): from the saved activations `acts`
// (b200gan_mapping_fwd) and the gradient of the last row, layer by layer from the top:
//     gz = g * sqrt(2) * (y > 0 ? 1 : 0.2)                      (fused leaky-ReLU from its OUTPUT)
//     gW[n][k] = scale * sum_m gz[m][n] * x[m][k],   gb[n] = bias_mul * sum_m gz[m][n],
//     gx[m][k] = scale * sum_n gz[m][n] * W[n][k]
// and finally the (per-slice) PixelNorm.  `gbuf` is a [2][batch][row_width] scratch: gbuf[0] holds dL/d(last row) on
// entry (the caller copies it there), the two rows ping-pong between layer input / output gradients.
__global__ void __launch_bounds__(256) mapping_bwd_kernel(const float* __restrict__ z, const float* __restrict__ acts,
                                                          float* __restrict__ gbuf, const b200gan_fc_layer* __restrict__ layers,
                                                          const b200gan_fc_layer_grad* __restrict__ grads, float* __restrict__ dz,
                                                          int n_groups, int n_layers, int batch, int z_dim, int row_width,
                                                          int normalize) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t layer_stride = (int64_t)batch * row_width;
    float* gcur = gbuf;
    float* gprev = gbuf + layer_stride;
    for (int l = n_layers - 1; l >= 0; --l) {
        const float* x = acts + (int64_t)l * layer_stride;
        const float* y = acts + (int64_t)(l + 1) * layer_stride;
        // phase A: activation gradient in place; clear the input-gradient row
        for (int64_t e = tid; e < layer_stride; e += nthreads) {
            const float yv = y[e];
            if (!(normalize & 2)) gcur[e] *= 1.4142135623730951f * (yv > 0.f ? 1.f : 0.2f);
            gprev[e] = 0.f;
        }
        grid.sync();
        // phase B: one warp item = (group, 32-column chunk of the input, 32-neuron chunk of the output)
        int total = 0;
        for (int g = 0; g < n_groups; ++g) {
            const b200gan_fc_layer L = layers[l * n_groups + g];
            total += ((L.in_dim + 31) / 32) * ((L.out_dim + 31) / 32);
        }
        for (int item = warp_global; item < total; item += n_warps) {
            int g = 0, j = item;
            for (;;) {
                const b200gan_fc_layer Lg = layers[l * n_groups + g];
                const int cnt = ((Lg.in_dim + 31) / 32) * ((Lg.out_dim + 31) / 32);
                if (j < cnt) break;
                j -= cnt;
                ++g;
            }
            const b200gan_fc_layer L = layers[l * n_groups + g];
            const b200gan_fc_layer_grad G = grads[l * n_groups + g];
            const int kchunks = (L.in_dim + 31) / 32;
            const int k = (j % kchunks) * 32 + lane, n0 = (j / kchunks) * 32;
            const int n1 = min(L.out_dim, n0 + 32);
            const bool kin = k < L.in_dim;
            // weight gradient rows n0..n1 at column k, and the bias gradient (by the first column chunk)
            for (int n = n0; n < n1; ++n) {
                float acc = 0.f, bsum = 0.f;
                for (int m = 0; m < batch; ++m) {
                    const float gz = gcur[(int64_t)m * row_width + L.out_off + n];
                    bsum += gz;
                    if (kin) acc = fmaf(gz, x[(int64_t)m * row_width + L.in_off + k], acc);
                }
                if (kin && G.gw) G.gw[(int64_t)n * L.in_dim + k] = acc * L.scale;
                if (G.gb && j % kchunks == 0 && lane == 0) G.gb[n] = bsum * L.bias_mul;
            }
            // input gradient: partial sums over this chunk of neurons
            if (kin) {
                for (int m0 = 0; m0 < batch; m0 += MB) {
                    float acc[MB];
#pragma unroll
                    for (int i = 0; i < MB; ++i) acc[i] = 0.f;
                    for (int n = n0; n < n1; ++n) {
                        const float wv = L.w[(int64_t)n * L.in_dim + k];
#pragma unroll
                        for (int i = 0; i < MB; ++i)
                            if (m0 + i < batch) acc[i] = fmaf(wv, gcur[(int64_t)(m0 + i) * row_width + L.out_off + n], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < MB; ++i)
                        if (m0 + i < batch) atomicAdd(gprev + (int64_t)(m0 + i) * row_width + L.in_off + k, acc[i] * L.scale);
                }
            }
        }
        grid.sync();
        float* t = gcur;
        gcur = gprev;
        gprev = t;
    }
    // gcur = dL/d(acts[0]); PixelNorm backward (gm.py:52-57): dz = r*g - xhat * (r/n) * sum(g * xhat)
    if (dz != nullptr) {
        for (int item = warp_global; item < batch * n_groups; item += n_warps) {
            const int b = item / n_groups, g = item % n_groups;
            const b200gan_fc_layer L = layers[g];
            const float* src = z + (int64_t)b * z_dim + L.in_off;
            const float* gr = gcur + (int64_t)b * row_width + L.in_off;
            const float* xh = acts + (int64_t)b * row_width + L.in_off;
            float* dst = dz + (int64_t)b * z_dim + L.in_off;
            if (!(normalize & 1)) {
                for (int k = lane; k < L.in_dim; k += 32) dst[k] = gr[k];
                continue;
            }
            float ss = 0.f, dot = 0.f;
            for (int k = lane; k < L.in_dim; k += 32) {
                ss += src[k] * src[k];
                dot += gr[k] * xh[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
                dot += __shfl_xor_sync(0xffffffffu, dot, o);
            }
            const float r = rsqrtf(ss / (float)L.in_dim + 1e-8f);
            for (int k = lane; k < L.in_dim; k += 32) dst[k] = r * gr[k] - xh[k] * (r / (float)L.in_dim) * dot;
        }
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

extern "C" int b200gan_mapping_bwd(const float* z, const float* acts, float* gbuf, const b200gan_fc_layer* layers,
                                   const b200gan_fc_layer_grad* grads, float* dz, int n_groups, int n_layers, int batch,
                                   int z_dim, int row_width, int normalize, void* stream) {
    using namespace b200gan;
    B200_REQUIRE(n_groups >= 1 && n_layers >= 1 && batch >= 0 && z_dim >= 1 && row_width >= 1, "mapping_bwd: bad shape");
    if (batch == 0) return 0;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mapping_bwd_kernel, 256, 0);
    B200_REQUIRE(per_sm >= 1, "mapping_bwd: kernel does not fit on an SM");
    int blocks = sm_count();
    void* args[] = {(void*)&z, (void*)&acts, (void*)&gbuf, (void*)&layers, (void*)&grads, (void*)&dz, (void*)&n_groups,
                    (void*)&n_layers, (void*)&batch, (void*)&z_dim, (void*)&row_width, (void*)&normalize};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)mapping_bwd_kernel, dim3(blocks), dim3(256), args, 0,
                                                (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("mapping_bwd: %s", cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}
