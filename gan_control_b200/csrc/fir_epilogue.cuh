// Optional epilogue of the FIR kernels: the StyledConv tail (demodulation scale gm.py:288-289, NoiseInjection
// gm.py:340-345, FusedLeakyReLU gm.py:32-35) applied to the filter output while it is still in registers,
//     y = gain * lrelu(v * rowscale[n][c] + noise_w * noise[n][pixel] + bias[c]),
// so the blurred up-convolution never makes an extra HBM round trip before its activation.
#pragma once
#include "common.cuh"

namespace b200gan {

struct FirEpilogue {
    const float* bias;       // [c] or null
    const float* rowscale;   // [n][c] or null
    const void* noise;       // [n][pixel] in the activation dtype, or null
    const float* noise_w;    // [1] (device) or null
    float slope, gain;
    int enabled;
};

// v[VEC]: channels c0..c0+VEC-1 of output pixel `pix` (linear index over n*out_h*out_w) of sample `b`
template <typename T, int VEC>
__device__ __forceinline__ void fir_epilogue(const FirEpilogue& ep, float* v, int b, int64_t pix, int c, int c0) {
    float nz = 0.f;
    if (ep.noise != nullptr) nz = (ep.noise_w ? *ep.noise_w : 1.f) * io<T>::ld(reinterpret_cast<const T*>(ep.noise) + pix);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float u = v[j];
        if (ep.rowscale) u *= ep.rowscale[(int64_t)b * c + c0 + j];
        u += nz + (ep.bias ? ep.bias[c0 + j] : 0.f);
        v[j] = ep.gain * (u > 0.f ? u : u * ep.slope);
    }
}

}  // namespace b200gan
