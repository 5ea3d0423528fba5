// extern "C" entry points of the convolution family; picks the engine per call.
#include "conv.cuh"


extern "C" int b200gan_conv_fwd(const void* x, const void* w, void* y, int dtype, int b, int in_h, int in_w, int ic,
                                int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                int w_per_sample, const float* bias, const float* rowscale, const void* noise,
                                const float* noise_w, float slope, float gain, void* stream) {
    using namespace b200gan;
    ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    return conv_fwd_simt(x, w, y, dtype, g, bias, rowscale, noise, noise_w, slope, gain, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_wgrad(const void* x, const void* gy, float* gw, int dtype, int b, int in_h, int in_w,
                                  int ic, int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                  int w_per_sample, void* stream) {
    using namespace b200gan;
    ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    return conv_wgrad_simt(x, gy, gw, dtype, g, (cudaStream_t)stream);
}
