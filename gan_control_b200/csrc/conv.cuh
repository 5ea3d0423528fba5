// Geometry shared by the convolution engines (conv_simt.cu, conv_umma.cu) and their C API.
#pragma once
#include "common.cuh"

namespace b200gan {

// y[b][oy][ox][o] = sum_{ky,kx,i} z[b][oy*down+ky-pad0][ox*down+kx-pad0][i] * w[wb][ky][kx][o][i]
// z = x zero-upsampled by `up`; see include/b200gan.h.
struct ConvGeom {
    int b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample;
    // "Packed" sides (up = down = 1 only): the convolution runs on a space-to-depth VIEW of a tensor that stays in its
    // plain NHWC layout in memory.  pack_in: the logical input (in_h, in_w, ic) is the physical tensor
    // (2*in_h, 2*in_w, ic/4) with logical channel (py*2+px)*ic/4 + c <-> physical pixel (2*iy+py, 2*ix+px), channel c.
    // pack_out: same for the output (the kernel stores depth-to-space).  A stride-2 convolution preceded by the FIR
    // blur, or a transposed stride-2 convolution followed by it, is a 3x3 stride-1 convolution between such views
    // with composite weights (ops.py), so the blurred / zero-inserted intermediate never exists.
    int pack_in = 0, pack_out = 0;
    int f16 = 0;             // 16-bit storage is IEEE half instead of bfloat16 (tcgen05 engines; set by the dispatcher)
};

// the epilogue of every forward engine (include/b200gan.h b200gan_conv_epilogue)
typedef b200gan_conv_epilogue ConvEp;
static inline bool ep_active(const ConvEp& e) {
    return e.bias || e.rowscale || e.noise || e.addend || e.gate || e.slope != 1.f || e.gain != 1.f;
}

int conv_fwd_simt(const void* x, const void* w, void* y, int dtype, const ConvGeom& g, const ConvEp& ep, cudaStream_t st);
// streaming 1x1 kernels for a <= 4-channel side (conv_pointwise.cu)
bool conv_fwd_pointwise_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y);
int conv_fwd_pointwise(const void* x, const void* w, void* y, int dtype, const ConvGeom& g, const ConvEp& ep, cudaStream_t st);
bool conv_wgrad_pointwise_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy);
int conv_wgrad_pointwise(const void* x, const void* gy, float* gw, int dtype, const ConvGeom& g, cudaStream_t st);
// tcgen05 engine (conv_umma.cu)
bool conv_fwd_umma_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y);
int conv_fwd_umma(const void* x, const void* w, void* y, const ConvGeom& g, const ConvEp& ep, cudaStream_t st);
bool conv_wgrad_umma_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy);
int conv_wgrad_umma(const void* x, const void* gy, float* gw, const ConvGeom& g, cudaStream_t st);
// halo-reuse variant for <= 64-channel stride-1 layers (conv_umma_halo.cu)
bool conv_fwd_halo_eligible(int dtype, const ConvGeom& g, const void* x, const void* w, const void* y);
int conv_fwd_halo(const void* x, const void* w, void* y, const ConvGeom& g, const ConvEp& ep, cudaStream_t st);
bool conv_wgrad_halo_eligible(int dtype, const ConvGeom& g, const void* x, const void* gy);
int conv_wgrad_halo(const void* x, const void* gy, float* gw, const ConvGeom& g, cudaStream_t st);
int conv_wgrad_simt(const void* x, const void* gy, float* gw, int dtype, const ConvGeom& g, cudaStream_t st);

}  // namespace b200gan
