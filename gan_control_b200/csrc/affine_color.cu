// ADA augmentation core (trainers/non_leaking.py): the geometric warp and the colour transform in ONE pass.
//
// The reference (non_leaking.py:314-371, 373-394) materialises a (N, H2, W2, 3) fp32 sampling grid (make_grid ->
// affine_grid -> rescale), calls F.grid_sample(bilinear, zeros, align_corners=False) on the 2x-upsampled image and,
// after the decimating FIR and the crop, multiplies every pixel with a 3x3 colour matrix and adds an offset.  The whole
// chain grid -> sample coordinate is one affine map per sample, so the kernel evaluates it per output pixel from six
// coefficients (fp64: the coordinates span ~2000 pixels and the bilinear weights need their fraction) and never reads a
// grid.  The colour transform is linear per pixel and the same at every pixel, so it commutes with the FIR that
// follows: it is applied here to the bilinear sample, with its offset divided by the FIR's DC gain by the caller.
// Any element strides (NCHW planar or channels-last) on both sides; <= 4 channels.
#include "common.cuh"

namespace b200gan {

struct AcParams {
    int n, c, ih, iw, oh, ow;
    int64_t xs[4], ys[4];           // element strides (n, c, h, w)
    const double* mat;              // [n][6]: source pixel (sx, sy) = (m0 ox + m1 oy + m2, m3 ox + m4 oy + m5)
    const float* color;             // [n][c][c + 1] rows (matrix | offset) or null
};

__device__ __forceinline__ void ac_coords(const AcParams& p, int b, int ox, int oy, int& x0, int& y0, float& fx, float& fy) {
    const double* m = p.mat + (int64_t)b * 6;
    const double sx = m[0] * ox + m[1] * oy + m[2];
    const double sy = m[3] * ox + m[4] * oy + m[5];
    const double flx = floor(sx), fly = floor(sy);
    // far outside: clamp so that the int conversion is defined; such taps are out of bounds anyway
    x0 = (int)fmin(fmax(flx, -2.0), (double)p.iw + 1.0);
    y0 = (int)fmin(fmax(fly, -2.0), (double)p.ih + 1.0);
    fx = (float)(sx - flx);
    fy = (float)(sy - fly);
}

template <typename T>
__global__ void __launch_bounds__(256) affine_color_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, AcParams p) {
    const int b = blockIdx.y;
    const int64_t total = (int64_t)p.oh * p.ow;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / p.ow), ox = (int)(i - (int64_t)oy * p.ow);
        int x0, y0;
        float fx, fy;
        ac_coords(p, b, ox, oy, x0, y0, fx, fy);
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        const bool vx0 = x0 >= 0 && x0 < p.iw, vx1 = x0 + 1 >= 0 && x0 + 1 < p.iw;
        const bool vy0 = y0 >= 0 && y0 < p.ih, vy1 = y0 + 1 >= 0 && y0 + 1 < p.ih;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const T* xb = x + (int64_t)b * p.xs[0] + (int64_t)y0 * p.xs[2] + (int64_t)x0 * p.xs[3];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (ch >= p.c) break;
            const T* q = xb + ch * p.xs[1];
            float a = 0.f;
            if (vy0 && vx0) a = fmaf(w00, io<T>::ld(q), a);
            if (vy0 && vx1) a = fmaf(w01, io<T>::ld(q + p.xs[3]), a);
            if (vy1 && vx0) a = fmaf(w10, io<T>::ld(q + p.xs[2]), a);
            if (vy1 && vx1) a = fmaf(w11, io<T>::ld(q + p.xs[2] + p.xs[3]), a);
            v[ch] = a;
        }
        T* yb = y + (int64_t)b * p.ys[0] + (int64_t)oy * p.ys[2] + (int64_t)ox * p.ys[3];
        if (p.color) {
            const float* cm = p.color + (int64_t)b * p.c * (p.c + 1);
            for (int o = 0; o < p.c; ++o) {
                float a = cm[o * (p.c + 1) + p.c];
                for (int ch = 0; ch < p.c; ++ch) a = fmaf(cm[o * (p.c + 1) + ch], v[ch], a);
                io<T>::st(yb + o * p.ys[1], a);
            }
        } else {
            for (int ch = 0; ch < p.c; ++ch) io<T>::st(yb + ch * p.ys[1], v[ch]);
        }
    }
}

// adjoint w.r.t. the image: gx (fp32, planar (n, c, ih, iw), zeroed by the caller) += scatter of C^T gy
template <typename T>
__global__ void __launch_bounds__(256) affine_color_bwd_kernel(const T* __restrict__ gy, float* __restrict__ gx, AcParams p) {
    const int b = blockIdx.y;
    const int64_t total = (int64_t)p.oh * p.ow;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / p.ow), ox = (int)(i - (int64_t)oy * p.ow);
        int x0, y0;
        float fx, fy;
        ac_coords(p, b, ox, oy, x0, y0, fx, fy);
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        const bool vx0 = x0 >= 0 && x0 < p.iw, vx1 = x0 + 1 >= 0 && x0 + 1 < p.iw;
        const bool vy0 = y0 >= 0 && y0 < p.ih, vy1 = y0 + 1 >= 0 && y0 + 1 < p.ih;
        const T* gb = gy + (int64_t)b * p.ys[0] + (int64_t)oy * p.ys[2] + (int64_t)ox * p.ys[3];
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < p.c; ++o) g[o] = io<T>::ld(gb + o * p.ys[1]);
        float v[4] = {g[0], g[1], g[2], g[3]};
        if (p.color) {
            const float* cm = p.color + (int64_t)b * p.c * (p.c + 1);
            for (int ch = 0; ch < p.c; ++ch) {
                float a = 0.f;
                for (int o = 0; o < p.c; ++o) a = fmaf(cm[o * (p.c + 1) + ch], g[o], a);
                v[ch] = a;
            }
        }
        float* xb = gx + ((int64_t)b * p.c * p.ih + y0) * p.iw + x0;
        for (int ch = 0; ch < p.c; ++ch) {
            float* q = xb + (int64_t)ch * p.ih * p.iw;
            if (vy0 && vx0) atomicAdd(q, w00 * v[ch]);
            if (vy0 && vx1) atomicAdd(q + 1, w01 * v[ch]);
            if (vy1 && vx0) atomicAdd(q + p.iw, w10 * v[ch]);
            if (vy1 && vx1) atomicAdd(q + p.iw + 1, w11 * v[ch]);
        }
    }
}

static int ac_fill(AcParams& p, int n, int c, int ih, int iw, int oh, int ow, const int64_t* xs, const int64_t* ys,
                   const double* mat, const float* color) {
    B200_REQUIRE(n >= 0 && c >= 1 && c <= 4 && ih >= 1 && iw >= 1 && oh >= 1 && ow >= 1, "affine_color: bad shape n=%d c=%d in=%dx%d out=%dx%d",
                 n, c, ih, iw, oh, ow);
    B200_REQUIRE(mat != nullptr && xs != nullptr && ys != nullptr, "affine_color: null matrix / strides");
    p.n = n; p.c = c; p.ih = ih; p.iw = iw; p.oh = oh; p.ow = ow; p.mat = mat; p.color = color;
    for (int i = 0; i < 4; ++i) { p.xs[i] = xs[i]; p.ys[i] = ys[i]; }
    return 0;
}

static dim3 ac_grid(const AcParams& p) {
    int64_t bx = cdiv((int64_t)p.oh * p.ow, 256 * 4);
    const int64_t cap = cdiv((int64_t)sm_count() * 16, p.n > 0 ? p.n : 1);
    if (bx > cap) bx = cap;
    return dim3((unsigned)(bx < 1 ? 1 : bx), (unsigned)p.n);
}

}  // namespace b200gan

extern "C" int b200gan_affine_color_fwd(const void* x, void* y, const double* mat, const float* color, int dtype, int n, int c,
                                        int in_h, int in_w, int out_h, int out_w, const int64_t* x_strides,
                                        const int64_t* y_strides, void* stream) {
    using namespace b200gan;
    AcParams p;
    if (int rc = ac_fill(p, n, c, in_h, in_w, out_h, out_w, x_strides, y_strides, mat, color)) return rc;
    if (n == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        affine_color_fwd_kernel<T><<<ac_grid(p), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, p);
        count_launch();
        return check_launch("affine_color_fwd");
    });
}

extern "C" int b200gan_affine_color_bwd(const void* gy, float* gx, const double* mat, const float* color, int dtype, int n, int c,
                                        int in_h, int in_w, int out_h, int out_w, const int64_t* gy_strides, void* stream) {
    using namespace b200gan;
    AcParams p;
    const int64_t planar[4] = {(int64_t)c * in_h * in_w, (int64_t)in_h * in_w, in_w, 1};
    if (int rc = ac_fill(p, n, c, in_h, in_w, out_h, out_w, planar, gy_strides, mat, color)) return rc;
    if (n == 0) return 0;
    return B200_DISPATCH(dtype, [&] {
        affine_color_bwd_kernel<T><<<ac_grid(p), 256, 0, (cudaStream_t)stream>>>((const T*)gy, gx, p);
        count_launch();
        return check_launch("affine_color_bwd");
    });
}
