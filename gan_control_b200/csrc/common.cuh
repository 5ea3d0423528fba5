// Shared helpers for libb200gan (sm_100a).  Internal -- the public surface is include/b200gan.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200gan.h"

namespace b200gan {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);   // cudaGetLastError -> return code (+ message)
int sm_count();

#define B200_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            b200gan::set_error(__VA_ARGS__);    \
            return B200GAN_EINVAL;              \
        }                                       \
    } while (0)

template <typename T> struct io;
template <> struct io<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct io<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct io<__half> {
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};

// VEC consecutive elements moved as one (up to 128-bit) access
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
    T v[VEC];
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// floor division / modulo for possibly negative numerators (device)
__device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// Division by a run-time-invariant divisor as multiply-high + shift (host computes the constants): the
// persistent kernels decode a tile index per tile in every warp role, and the compiler's generic 32-bit
// division is ~20 dependent instructions (I2F / MUFU.RCP / F2I ...) on an issue-bound warp.
struct FastDiv {
    uint32_t mul, shift, div;
    __device__ __forceinline__ uint32_t quot(uint32_t x) const { return div == 1 ? x : (__umulhi(x, mul) >> shift); }
    __device__ __forceinline__ void divmod(uint32_t x, uint32_t& q, uint32_t& r) const {
        q = quot(x);
        r = x - q * div;
    }
};
// exact for 0 <= x < 2^31 (round-up method with a 32-bit magic number)
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.div = d;
    if (d <= 1) { f.mul = 0; f.shift = 0; f.div = 1; return f; }
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                       // ceil(log2 d)
    const uint64_t m = ((1ull << (31 + l)) + d - 1) / d;   // ceil(2^(31+l) / d) < 2^32
    f.mul = (uint32_t)m;
    f.shift = l - 1;                                   // x * m >> (32 + l - 1)
    return f;
}

}  // namespace b200gan

// dtype dispatch: binds `T` inside the lambda body
#define B200_DISPATCH(dtype, ...)                                   \
    [&]() -> int {                                                  \
        if ((dtype) == B200GAN_F32) {                               \
            using T = float;                                        \
            return __VA_ARGS__();                                   \
        } else if ((dtype) == B200GAN_BF16) {                       \
            using T = __nv_bfloat16;                                \
            return __VA_ARGS__();                                   \
        } else if ((dtype) == B200GAN_F16) {                        \
            using T = __half;                                       \
            return __VA_ARGS__();                                   \
        }                                                           \
        b200gan::set_error("unsupported dtype %d", (int)(dtype));   \
        return B200GAN_EINVAL;                                      \
    }()
