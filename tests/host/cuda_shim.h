// TEST INFRASTRUCTURE ONLY.  Host stand-ins for the CUDA built-ins used by the per-point device functions of the fused
// alignment kernel (super_primitive_b200/csrc/spb_fast.cuh, spb_gn_packed.cuh), so that g++ can compile those functions
// from the SAME source for the CPU tests (tests/test_kernel_math_host_cpu.py).  Approximate device intrinsics
// (ex2.approx, rcp.approx, flush-to-zero) become their IEEE counterparts: results agree to float32 rounding, not bit
// for bit.  Build with:  -D__device__= -D__forceinline__=inline -D__global__= -D__restrict__= -DSPB_TAP_L2_256=0
#pragma once
#include <cuda_runtime.h>      // float2 / float4 / make_float4: plain structs on the host
#include <math.h>
#include <stdint.h>
#include <string.h>

struct HostIdx { unsigned x, y, z; };
static HostIdx threadIdx, blockIdx, blockDim, gridDim;

#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
