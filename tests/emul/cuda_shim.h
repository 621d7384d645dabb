// Host stand-ins for the few CUDA built-ins that csrc/geom.cuh and csrc/clip.cuh use, so that the SAME source can be
// compiled by g++ and stepped through on the CPU (tests/emul/clip_emul.cpp).  Test infrastructure only.
// Compile with -ffp-contract=off: every *_rn intrinsic is one correctly rounded IEEE operation, as on the device.
#pragma once
#define GLENET_HOST_EMUL 1
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define CUDART_NAN_F __builtin_nanf("")
#define CUDART_INF_F __builtin_inff()
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }   // approximate on the device: only feeds the vertex ORDER
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
using std::min;
using std::max;
