// Shared device/host helpers for libdimsum_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dimsum_b200.h"

namespace dimsum {

// ------------------------------------------------------------------------------------------------
// host side: error string (thread local) + launch counter
// ------------------------------------------------------------------------------------------------
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);   // cudaGetLastError -> DIMSUM_ERR_CUDA, and counts the launch

#define DIMSUM_REQUIRE(cond, code, ...)                   \
    do {                                                  \
        if (!(cond)) return ::dimsum::fail(code, __VA_ARGS__); \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
#define DEV __device__ __forceinline__

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Blackwell packed fp32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2): one issue slot, two lanes.
DEV float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)),
          "l"(reinterpret_cast<uint64_t &>(c)));
    return d;
}
DEV float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
DEV float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)));
    return d;
}
DEV float2 splat2(float v) { return make_float2(v, v); }

// MUFU primitives (no fast-math flag: these are the only approximations, each ~2^-22 relative).
DEV float ex2_mufu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
DEV float rcp_mufu(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus(x) = max(x,0) + log1p(exp(-|x|)),  log1p(w) = 2 atanh(w/(2+w)), w in (0,1].
// Agrees with the reference's `x <= 20 ? log1pf(expf(x)) : x` (selective_scan_fwd_kernel.cuh:153-156)
// to ~1e-7 relative over the whole range (for x > 20 the log term is below half an ulp of x).
DEV float softplus_f(float x) {
    float w = ex2_mufu(-fabsf(x) * kLog2e);
    float s = w * rcp_mufu(2.0f + w);
    float s2 = s * s;
    float p = 1.0f / 13.0f;
    p = fmaf(p, s2, 1.0f / 11.0f);
    p = fmaf(p, s2, 1.0f / 9.0f);
    p = fmaf(p, s2, 1.0f / 7.0f);
    p = fmaf(p, s2, 1.0f / 5.0f);
    p = fmaf(p, s2, 1.0f / 3.0f);
    p = fmaf(p, s2, 1.0f);
    return fmaf(2.0f * s, p, fmaxf(x, 0.0f));
}

// sigmoid / silu
DEV float sigmoid_f(float z) { return rcp_mufu(1.0f + ex2_mufu(-z * kLog2e)); }
DEV float silu_f(float z) { return z * sigmoid_f(z); }

// Cheaper variants for 16-bit outputs (bf16 eps = 2^-8, fp16 eps = 2^-11): one MUFU less each, no polynomial.
//   silu:     z (0.5 + 0.5 tanh(z/2)) with tanh.approx (relative error ~2^-11)
//   softplus: max(x,0) + ln2 * lg2(1 + 2^(-|x| log2 e))  (absolute error ~2e-7)
DEV float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
DEV float lg2_mufu(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool kFast>
DEV float silu_t(float z) {
    if (kFast) {
        const float hz = 0.5f * z;
        return fmaf(hz, tanh_mufu(hz), hz);
    }
    return silu_f(z);
}
template <bool kFast>
DEV float softplus_t(float x) {
    if (kFast) return fmaf(kLn2, lg2_mufu(1.0f + ex2_mufu(-fabsf(x) * kLog2e)), fmaxf(x, 0.0f));
    return softplus_f(x);
}

// ------------------------------------------------------------------------------------------------
// 16-byte vector I/O in the tensor's storage type, fp32 in registers
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Io;

template <>
struct Io<float> {
    static constexpr int kVec = 4;
    static constexpr int kDtype = DIMSUM_F32;
    static DEV float ld(const float *p) { return *p; }
    static DEV void st(float *p, float v) { *p = v; }
    static DEV void ldv(const float *p, float (&v)[4]) {
        float4 r = *reinterpret_cast<const float4 *>(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static DEV void stv(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    // four consecutive elements (16 bytes here, 8 bytes for the 16-bit types)
    static DEV void ld4(const float *p, float (&v)[4]) { ldv(p, v); }
    static DEV void st4(float *p, const float (&v)[4]) { stv(p, v); }
};

template <>
struct Io<__nv_bfloat16> {
    static constexpr int kVec = 8;
    static constexpr int kDtype = DIMSUM_BF16;
    static DEV float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static DEV void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
    static DEV void ldv(const __nv_bfloat16 *p, float (&v)[8]) {
        uint4 r = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static DEV void stv(__nv_bfloat16 *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = reinterpret_cast<uint32_t &>(h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    static DEV void ld4(const __nv_bfloat16 *p, float (&v)[4]) {
        const uint2 r = *reinterpret_cast<const uint2 *>(p);
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
    static DEV void st4(__nv_bfloat16 *p, const float (&v)[4]) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(reinterpret_cast<uint32_t &>(h0), reinterpret_cast<uint32_t &>(h1));
    }
};

template <>
struct Io<__half> {
    static constexpr int kVec = 8;
    static constexpr int kDtype = DIMSUM_F16;
    static DEV float ld(const __half *p) { return __half2float(*p); }
    static DEV void st(__half *p, float v) { *p = __float2half_rn(v); }
    static DEV void ldv(const __half *p, float (&v)[8]) {
        uint4 r = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(reinterpret_cast<const __half2 &>(w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static DEV void stv(__half *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = reinterpret_cast<uint32_t &>(h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    static DEV void ld4(const __half *p, float (&v)[4]) {
        const uint2 r = *reinterpret_cast<const uint2 *>(p);
        const float2 f0 = __half22float2(reinterpret_cast<const __half2 &>(r.x)), f1 = __half22float2(reinterpret_cast<const __half2 &>(r.y));
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
    }
    static DEV void st4(__half *p, const float (&v)[4]) {
        __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(reinterpret_cast<uint32_t &>(h0), reinterpret_cast<uint32_t &>(h1));
    }
};

}  // namespace dimsum
