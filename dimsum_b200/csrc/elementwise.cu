// Token-major glue around the mixer for sm_100a: adaLN modulate and gated residual with the scan order folded into the
// row index (so the transpose / flip / table orders of DiMSUM cost no extra pass and no permuted copy), and the fused
// residual-add + RMSNorm that the reference runs as a Triton kernel (mamba_ssm/ops/triton/layernorm.py:62-118).
// All three are pure streaming kernels: 16-byte coalesced row accesses, one pass.
#include "common.cuh"

namespace dimsum {
namespace {

// 8 consecutive channels of a tensor whose dtype is only known at run time (uniform branch): 32 B (fp32) or 16 B.
DEV void ld8(const void *base, int dtype, int64_t idx, float (&v)[8]) {
    if (dtype == DIMSUM_F32) {
        const float *p = reinterpret_cast<const float *>(base) + idx;
        Io<float>::ldv(p, reinterpret_cast<float(&)[4]>(v[0]));
        Io<float>::ldv(p + 4, reinterpret_cast<float(&)[4]>(v[4]));
    } else if (dtype == DIMSUM_BF16) {
        Io<__nv_bfloat16>::ldv(reinterpret_cast<const __nv_bfloat16 *>(base) + idx, v);
    } else {
        Io<__half>::ldv(reinterpret_cast<const __half *>(base) + idx, v);
    }
}
DEV void st8(void *base, int dtype, int64_t idx, const float (&v)[8]) {
    if (dtype == DIMSUM_F32) {
        float *p = reinterpret_cast<float *>(base) + idx;
        Io<float>::stv(p, reinterpret_cast<const float(&)[4]>(v[0]));
        Io<float>::stv(p + 4, reinterpret_cast<const float(&)[4]>(v[4]));
    } else if (dtype == DIMSUM_BF16) {
        Io<__nv_bfloat16>::stv(reinterpret_cast<__nv_bfloat16 *>(base) + idx, v);
    } else {
        Io<__half>::stv(reinterpret_cast<__half *>(base) + idx, v);
    }
}

// 8 consecutive channels as raw bits (32 B fp32 / 16 B 16-bit) and their conversion: see ld_raw4 below for why they are apart
struct Raw8 { uint4 a, b; };
DEV Raw8 ld_raw8(const void *base, int dtype, int64_t idx) {
    Raw8 r;
    r.b = make_uint4(0u, 0u, 0u, 0u);
    if (dtype == DIMSUM_F32) {
        const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(base) + idx);
        r.a = p[0]; r.b = p[1];
    } else {
        r.a = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(base) + idx);
    }
    return r;
}
DEV void cvt_raw8(const Raw8 &r, int dtype, float (&v)[8]) {
    if (dtype == DIMSUM_F32) {
        v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
        v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
    } else {
        const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (dtype == DIMSUM_BF16) {
                v[2 * i] = __uint_as_float(w[i] << 16);
                v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 f = __half22float2(reinterpret_cast<const __half2 &>(w[i]));
                v[2 * i] = f.x; v[2 * i + 1] = f.y;
            }
        }
    }
}

// one thread = 8 consecutive channels of one token; all loads are issued before the arithmetic
template <bool kGate>
__global__ void __launch_bounds__(256) rowwise_kernel(const dimsum_rowwise_params p) {
    // 32-bit indexing inside a batch row (blockIdx.y = batch): a flat 64-bit index costs three 64-bit divisions per thread
    const unsigned tpt = (unsigned)(p.channels / 8);          // threads per token row
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)p.seqlen * tpt) return;
    const int l = (int)(idx / tpt);
    const int v = (int)(idx - (unsigned)l * tpt);
    const int64_t b = blockIdx.y;
    const int src_l = p.idx != nullptr ? p.idx[l] : l;
    const int c0 = v * 8;
    float a[8], q[8], r[8], o[8];
    // the three loads go out back to back as raw bits; the conversions (their first use) come after the last one is issued
    const int x_dt = (int)p.x_dtype, aux_dt = (int)p.aux_dtype;
    const Raw8 ra = ld_raw8(p.x, x_dt, b * p.x_batch_stride + (int64_t)(kGate ? l : src_l) * p.x_token_stride + c0);
    const Raw8 rq = kGate ? ld_raw8(p.m, aux_dt, b * p.m_batch_stride + (int64_t)src_l * p.m_token_stride + c0)
                          : ld_raw8(p.shift, aux_dt, b * p.vec_row_stride + c0);
    const Raw8 rr = ld_raw8(kGate ? p.gate : p.scale, aux_dt, b * p.vec_row_stride + c0);
    cvt_raw8(ra, x_dt, a);
    cvt_raw8(rq, aux_dt, q);
    cvt_raw8(rr, aux_dt, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = kGate ? fmaf(r[i], q[i], a[i]) : fmaf(a[i], 1.f + r[i], q[i]);
    st8(p.dst, (int)p.dst_dtype, b * p.dst_batch_stride + (int64_t)l * p.dst_token_stride + c0, o);
}

// VEC consecutive channels of a tensor whose dtype is only known at run time (uniform branch)
template <int VEC>
DEV void ld_rt(const void *base, int dtype, int64_t idx, float (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
        float (&q)[4] = reinterpret_cast<float(&)[4]>(v[i]);
        if (dtype == DIMSUM_F32) Io<float>::ld4(reinterpret_cast<const float *>(base) + idx + i, q);
        else if (dtype == DIMSUM_BF16) Io<__nv_bfloat16>::ld4(reinterpret_cast<const __nv_bfloat16 *>(base) + idx + i, q);
        else Io<__half>::ld4(reinterpret_cast<const __half *>(base) + idx + i, q);
    }
}
template <int VEC>
DEV void st_rt(void *base, int dtype, int64_t idx, const float (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
        const float (&q)[4] = reinterpret_cast<const float(&)[4]>(v[i]);
        if (dtype == DIMSUM_F32) Io<float>::st4(reinterpret_cast<float *>(base) + idx + i, q);
        else if (dtype == DIMSUM_BF16) Io<__nv_bfloat16>::st4(reinterpret_cast<__nv_bfloat16 *>(base) + idx + i, q);
        else Io<__half>::st4(reinterpret_cast<__half *>(base) + idx + i, q);
    }
}

// Four consecutive channels as RAW bits (16 B for fp32, 8 B for the 16-bit types) and their conversion, kept apart on purpose:
// a converting load is used the moment it is issued, and an in-order warp then waits for it before it issues the next one.
// Kernels that depend on many loads in flight issue all their ld_raw4 first and convert afterwards.
DEV uint4 ld_raw4(const void *base, int dtype, int64_t idx) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (dtype == DIMSUM_F32) {
        r = *reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(base) + idx);
    } else {
        const uint2 h = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(base) + idx);
        r.x = h.x; r.y = h.y;
    }
    return r;
}
DEV void cvt_raw4(const uint4 r, int dtype, float (&v)[4]) {
    if (dtype == DIMSUM_F32) {
        v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    } else if (dtype == DIMSUM_BF16) {
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    } else {
        const float2 f0 = __half22float2(reinterpret_cast<const __half2 &>(r.x)), f1 = __half22float2(reinterpret_cast<const __half2 &>(r.y));
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
    }
}

// residual add + RMSNorm / LayerNorm (+ adaLN modulate): one warp per row, the row lives in registers between the
// passes (channels <= 32 lanes * 8 * VEC: 1024 fp32, 2048 16-bit)
// kMaxIter = 16-byte vectors per lane (1, 2, 4 or 8): sized to the row so that no predicated-off iterations are issued
// one 16-byte vector of T as raw bits -> Io<T>::kVec floats
template <typename T>
DEV void cvt_vec(const uint4 r, float (&v)[Io<T>::kVec]) {
    if constexpr (Io<T>::kVec == 4) {
        cvt_raw4(r, DIMSUM_F32, v);
    } else {
        Raw8 q;
        q.a = r; q.b = r;
        cvt_raw8(q, Io<T>::kDtype, v);
    }
}

template <typename T, bool kLayerNorm, int kMaxIter>
__global__ void __launch_bounds__(256) norm_kernel(const dimsum_norm_modulate_params p) {
    constexpr int VEC = Io<T>::kVec, Q = VEC / 4;
    const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= p.rows) return;                         // warp-uniform: the fences below see all 32 lanes
    const int nvec = (int)(p.channels / VEC);
    const T *x = reinterpret_cast<const T *>(p.x) + (int64_t)warp * p.x_row_stride;
    const float *res = p.residual != nullptr ? reinterpret_cast<const float *>(p.residual) + (int64_t)warp * p.channels : nullptr;
    float *res_out = p.res_out != nullptr ? reinterpret_cast<float *>(p.res_out) + (int64_t)warp * p.channels : nullptr;
    float vals[kMaxIter][VEC];
    float ss = 0.f, sum = 0.f;
    {
        // every load of the row goes out before the first value is used (raw bits now, conversion / add after the fence):
        // vectors past the end of a short row re-read the last one and are zeroed
        uint4 xr[kMaxIter], rr[kMaxIter][Q];
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = min(lane + it * 32, nvec - 1);
            xr[it] = *reinterpret_cast<const uint4 *>(x + v * VEC);
            if (res != nullptr) {
#pragma unroll
                for (int q = 0; q < Q; ++q) rr[it][q] = *reinterpret_cast<const uint4 *>(res + v * VEC + 4 * q);
            }
        }
        __syncwarp(__activemask());                     // scheduling fence (see colsum_kernel)
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = lane + it * 32;
            cvt_vec<T>(xr[it], vals[it]);
            if (res != nullptr) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    float r[4];
                    cvt_raw4(rr[it][q], DIMSUM_F32, r);
#pragma unroll
                    for (int i = 0; i < 4; ++i) vals[it][4 * q + i] += r[i];
                }
            }
            if (v < nvec) {
                if (res_out != nullptr) {
#pragma unroll
                    for (int i = 0; i < VEC; i += 4)
                        *reinterpret_cast<float4 *>(res_out + v * VEC + i) =
                            make_float4(vals[it][i], vals[it][i + 1], vals[it][i + 2], vals[it][i + 3]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) vals[it][i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                if (kLayerNorm) sum += vals[it][i];
                else ss = fmaf(vals[it][i], vals[it][i], ss);
            }
        }
    }
    // the adaLN rows (L2 hits, shared by the tokens of a batch row) are requested now, all of them, and used after the reductions
    const bool mod = p.scale != nullptr;
    const int aux_dt = (int)p.aux_dtype;
    const int64_t mrow = mod ? (warp / p.rows_per_batch) * p.vec_row_stride : 0;
    uint4 shr[kMaxIter][Q], scr[kMaxIter][Q];
    if (mod) {
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = min(lane + it * 32, nvec - 1);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                shr[it][q] = ld_raw4(p.shift, aux_dt, mrow + v * VEC + 4 * q);
                scr[it][q] = ld_raw4(p.scale, aux_dt, mrow + v * VEC + 4 * q);
            }
        }
        __syncwarp(__activemask());
    }
    float mean = 0.f;
    if (kLayerNorm) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mean = sum / (float)p.channels;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            if (lane + it * 32 < nvec) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) { vals[it][i] -= mean; ss = fmaf(vals[it][i], vals[it][i], ss); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)p.channels + p.eps);
    const float *w = reinterpret_cast<const float *>(p.weight);
    const int64_t yrow = (int64_t)warp * p.y_row_stride;
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
        const int v = lane + it * 32;
        if (v < nvec) {
            float o[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = vals[it][i] * rstd;
            if (!kLayerNorm) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) o[i] *= w[v * VEC + i];
            }
            if (mod) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    float sh[4], sc[4];
                    cvt_raw4(shr[it][q], aux_dt, sh);
                    cvt_raw4(scr[it][q], aux_dt, sc);
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[4 * q + i] = fmaf(o[4 * q + i], 1.f + sc[i], sh[i]);
                }
            }
            st_rt<VEC>(p.y, (int)p.y_dtype, yrow + v * VEC, o);
        }
    }
}

int run_norm(const dimsum_norm_modulate_params *p, cudaStream_t stream, const char *who) {
    DIMSUM_REQUIRE(p != nullptr && p->x && p->y, DIMSUM_ERR_INVALID, "%s: null pointer", who);
    DIMSUM_REQUIRE(p->rows >= 0 && p->channels > 0, DIMSUM_ERR_INVALID, "%s: bad sizes", who);
    DIMSUM_REQUIRE(p->x_dtype >= 0 && p->x_dtype <= 2 && p->y_dtype >= 0 && p->y_dtype <= 2, DIMSUM_ERR_INVALID, "%s: unknown dtype", who);
    DIMSUM_REQUIRE(p->norm_kind == 0 || p->norm_kind == 1, DIMSUM_ERR_INVALID, "%s: unknown norm_kind", who);
    DIMSUM_REQUIRE(p->norm_kind == 1 || p->weight != nullptr, DIMSUM_ERR_INVALID, "%s: RMSNorm needs a weight", who);
    DIMSUM_REQUIRE((p->shift == nullptr) == (p->scale == nullptr), DIMSUM_ERR_INVALID, "%s: shift and scale go together", who);
    DIMSUM_REQUIRE(p->scale == nullptr || (p->rows_per_batch > 0 && p->aux_dtype >= 0 && p->aux_dtype <= 2), DIMSUM_ERR_INVALID,
                   "%s: modulate needs rows_per_batch and aux_dtype", who);
    const int vec = p->x_dtype == DIMSUM_F32 ? 4 : 8;
    DIMSUM_REQUIRE(p->channels % vec == 0 && p->channels <= 32 * 8 * vec, DIMSUM_ERR_UNSUPPORTED,
                   "%s: channels=%lld must be a multiple of %d and at most %d", who, (long long)p->channels, vec, 32 * 8 * vec);
    DIMSUM_REQUIRE(aligned16(p->x) && aligned16(p->y) && p->x_row_stride % vec == 0 && p->y_row_stride % 4 == 0 &&
                       (p->residual == nullptr || aligned16(p->residual)) && (p->res_out == nullptr || aligned16(p->res_out)) &&
                       (p->scale == nullptr || (aligned16(p->shift) && aligned16(p->scale) && p->vec_row_stride % 4 == 0)),
                   DIMSUM_ERR_UNSUPPORTED, "%s: rows must be 16-byte aligned", who);
    if (p->rows == 0) return DIMSUM_OK;
    const unsigned blocks = (unsigned)((p->rows * 32 + 255) / 256);
    const int per_lane = (int)((p->channels / vec + 31) / 32);
#define NKI(T, I)                                                                      \
    if (p->norm_kind == 1) norm_kernel<T, true, I><<<blocks, 256, 0, stream>>>(*p);    \
    else norm_kernel<T, false, I><<<blocks, 256, 0, stream>>>(*p);
#define NK(T)                                     \
    if (per_lane <= 1) { NKI(T, 1) }              \
    else if (per_lane <= 2) { NKI(T, 2) }         \
    else if (per_lane <= 4) { NKI(T, 4) }         \
    else { NKI(T, 8) }
    if (p->x_dtype == DIMSUM_F32) { NK(float) }
    else if (p->x_dtype == DIMSUM_BF16) { NK(__nv_bfloat16) }
    else { NK(__half) }
#undef NK
#undef NKI
    return check_launch(who);
}

// tanh-approximated GELU with an accurate tanh: tanh(y) = 1 - 2 / (exp(2y) + 1)
DEV float gelu_tanh_f(float x) {
    const float y = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    const float t = 1.f - 2.f * rcp_mufu(1.f + ex2_mufu(2.f * kLog2e * y));
    return 0.5f * x * (1.f + t);
}

template <typename T, int kNV>
__global__ void __launch_bounds__(256) gelu_mul_kernel(const dimsum_gelu_mul_params p) {
    constexpr int VEC = Io<T>::kVec;
    const int tpr = (int)(p.hidden / (VEC * kNV));
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.rows * tpr) return;
    const int v = (int)(gid % tpr);
    const int64_t r = gid / tpr;
    const T *x = reinterpret_cast<const T *>(p.x) + r * p.x_row_stride + v * VEC * kNV;
    float a[kNV][VEC], b[kNV][VEC], o[VEC];
#pragma unroll
    for (int j = 0; j < kNV; ++j) { Io<T>::ldv(x + j * VEC, a[j]); Io<T>::ldv(x + p.hidden + j * VEC, b[j]); }
    T *y = reinterpret_cast<T *>(p.y) + r * p.y_row_stride + v * VEC * kNV;
#pragma unroll
    for (int j = 0; j < kNV; ++j) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = gelu_tanh_f(a[j][i]) * b[j][i];
        Io<T>::stv(y + j * VEC, o);
    }
}

// CFG combine + Euler update of the sampler in one pass: v = uncond + s (cond - uncond) on the first `channels` channels of the
// model output, x_new = x + dt v for BOTH halves of the CFG batch (they stay identical copies, sample_ddp.py:168-173).
// Arithmetic order and roundings are those of the PyTorch expressions it replaces (no FMA contraction).
template <typename T>
__global__ void __launch_bounds__(256) cfg_euler_kernel(const dimsum_cfg_euler_params p) {
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t per_row = p.channels * p.hw;                  // elements of one latent (guided channels only)
    if (i4 >= p.half_batch * per_row) return;
    const int64_t r = i4 / per_row, e = i4 - r * per_row;
    const T *oc = reinterpret_cast<const T *>(p.model_out) + r * p.out_row_stride + e;
    const T *ou = oc + p.half_batch * p.out_row_stride;
    float c[4], u[4], x[4], o[4];
    Io<T>::ld4(oc, c);
    Io<T>::ld4(ou, u);
    const float *xp = reinterpret_cast<const float *>(p.x) + r * per_row + e;
    *reinterpret_cast<float4 *>(x) = *reinterpret_cast<const float4 *>(xp);
    const float dt = *p.dt, s = p.cfg_scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float v = __fadd_rn(u[i], __fmul_rn(s, __fsub_rn(c[i], u[i])));
        o[i] = __fadd_rn(x[i], __fmul_rn(dt, v));
    }
    float *d0 = reinterpret_cast<float *>(p.x_new) + r * per_row + e;
    *reinterpret_cast<float4 *>(d0) = *reinterpret_cast<const float4 *>(o);
    *reinterpret_cast<float4 *>(d0 + p.half_batch * per_row) = *reinterpret_cast<const float4 *>(o);
    if (p.v_out != nullptr) {                                   // the guided drift itself, for callers that integrate differently
        float vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) vv[i] = __fadd_rn(u[i], __fmul_rn(s, __fsub_rn(c[i], u[i])));
        float *v0 = reinterpret_cast<float *>(p.v_out) + r * per_row + e;
        *reinterpret_cast<float4 *>(v0) = *reinterpret_cast<const float4 *>(vv);
        *reinterpret_cast<float4 *>(v0 + p.half_batch * per_row) = *reinterpret_cast<const float4 *>(vv);
    }
}

// column sums over the tokens of a batch row: CTA = (128 channels, batch row); kColsumWarps warps stride over the tokens
// with kColsumUnroll rows of loads in flight per lane (a (32, 256, 512) bf16 gradient is only 128 CTAs: the bytes in flight
// come from the unroll, not from the grid), a lane owns 4 consecutive channels, partial sums meet in shared memory
constexpr int kColsumWarps = 16;
template <bool kWantX>
__global__ void __launch_bounds__(kColsumWarps * 32) colsum_kernel(const dimsum_colsum_params p) {
    constexpr int kColsumUnroll = kWantX ? 8 : 16;          // 16 loads in flight per lane either way
    __shared__ float red[2][kColsumWarps][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128 + lane * 4;
    const int64_t b = blockIdx.y;
    constexpr bool want_x = kWantX;
    const int g_dt = (int)p.g_dtype, x_dt = (int)p.x_dtype;
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < p.channels) {
        const int64_t gbase = b * p.g_batch_stride + c0, xbase = b * p.x_batch_stride + c0;
        const int last = (int)p.seqlen - 1;
        for (int l0 = warp; l0 < p.seqlen; l0 += kColsumWarps * kColsumUnroll) {
            // rows past the end re-read the last row (unconditional loads: nothing for the later ones to queue behind) and
            // are masked out of the sums
            int lx[kColsumUnroll];
#pragma unroll
            for (int j = 0; j < kColsumUnroll; ++j) {
                const int l = min(l0 + j * kColsumWarps, last);
                lx[j] = (want_x && p.x_idx != nullptr) ? __ldg(p.x_idx + l) : l;
            }
            uint4 graw[kColsumUnroll], xraw[kColsumUnroll];
#pragma unroll
            for (int j = 0; j < kColsumUnroll; ++j)
                graw[j] = ld_raw4(p.g, g_dt, gbase + (int64_t)min(l0 + j * kColsumWarps, last) * p.g_token_stride);
            if (want_x) {
#pragma unroll
                for (int j = 0; j < kColsumUnroll; ++j) xraw[j] = ld_raw4(p.x, x_dt, xbase + (int64_t)lx[j] * p.x_token_stride);
            }
            // scheduling fence: keeps ptxas from sinking the loads between the sums to save registers.  Only the lanes that
            // own channels are here (channels % 128 != 0 leaves the others outside the branch), hence the active mask
            __syncwarp(__activemask());
#pragma unroll
            for (int j = 0; j < kColsumUnroll; ++j) {
                const float keep = (l0 + j * kColsumWarps <= last) ? 1.f : 0.f;
                float g[4], x[4];
                cvt_raw4(graw[j], g_dt, g);
#pragma unroll
                for (int i = 0; i < 4; ++i) { g[i] *= keep; sg[i] += g[i]; }
                if (want_x) {
                    cvt_raw4(xraw[j], x_dt, x);
#pragma unroll
                    for (int i = 0; i < 4; ++i) sx[i] = fmaf(g[i], x[i], sx[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[0][warp][lane * 4 + i] = sg[i]; red[1][warp][lane * 4 + i] = sx[i]; }
    __syncthreads();
    if (threadIdx.x < 256) {                                             // 2 outputs x 128 channels
        const int which = threadIdx.x >> 7, c = threadIdx.x & 127;
        void *dst = which ? p.sum_gx : p.sum_g;
        if (dst != nullptr && blockIdx.x * 128 + c < p.channels) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kColsumWarps; ++w) t += red[which][w][c];
            const int64_t o = b * p.out_row_stride + blockIdx.x * 128 + c;
            if (p.out_dtype == DIMSUM_F32) reinterpret_cast<float *>(dst)[o] = t;
            else if (p.out_dtype == DIMSUM_BF16) reinterpret_cast<__nv_bfloat16 *>(dst)[o] = __float2bfloat16_rn(t);
            else reinterpret_cast<__half *>(dst)[o] = __float2half_rn(t);
        }
    }
}

// d/dx of the tanh-approximated GELU, sharing the tanh with the value
DEV void gelu_tanh_with_grad(float x, float &val, float &grad) {
    const float k = 0.7978845608028654f, c = 0.044715f;
    const float x2 = x * x;
    const float y = k * fmaf(c * x2, x, x);
    const float t = 1.f - 2.f * rcp_mufu(1.f + ex2_mufu(2.f * kLog2e * y));
    val = 0.5f * x * (1.f + t);
    grad = 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * fmaf(3.f * c, x2, 1.f);
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_mul_bwd_kernel(const dimsum_gelu_mul_bwd_params p) {
    constexpr int VEC = Io<T>::kVec;
    const unsigned tpr = (unsigned)(p.hidden / VEC);
    const int64_t r = blockIdx.y + (int64_t)blockIdx.z * gridDim.y;
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= tpr || r >= p.rows) return;
    const T *x = reinterpret_cast<const T *>(p.x) + r * p.x_row_stride + v * VEC;
    float a[VEC], b[VEC], g[VEC], da[VEC], db[VEC];
    Io<T>::ldv(x, a);
    Io<T>::ldv(x + p.hidden, b);
    Io<T>::ldv(reinterpret_cast<const T *>(p.dy) + r * p.dy_row_stride + v * VEC, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float val, grad;
        gelu_tanh_with_grad(a[i], val, grad);
        da[i] = g[i] * b[i] * grad;
        db[i] = g[i] * val;
    }
    T *dx = reinterpret_cast<T *>(p.dx) + r * p.dx_row_stride + v * VEC;
    Io<T>::stv(dx, da);
    Io<T>::stv(dx + p.hidden, db);
}

// RMSNorm backward: one warp per row (the fp32 row h lives in registers), a CTA's 8 warps walk rows blockIdx.x*8+warp,
// += gridDim.x*8, ... and keep their dweight contributions in registers; one shared-memory reduction per CTA at the end
// (the weight is re-read through L1 for every row rather than kept in 32 more registers: at 1024 channels that is the
// difference between one and two resident CTAs per SM, and the kernel is bound by rows in flight)
template <int kMaxIter>
__global__ void __launch_bounds__(256, 2) rmsnorm_bwd_kernel(const dimsum_rmsnorm_bwd_params p) {
    __shared__ float red[8][32 * kMaxIter * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = (int)(p.channels / 4);
    const float *w = reinterpret_cast<const float *>(p.weight);
    float dw[kMaxIter][4];
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) dw[it][i] = 0.f;
    const float inv_c = 1.f / (float)p.channels;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < p.rows; row += (int64_t)gridDim.x * 8) {
        const float *h = reinterpret_cast<const float *>(p.h) + row * p.channels;
        float hv[kMaxIter][4], gv[kMaxIter][4];
        float ss = 0.f;
        {
            // every load of the row is issued before the first value is used (vectors past the end of a short row re-read
            // the last one and are zeroed afterwards)
            uint4 hraw[kMaxIter], graw[kMaxIter];
#pragma unroll
            for (int it = 0; it < kMaxIter; ++it) {
                const int v = min(lane + it * 32, nvec - 1);
                hraw[it] = *reinterpret_cast<const uint4 *>(h + v * 4);
                graw[it] = ld_raw4(p.dy, (int)p.dy_dtype, row * p.dy_row_stride + v * 4);
            }
            __syncwarp(__activemask());           // scheduling fence, as in colsum_kernel (the row loop is warp-uniform)
#pragma unroll
            for (int it = 0; it < kMaxIter; ++it) {
                const float keep = (lane + it * 32 < nvec) ? 1.f : 0.f;
                cvt_raw4(hraw[it], DIMSUM_F32, hv[it]);
                cvt_raw4(graw[it], (int)p.dy_dtype, gv[it]);
#pragma unroll
                for (int i = 0; i < 4; ++i) { hv[it][i] *= keep; gv[it][i] *= keep; ss = fmaf(hv[it][i], hv[it][i], ss); }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss * inv_c + p.eps);
        float c1 = 0.f;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = lane + it * 32;
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nvec) wv = __ldg(reinterpret_cast<const float4 *>(w) + v);
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                hv[it][i] *= rstd;                                   // xhat
                dw[it][i] = fmaf(gv[it][i], hv[it][i], dw[it][i]);
                gv[it][i] *= wr[i];                                  // weight * dy
                c1 = fmaf(gv[it][i], hv[it][i], c1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c1 *= inv_c;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = lane + it * 32;
            if (v < nvec) {
                float d[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = (gv[it][i] - hv[it][i] * c1) * rstd;
                if (p.dres_in != nullptr) {
                    float r[4];
                    Io<float>::ld4(reinterpret_cast<const float *>(p.dres_in) + row * p.channels + v * 4, r);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[i] += r[i];
                }
                st_rt<4>(p.dx, (int)p.dx_dtype, row * p.dx_row_stride + v * 4, d);
                if (p.dres_out != nullptr) Io<float>::st4(reinterpret_cast<float *>(p.dres_out) + row * p.channels + v * 4, d);
            }
        }
    }
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp][(lane + it * 32) * 4 + i] = dw[it][i];
    __syncthreads();
    float *dst = reinterpret_cast<float *>(p.dweight_partial) + (int64_t)blockIdx.x * p.channels;
    for (int c = threadIdx.x; c < p.channels; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int wp = 0; wp < 8; ++wp) t += red[wp][c];
        dst[c] = t;
    }
}

int rowwise_entry(const dimsum_rowwise_params *p, bool gate, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const char *who = gate ? "gate_residual" : "modulate";
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    DIMSUM_REQUIRE(p != nullptr && p->x && p->dst, DIMSUM_ERR_INVALID, "%s: null pointer", who);
    DIMSUM_REQUIRE(gate ? (p->m && p->gate) : (p->shift && p->scale), DIMSUM_ERR_INVALID, "%s: null operand", who);
    DIMSUM_REQUIRE(p->batch >= 0 && p->seqlen > 0 && p->channels > 0, DIMSUM_ERR_INVALID, "%s: bad sizes", who);
    auto dt_ok = [](int64_t d) { return d >= 0 && d <= 2; };
    DIMSUM_REQUIRE(dt_ok(p->x_dtype) && dt_ok(p->aux_dtype) && dt_ok(p->dst_dtype), DIMSUM_ERR_INVALID, "%s: unknown dtype", who);
    bool ok = p->channels % 8 == 0 && aligned16(p->x) && aligned16(p->dst) && p->x_batch_stride % 8 == 0 &&
              p->x_token_stride % 8 == 0 && p->dst_batch_stride % 8 == 0 && p->dst_token_stride % 8 == 0 &&
              p->vec_row_stride % 8 == 0;
    if (gate) ok = ok && aligned16(p->m) && aligned16(p->gate) && p->m_batch_stride % 8 == 0 && p->m_token_stride % 8 == 0;
    else ok = ok && aligned16(p->shift) && aligned16(p->scale);
    DIMSUM_REQUIRE(ok, DIMSUM_ERR_UNSUPPORTED, "%s: rows must be 16-byte aligned and channels a multiple of 8", who);
    DIMSUM_REQUIRE(p->dst != p->x || p->idx == nullptr || gate, DIMSUM_ERR_INVALID, "%s: in-place gather is not supported", who);
    if (p->batch == 0) return DIMSUM_OK;
    const int64_t per_batch = p->seqlen * (p->channels / 8);
    DIMSUM_REQUIRE(per_batch < ((int64_t)1 << 31) && p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED,
                   "%s: more than 65535 batch rows or 2^31 vectors per batch row", who);
    const dim3 blocks((unsigned)((per_batch + 255) / 256), (unsigned)p->batch);
    if (gate) rowwise_kernel<true><<<blocks, 256, 0, stream>>>(*p);
    else rowwise_kernel<false><<<blocks, 256, 0, stream>>>(*p);
    return check_launch(who);
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_modulate(const dimsum_rowwise_params *p, void *stream) { return rowwise_entry(p, false, stream); }
extern "C" int dimsum_gate_residual(const dimsum_rowwise_params *p, void *stream) { return rowwise_entry(p, true, stream); }

extern "C" int dimsum_add_rmsnorm(const dimsum_rmsnorm_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "add_rmsnorm: null params");
    if (p->rows == 0) return DIMSUM_OK;
    dimsum_norm_modulate_params q{};
    q.rows = p->rows; q.channels = p->channels; q.rows_per_batch = 1;
    q.x_dtype = p->dtype; q.y_dtype = p->dtype; q.aux_dtype = DIMSUM_F32; q.norm_kind = 0;
    q.x_row_stride = p->x_row_stride; q.y_row_stride = p->y_row_stride; q.vec_row_stride = 0;
    q.x = p->x; q.residual = p->residual; q.weight = p->weight; q.shift = nullptr; q.scale = nullptr;
    q.y = p->y; q.res_out = p->res_out; q.eps = p->eps;
    DIMSUM_REQUIRE(p->dtype >= 0 && p->dtype <= 2, DIMSUM_ERR_INVALID, "add_rmsnorm: bad arguments");
    const int vec = p->dtype == DIMSUM_F32 ? 4 : 8;
    DIMSUM_REQUIRE(p->y_row_stride % vec == 0, DIMSUM_ERR_UNSUPPORTED, "add_rmsnorm: rows must be 16-byte aligned");
    return run_norm(&q, stream, "add_rmsnorm");
}

extern "C" int dimsum_norm_modulate(const dimsum_norm_modulate_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->rows == 0) return DIMSUM_OK;
    return run_norm(p, stream, "norm_modulate");
}

extern "C" int dimsum_gelu_mul(const dimsum_gelu_mul_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->rows == 0) return DIMSUM_OK;
    DIMSUM_REQUIRE(p != nullptr && p->x && p->y, DIMSUM_ERR_INVALID, "gelu_mul: null pointer");
    DIMSUM_REQUIRE(p->rows >= 0 && p->hidden > 0 && p->dtype >= 0 && p->dtype <= 2, DIMSUM_ERR_INVALID, "gelu_mul: bad arguments");
    const int vec = p->dtype == DIMSUM_F32 ? 4 : 8;
    DIMSUM_REQUIRE(p->hidden % vec == 0 && aligned16(p->x) && aligned16(p->y) && p->x_row_stride % vec == 0 &&
                       p->y_row_stride % vec == 0, DIMSUM_ERR_UNSUPPORTED, "gelu_mul: rows must be 16-byte aligned");
    if (p->rows == 0) return DIMSUM_OK;
    const bool two = (p->hidden / vec) % 2 == 0;
    const int64_t total = p->rows * (p->hidden / (vec * (two ? 2 : 1)));
    const unsigned blocks = (unsigned)((total + 255) / 256);
#define GM(T)                                                                       \
    if (two) gelu_mul_kernel<T, 2><<<blocks, 256, 0, stream>>>(*p);                 \
    else gelu_mul_kernel<T, 1><<<blocks, 256, 0, stream>>>(*p);
    if (p->dtype == DIMSUM_F32) { GM(float) }
    else if (p->dtype == DIMSUM_BF16) { GM(__nv_bfloat16) }
    else { GM(__half) }
#undef GM
    return check_launch("gelu_mul");
}

extern "C" int dimsum_cfg_euler_step(const dimsum_cfg_euler_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->half_batch == 0) return DIMSUM_OK;
    DIMSUM_REQUIRE(p != nullptr && p->model_out && p->x && p->x_new && p->dt, DIMSUM_ERR_INVALID, "cfg_euler_step: null pointer");
    DIMSUM_REQUIRE(p->half_batch > 0 && p->channels > 0 && p->hw > 0 && p->out_dtype >= 0 && p->out_dtype <= 2, DIMSUM_ERR_INVALID,
                   "cfg_euler_step: bad arguments");
    DIMSUM_REQUIRE((p->channels * p->hw) % 4 == 0 && p->out_row_stride % 4 == 0 && aligned16(p->x) && aligned16(p->x_new) &&
                       (reinterpret_cast<uintptr_t>(p->model_out) & (p->out_dtype == DIMSUM_F32 ? 15u : 7u)) == 0,
                   DIMSUM_ERR_UNSUPPORTED, "cfg_euler_step: latents must be 16-byte aligned with a multiple of 4 elements");
    const int64_t total = p->half_batch * p->channels * p->hw / 4;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (p->out_dtype == DIMSUM_F32) cfg_euler_kernel<float><<<blocks, 256, 0, stream>>>(*p);
    else if (p->out_dtype == DIMSUM_BF16) cfg_euler_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(*p);
    else cfg_euler_kernel<__half><<<blocks, 256, 0, stream>>>(*p);
    return check_launch("cfg_euler_step");
}

extern "C" int dimsum_token_colsum(const dimsum_colsum_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;
    DIMSUM_REQUIRE(p != nullptr && p->g && (p->sum_g || p->sum_gx), DIMSUM_ERR_INVALID, "token_colsum: null pointer");
    DIMSUM_REQUIRE(p->sum_gx == nullptr || p->x != nullptr, DIMSUM_ERR_INVALID, "token_colsum: sum_gx needs x");
    DIMSUM_REQUIRE(p->batch > 0 && p->seqlen > 0 && p->channels > 0, DIMSUM_ERR_INVALID, "token_colsum: bad sizes");
    auto dt_ok = [](int64_t d) { return d >= 0 && d <= 2; };
    DIMSUM_REQUIRE(dt_ok(p->g_dtype) && dt_ok(p->out_dtype) && (p->x == nullptr || dt_ok(p->x_dtype)), DIMSUM_ERR_INVALID,
                   "token_colsum: unknown dtype");
    DIMSUM_REQUIRE(p->channels % 4 == 0 && aligned16(p->g) && p->g_batch_stride % 4 == 0 && p->g_token_stride % 4 == 0 &&
                       (p->x == nullptr || (aligned16(p->x) && p->x_batch_stride % 4 == 0 && p->x_token_stride % 4 == 0)),
                   DIMSUM_ERR_UNSUPPORTED, "token_colsum: rows must be 16-byte aligned and channels a multiple of 4");
    DIMSUM_REQUIRE(p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "token_colsum: batch > 65535");
    const dim3 grid((unsigned)((p->channels + 127) / 128), (unsigned)p->batch);
    if (p->sum_gx != nullptr) colsum_kernel<true><<<grid, kColsumWarps * 32, 0, stream>>>(*p);
    else colsum_kernel<false><<<grid, kColsumWarps * 32, 0, stream>>>(*p);
    return check_launch("token_colsum");
}

extern "C" int dimsum_gelu_mul_bwd(const dimsum_gelu_mul_bwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->rows == 0) return DIMSUM_OK;
    DIMSUM_REQUIRE(p != nullptr && p->x && p->dy && p->dx, DIMSUM_ERR_INVALID, "gelu_mul_bwd: null pointer");
    DIMSUM_REQUIRE(p->rows > 0 && p->hidden > 0 && p->dtype >= 0 && p->dtype <= 2, DIMSUM_ERR_INVALID, "gelu_mul_bwd: bad arguments");
    const int vec = p->dtype == DIMSUM_F32 ? 4 : 8;
    DIMSUM_REQUIRE(p->hidden % vec == 0 && aligned16(p->x) && aligned16(p->dy) && aligned16(p->dx) && p->x_row_stride % vec == 0 &&
                       p->dy_row_stride % vec == 0 && p->dx_row_stride % vec == 0,
                   DIMSUM_ERR_UNSUPPORTED, "gelu_mul_bwd: rows must be 16-byte aligned");
    const unsigned tpr = (unsigned)(p->hidden / vec);
    const unsigned gy = (unsigned)(p->rows < 65535 ? p->rows : 65535), gz = (unsigned)((p->rows + gy - 1) / gy);
    DIMSUM_REQUIRE(gz <= 65535, DIMSUM_ERR_UNSUPPORTED, "gelu_mul_bwd: too many rows");
    const dim3 grid((tpr + 255) / 256, gy, gz);
    if (p->dtype == DIMSUM_F32) gelu_mul_bwd_kernel<float><<<grid, 256, 0, stream>>>(*p);
    else if (p->dtype == DIMSUM_BF16) gelu_mul_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(*p);
    else gelu_mul_bwd_kernel<__half><<<grid, 256, 0, stream>>>(*p);
    return check_launch("gelu_mul_bwd");
}

extern "C" int dimsum_add_rmsnorm_bwd(const dimsum_rmsnorm_bwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr && p->h && p->weight && p->dy && p->dx && p->dweight_partial, DIMSUM_ERR_INVALID,
                   "add_rmsnorm_bwd: null pointer");
    DIMSUM_REQUIRE(p->rows > 0 && p->channels > 0 && p->n_partials > 0, DIMSUM_ERR_INVALID, "add_rmsnorm_bwd: bad sizes");
    DIMSUM_REQUIRE(p->dy_dtype >= 0 && p->dy_dtype <= 2 && p->dx_dtype >= 0 && p->dx_dtype <= 2, DIMSUM_ERR_INVALID,
                   "add_rmsnorm_bwd: unknown dtype");
    DIMSUM_REQUIRE(p->channels % 4 == 0 && p->channels <= 1024, DIMSUM_ERR_UNSUPPORTED,
                   "add_rmsnorm_bwd: channels=%lld must be a multiple of 4 and at most 1024", (long long)p->channels);
    DIMSUM_REQUIRE(aligned16(p->h) && aligned16(p->weight) && aligned16(p->dy) && aligned16(p->dx) && p->dy_row_stride % 4 == 0 && p->dx_row_stride % 4 == 0 &&
                       (p->dres_in == nullptr || aligned16(p->dres_in)) && (p->dres_out == nullptr || aligned16(p->dres_out)),
                   DIMSUM_ERR_UNSUPPORTED, "add_rmsnorm_bwd: rows must be 16-byte aligned");
    const unsigned blocks = (unsigned)p->n_partials;
    const int per_lane = (int)((p->channels / 4 + 31) / 32);
    if (per_lane <= 2) rmsnorm_bwd_kernel<2><<<blocks, 256, 0, stream>>>(*p);
    else if (per_lane <= 4) rmsnorm_bwd_kernel<4><<<blocks, 256, 0, stream>>>(*p);
    else rmsnorm_bwd_kernel<8><<<blocks, 256, 0, stream>>>(*p);
    return check_launch("add_rmsnorm_bwd");
}
