// Token gather and the fused 2-level Haar wavelet-packet split / merge for sm_100a.
//
// Replaces WaveDiMBlock._dwt_fast / _idwt_fast (dimsum/models_dim.py:572-604: 8 grouped conv2d + 2 cat + chunk
// reorder + 2 rearranges forward, 2 conv_transpose2d + the same glue inverse -- about ten full-tensor passes) and
// the local_scan / local_reverse copies around it (dimsum/scanning_orders.py:347-416) with ONE pass each way.
//
// Because level 2 is applied to all four level-1 sub-bands, the transform is a 16-point +-1 butterfly over every
// 4x4 patch of the token image, per channel, and its 16 outputs land on the SAME 16 tokens:
//     coef[k1][k2] = 1/16 * sum_{j1,j2} S[k1][j1] S[k2][j2] X[pixel(j1,j2)]          S = Haar sign matrix
//     OUT[token (4h+p1, 4w+p2), channel (4 k1 + k2) * C/16 + c/16] = coef,   p1 = (c % 16) / 4, p2 = c % 4.
// So a CTA owns one (batch, 4x4 token patch): 16 rows of C channels in, 16 rows out, everything coalesced, the
// channel scramble done through a padded shared tile.  `pos` places / finds each token at its window-scan
// sequence position, which fuses local_scan (forward) and local_reverse (inverse) into the store / load.
#include "common.cuh"

namespace dimsum {
namespace {

struct WaveArgs {
    const void *src;
    void *dst;
    const int32_t *pos;
    int64_t s_bs, s_ts, d_bs, d_ts;
    int grid, channels;
    float scale;
};

// 4-point Haar sign transform (ll, lh, hl, hh) of (p, q, r, s); self-inverse up to a factor 4.
DEV void haar4(float &p, float &q, float &r, float &s) {
    const float a = p + q, b = r + s, c = p - q, d = r - s;
    p = a + b; q = a - b; r = c + d; s = c - d;
}

// v[4*j2 + j1] (j1 = position inside the level-1 2x2 block, j2 = which block) -> v[4*k1 + k2]
DEV void packet16(float (&v)[16]) {
#pragma unroll
    for (int j2 = 0; j2 < 4; ++j2) haar4(v[4 * j2], v[4 * j2 + 1], v[4 * j2 + 2], v[4 * j2 + 3]);   // level 1: j1 -> k1
    float o[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float a = v[k1], b = v[4 + k1], c = v[8 + k1], d = v[12 + k1];                                // level 2: j2 -> k2
        haar4(a, b, c, d);
        o[4 * k1] = a; o[4 * k1 + 1] = b; o[4 * k1 + 2] = c; o[4 * k1 + 3] = d;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// token index inside the 4x4 patch for (j2, j1): pixel row 2*i2 + i1, col 2*jj2 + jj1
DEV int patch_pixel(int j2, int j1) {
    const int row = 2 * (j2 >> 1) + (j1 >> 1), col = 2 * (j2 & 1) + (j1 & 1);
    return row * 4 + col;
}

constexpr int kWaveThreads = 256;   // upper bound; the launch uses one thread per channel (group of 4), so none idles in the butterfly

// 4 consecutive channels: one 16-byte (fp32) or 8-byte (16-bit) access
template <typename T> DEV void ld4(const T *p, float (&v)[4]);
template <typename T> DEV void st4(T *p, const float (&v)[4]);
template <> DEV void ld4<float>(const float *p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4 *>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> DEV void st4<float>(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> DEV void ld4<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <> DEV void st4<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2 *>(p) = make_uint2(reinterpret_cast<uint32_t &>(a), reinterpret_cast<uint32_t &>(b));
}
template <> DEV void ld4<__half>(const __half *p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __half22float2(reinterpret_cast<const __half2 &>(r.x)), b = __half22float2(reinterpret_cast<const __half2 &>(r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> DEV void st4<__half>(__half *p, const float (&v)[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2 *>(p) = make_uint2(reinterpret_cast<uint32_t &>(a), reinterpret_cast<uint32_t &>(b));
}

// kVec4: every thread moves 4 consecutive channels per access (needs 16-byte aligned fp32 rows / 8-byte aligned 16-bit rows)
template <typename T, bool kInverse, bool kVec4>
__global__ void __launch_bounds__(kWaveThreads, 3) wavelet_kernel(const WaveArgs a) {
    extern __shared__ __align__(16) float tile[];   // [16][C + 4] fp32
    constexpr int CH = kVec4 ? 4 : 1;
    const int C = a.channels, pitch = C + 4, Cq = C / 16;
    const int g = a.grid / 4;
    const int b = blockIdx.y;
    const int ph = blockIdx.x / g, pw = blockIdx.x % g;
    const T *src = reinterpret_cast<const T *>(a.src) + b * a.s_bs;
    T *dst = reinterpret_cast<T *>(a.dst) + b * a.d_bs;

    auto token_of = [&](int t16) { return (ph * 4 + (t16 >> 2)) * a.grid + pw * 4 + (t16 & 3); };
    auto seq_of = [&](int t16) { const int tok = token_of(t16); return a.pos != nullptr ? a.pos[tok] : tok; };
    auto ldc = [&](const T *p, float (&v)[CH]) {
        if constexpr (kVec4) ld4<T>(p, v); else v[0] = Io<T>::ld(p);
    };
    auto stc = [&](T *p, const float (&v)[CH]) {
        if constexpr (kVec4) st4<T>(p, v); else Io<T>::st(p, v[0]);
    };

    if (!kInverse) {
        // image tokens -> butterfly -> tile[p1p2][k * Cq + c / 16] -> coefficient tokens at pos[token]
        for (int c0 = threadIdx.x * CH; c0 < C; c0 += blockDim.x * CH) {
            float px[16][CH];
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2)
#pragma unroll
                for (int j1 = 0; j1 < 4; ++j1)
                    ldc(src + (int64_t)token_of(patch_pixel(j2, j1)) * a.s_ts + c0, px[4 * j2 + j1]);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = px[k][i];
                packet16(v);
                const int c = c0 + i, t16 = c & 15, cq = c >> 4;
#pragma unroll
                for (int k = 0; k < 16; ++k) tile[t16 * pitch + k * Cq + cq] = v[k] * a.scale;
            }
        }
        __syncthreads();
        // copy-out: a thread keeps its channel group and walks the 16 token rows (no div / mod, 16 independent LDS + STG)
        for (int c0 = threadIdx.x * CH; c0 < C; c0 += blockDim.x * CH) {
#pragma unroll
            for (int t16 = 0; t16 < 16; ++t16) {
                float v[CH];
                if constexpr (kVec4) {
                    const float4 r = *reinterpret_cast<const float4 *>(&tile[t16 * pitch + c0]);
                    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
                } else {
                    v[0] = tile[t16 * pitch + c0];
                }
                stc(dst + (int64_t)seq_of(t16) * a.d_ts + c0, v);
            }
        }
    } else {
        // all 16 token rows of a channel group are requested before the first one is parked (16 loads in flight per thread)
        for (int c0 = threadIdx.x * CH; c0 < C; c0 += blockDim.x * CH) {
            float v[16][CH];
#pragma unroll
            for (int t16 = 0; t16 < 16; ++t16) ldc(src + (int64_t)seq_of(t16) * a.s_ts + c0, v[t16]);
#pragma unroll
            for (int t16 = 0; t16 < 16; ++t16) {
                if constexpr (kVec4) {
                    *reinterpret_cast<float4 *>(&tile[t16 * pitch + c0]) = make_float4(v[t16][0], v[t16][1], v[t16][2], v[t16][3]);
                } else {
                    tile[t16 * pitch + c0] = v[t16][0];
                }
            }
        }
        __syncthreads();
        for (int c0 = threadIdx.x * CH; c0 < C; c0 += blockDim.x * CH) {
            float px[16][CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = c0 + i, t16 = c & 15, cq = c >> 4;
                // inverse = the same sign butterfly with the roles of (k1,k2) and (j1,j2) exchanged
                float tmp[16];
#pragma unroll
                for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) tmp[4 * k2 + k1] = tile[t16 * pitch + (4 * k1 + k2) * Cq + cq];
                packet16(tmp);   // tmp[4*j1 + j2]
#pragma unroll
                for (int j1 = 0; j1 < 4; ++j1)
#pragma unroll
                    for (int j2 = 0; j2 < 4; ++j2) px[patch_pixel(j2, j1)][i] = tmp[4 * j1 + j2] * a.scale;
                // keep the four channels' gathers from being hoisted together (136 registers -> 3 CTAs/SM otherwise)
                asm volatile("" ::: "memory");
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) stc(dst + (int64_t)token_of(t) * a.d_ts + c0, px[t]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const T *src, T *dst, const int32_t *index, int64_t s_bs, int64_t s_ts,
                                                     int64_t d_bs, int64_t d_ts, int seqlen, int vecs_per_token) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;        // 32-bit inside a batch row, blockIdx.y = batch
    if (idx >= (unsigned)seqlen * (unsigned)vecs_per_token) return;
    const unsigned l = idx / (unsigned)vecs_per_token;
    const unsigned v = idx - l * (unsigned)vecs_per_token;
    const int64_t b = blockIdx.y;
    const uint4 val = *reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(src + b * s_bs + (int64_t)index[l] * s_ts) + 16 * v);
    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(dst + b * d_bs + (int64_t)l * d_ts) + 16 * v) = val;
}

template <typename T>
int run_wavelet(const dimsum_wavelet_params *p, bool inverse, cudaStream_t stream) {
    WaveArgs a;
    a.src = p->src; a.dst = p->dst; a.pos = p->pos;
    a.s_bs = p->src_batch_stride; a.s_ts = p->src_token_stride; a.d_bs = p->dst_batch_stride; a.d_ts = p->dst_token_stride;
    a.grid = (int)p->grid; a.channels = (int)p->channels; a.scale = p->scale;
    const int g = a.grid / 4;
    const int smem = 16 * (a.channels + 4) * (int)sizeof(float);
    dim3 grid(g * g, (unsigned)p->batch);
    const uintptr_t align = 4 * sizeof(T) - 1;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(p->src) | reinterpret_cast<uintptr_t>(p->dst)) & align) == 0 &&
                      a.s_bs % 4 == 0 && a.s_ts % 4 == 0 && a.d_bs % 4 == 0 && a.d_ts % 4 == 0;
    auto go = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int per = vec4 ? 4 : 1;
        const int threads = min(kWaveThreads, max(64, ((a.channels + per - 1) / per + 31) / 32 * 32));
        kern<<<grid, threads, smem, stream>>>(a);
    };
    if (inverse) {
        if (vec4) go(wavelet_kernel<T, true, true>); else go(wavelet_kernel<T, true, false>);
    } else {
        if (vec4) go(wavelet_kernel<T, false, true>); else go(wavelet_kernel<T, false, false>);
    }
    return check_launch(inverse ? "wavelet_packet_inv" : "wavelet_packet_fwd");
}

int wavelet_entry(const dimsum_wavelet_params *p, bool inverse, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const char *who = inverse ? "wavelet_packet_inv" : "wavelet_packet_fwd";
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    DIMSUM_REQUIRE(p != nullptr && p->src && p->dst, DIMSUM_ERR_INVALID, "%s: null pointer", who);
    DIMSUM_REQUIRE(p->grid > 0 && p->grid % 4 == 0, DIMSUM_ERR_INVALID, "%s: token grid %lld must be a multiple of 4", who,
                   (long long)p->grid);
    DIMSUM_REQUIRE(p->channels > 0 && p->channels % 16 == 0, DIMSUM_ERR_INVALID, "%s: channels %lld must be a multiple of 16",
                   who, (long long)p->channels);
    DIMSUM_REQUIRE(16 * (p->channels + 4) * 4 <= 200 * 1024, DIMSUM_ERR_UNSUPPORTED, "%s: channels %lld too many", who,
                   (long long)p->channels);
    DIMSUM_REQUIRE(p->batch >= 0 && p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "%s: batch out of range", who);
    DIMSUM_REQUIRE(p->src != p->dst, DIMSUM_ERR_INVALID, "%s: in-place operation is not supported", who);
    if (p->batch == 0) return DIMSUM_OK;
    switch (p->dtype) {
        case DIMSUM_F32: return run_wavelet<float>(p, inverse, stream);
        case DIMSUM_BF16: return run_wavelet<__nv_bfloat16>(p, inverse, stream);
        case DIMSUM_F16: return run_wavelet<__half>(p, inverse, stream);
        default: return fail(DIMSUM_ERR_INVALID, "%s: unknown dtype", who);
    }
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_wavelet_packet_fwd(const dimsum_wavelet_params *p, void *stream) { return wavelet_entry(p, false, stream); }
extern "C" int dimsum_wavelet_packet_inv(const dimsum_wavelet_params *p, void *stream) { return wavelet_entry(p, true, stream); }

extern "C" int dimsum_token_gather(const dimsum_gather_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    DIMSUM_REQUIRE(p != nullptr && p->src && p->dst && p->index, DIMSUM_ERR_INVALID, "token_gather: null pointer");
    DIMSUM_REQUIRE(p->batch >= 0 && p->seqlen > 0 && p->channels > 0, DIMSUM_ERR_INVALID, "token_gather: bad sizes");
    DIMSUM_REQUIRE(p->dtype >= 0 && p->dtype <= 2, DIMSUM_ERR_INVALID, "token_gather: unknown dtype");
    const int esz = p->dtype == DIMSUM_F32 ? 4 : 2;
    const int vec = 16 / esz;
    DIMSUM_REQUIRE(p->channels % vec == 0 && aligned16(p->src) && aligned16(p->dst) && p->src_token_stride % vec == 0 &&
                       p->dst_token_stride % vec == 0 && p->src_batch_stride % vec == 0 && p->dst_batch_stride % vec == 0,
                   DIMSUM_ERR_UNSUPPORTED, "token_gather: rows must be 16-byte aligned multiples of 16 bytes");
    DIMSUM_REQUIRE(p->src != p->dst, DIMSUM_ERR_INVALID, "token_gather: in-place operation is not supported");
    if (p->batch == 0) return DIMSUM_OK;
    const int vpt = (int)(p->channels / vec);
    const int64_t per_batch = p->seqlen * vpt;
    DIMSUM_REQUIRE(per_batch < ((int64_t)1 << 31) && p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED,
                   "token_gather: more than 65535 batch rows or 2^31 vectors per batch row");
    const dim3 blocks((unsigned)((per_batch + 255) / 256), (unsigned)p->batch);
    if (esz == 4) {
        gather_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float *>(p->src), reinterpret_cast<float *>(p->dst),
                                                        p->index, p->src_batch_stride, p->src_token_stride, p->dst_batch_stride,
                                                        p->dst_token_stride, (int)p->seqlen, vpt);
    } else {
        gather_kernel<uint16_t><<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint16_t *>(p->src),
                                                           reinterpret_cast<uint16_t *>(p->dst), p->index, p->src_batch_stride,
                                                           p->src_token_stride, p->dst_batch_stride, p->dst_token_stride,
                                                           (int)p->seqlen, vpt);
    }
    return check_launch("token_gather");
}
