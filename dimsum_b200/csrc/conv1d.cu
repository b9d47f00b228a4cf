// Causal depthwise conv1d (+SiLU), forward and backward, for sm_100a.
//
// Replaces causal_conv1d_fwd_kernel / causal_conv1d_bwd_kernel (causal-conv1d/csrc/causal_conv1d_fwd.cu:39-130,
// causal_conv1d_bwd.cu:46-240).  The reference gives one 128-thread CTA to every 256-element row, so at the
// model's L=256 half to three quarters of its threads idle; here the (row, 16-byte vector) pairs of the whole
// tensor are flattened over the grid, every thread owns one vector, and the 3-element halo travels by warp
// shuffle (lane 0 and row starts re-read it through L1).  Pure streaming: 2 * R*D*L*s bytes.
#include "common.cuh"

namespace dimsum {
namespace {

constexpr int kMaxW = 4;

struct ConvArgs {
    const void *x, *weight, *bias, *dout;
    void *out, *dx;
    float *dweight, *dbias;
    const int32_t *perm;
    int64_t x_bs, x_ds, o_bs, o_ds, g_bs, g_ds, dx_bs, dx_ds, w_ds, w_ws;
    int batch, dim, seqlen, width, silu, w_dtype, vec_ok, out_vec_ok;
};

DEV float load_w(const void *w, int dtype, int64_t idx) {
    if (dtype == DIMSUM_F32) return reinterpret_cast<const float *>(w)[idx];
    if (dtype == DIMSUM_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(w)[idx]);
    return __half2float(reinterpret_cast<const __half *>(w)[idx]);
}

// ---------------------------------------------------------------------------------------------- forward
// kNV consecutive 16-byte vectors per thread (2 when the row length allows it: twice the bytes in flight per thread,
// which is what a pure streaming kernel at full occupancy needs to cover HBM latency on B200).
// Indexing is 32-bit inside a batch row (blockIdx.y walks the batch): the 64-bit div / mod chain of a flat index cost more
// instructions than the convolution itself and made the 16-bit variant issue-bound.
template <typename T, int kNV, bool kSilu>
__global__ void __launch_bounds__(256) conv_fwd_vec_kernel(const ConvArgs a) {
    constexpr int VEC = Io<T>::kVec;
    constexpr int W = kNV * VEC;                      // elements per thread
    const unsigned tpr = (unsigned)a.seqlen / W;      // threads per row
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = idx < (unsigned)a.dim * tpr;
    const int lane = threadIdx.x & 31;
    const unsigned d = active ? idx / tpr : 0u;
    const unsigned v = active ? idx - d * tpr : 0u;
    float w[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW; ++i) {   // right-align the taps: w[3] multiplies x[l]
        const int wi = i - (kMaxW - a.width);
        w[i] = wi >= 0 ? load_w(a.weight, a.w_dtype, (int64_t)d * a.w_ds + wi * a.w_ws) : 0.f;
    }
    const float bias = a.bias != nullptr ? load_w(a.bias, a.w_dtype, d) : 0.f;
    for (int b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const T *xr = reinterpret_cast<const T *>(a.x) + (int64_t)b * a.x_bs + (int64_t)d * a.x_ds + v * W;
        float xv[W + kMaxW - 1];   // [0..2] halo, [3..] own elements
#pragma unroll
        for (int i = 0; i < W + kMaxW - 1; ++i) xv[i] = 0.f;
        if (active) {
#pragma unroll
            for (int j = 0; j < kNV; ++j) Io<T>::ldv(xr + j * VEC, reinterpret_cast<float(&)[VEC]>(xv[kMaxW - 1 + j * VEC]));
        }
        // halo: last 3 elements of the previous thread of the same row
        const float h0 = __shfl_up_sync(0xffffffffu, xv[W + kMaxW - 4], 1);
        const float h1 = __shfl_up_sync(0xffffffffu, xv[W + kMaxW - 3], 1);
        const float h2 = __shfl_up_sync(0xffffffffu, xv[W + kMaxW - 2], 1);
        if (active && v > 0) {
            if (lane > 0) {
                xv[0] = h0; xv[1] = h1; xv[2] = h2;
            } else {
                xv[0] = Io<T>::ld(xr - 3);
                xv[1] = Io<T>::ld(xr - 2);
                xv[2] = Io<T>::ld(xr - 1);
            }
        }
        if (active) {
            T *orow = reinterpret_cast<T *>(a.out) + (int64_t)b * a.o_bs + (int64_t)d * a.o_ds + v * W;
#pragma unroll
            for (int j = 0; j < kNV; ++j) {
                float ov[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float acc = bias;
#pragma unroll
                    for (int k = 0; k < kMaxW; ++k) acc = fmaf(w[k], xv[j * VEC + i + k], acc);
                    ov[i] = kSilu ? silu_t<sizeof(T) == 2>(acc) : acc;
                }
                Io<T>::stv(orow + j * VEC, ov);
            }
        }
    }
}

// scalar / gathered path.  A thread produces kGW consecutive outputs of one row from kGW + 3 inputs fetched once each
// through the optional token order (4-byte gathers inside a row that stays in L1), and stores them with one vector
// write when the row is aligned: ~3x fewer loads than one thread per output.
constexpr int kGW = 8;

template <typename T>
__global__ void __launch_bounds__(256) conv_fwd_scalar_kernel(const ConvArgs a) {
    const unsigned tpr = (unsigned)(a.seqlen + kGW - 1) / kGW;   // threads per row
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit inside a batch row, blockIdx.y = batch
    if (idx >= (unsigned)a.dim * tpr) return;
    const unsigned d = idx / tpr;
    const int l0 = (int)(idx - d * tpr) * kGW;
    const int64_t b = blockIdx.y;
    const T *xr = reinterpret_cast<const T *>(a.x) + b * a.x_bs + (int64_t)d * a.x_ds;
    float xs[kGW + kMaxW - 1];
#pragma unroll
    for (int i = 0; i < kGW + kMaxW - 1; ++i) {
        const int ls = l0 - (kMaxW - 1) + i;
        float v = 0.f;
        if (ls >= 0 && ls < a.seqlen) v = Io<T>::ld(xr + (a.perm != nullptr ? a.perm[ls] : ls));
        xs[i] = v;
    }
    float w[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW; ++i) {
        const int wi = i - (kMaxW - a.width);
        w[i] = wi >= 0 ? load_w(a.weight, a.w_dtype, (int64_t)d * a.w_ds + wi * a.w_ws) : 0.f;
    }
    const float bias = a.bias != nullptr ? load_w(a.bias, a.w_dtype, d) : 0.f;
    float ov[kGW];
#pragma unroll
    for (int i = 0; i < kGW; ++i) {
        float acc = bias;
#pragma unroll
        for (int k = 0; k < kMaxW; ++k) acc = fmaf(w[k], xs[i + k], acc);
        ov[i] = a.silu ? silu_t<sizeof(T) == 2>(acc) : acc;
    }
    T *orow = reinterpret_cast<T *>(a.out) + b * a.o_bs + (int64_t)d * a.o_ds + l0;
    if (a.out_vec_ok && l0 + kGW <= a.seqlen) {
#pragma unroll
        for (int i = 0; i < kGW; i += Io<T>::kVec) Io<T>::stv(orow + i, reinterpret_cast<const float(&)[Io<T>::kVec]>(ov[i]));
    } else {
#pragma unroll
        for (int i = 0; i < kGW; ++i)
            if (l0 + i < a.seqlen) Io<T>::st(orow + i, ov[i]);
    }
}

// ---------------------------------------------------------------------------------------------- backward
// grid = (dim, batch slabs); a CTA owns one channel for kBatchPerCta batch rows so dweight / dbias are reduced
// in registers + one block reduction and leave as 5 atomics per CTA (the reference: 5 atomics per (b, d) row).
constexpr int kBwdThreads = 128;
constexpr int kBatchPerCta = 8;

// kNV = 16-byte vectors per thread on the vector path (2 when the row length allows: halves the halo shuffles and the
// address arithmetic per element -- the kernel is issue-bound), 0 = scalar path for unaligned / ragged rows
template <typename T, int kNV, bool kSilu>
__global__ void __launch_bounds__(kBwdThreads) conv_bwd_kernel(const ConvArgs a) {
    const int d = blockIdx.x;
    const int b0 = blockIdx.y * kBatchPerCta;
    const int nb = min(kBatchPerCta, a.batch - b0);
    const int L = a.seqlen, W = a.width;
    float w[kMaxW];
#pragma unroll
    for (int i = 0; i < kMaxW; ++i) {
        const int wi = i - (kMaxW - W);
        w[i] = wi >= 0 ? load_w(a.weight, a.w_dtype, d * a.w_ds + wi * a.w_ws) : 0.f;
    }
    const float bias = a.bias != nullptr ? load_w(a.bias, a.w_dtype, d) : 0.f;
    float dw[kMaxW] = {0.f, 0.f, 0.f, 0.f};
    float db = 0.f;
    // g = dout * act'(pre)
    auto gate = [&](float gout, float pre) {
        if (!kSilu) return gout;
        const float sg = sigmoid_f(pre);
        return gout * (sg * fmaf(pre, 1.f - sg, 1.f));
    };

    // g[l] = dout[l] * act'(pre[l]);  dx[l] = sum_j w[3-j] g[l+j];  dw[k] += g[l] x[l-3+k]
    if constexpr (kNV > 0) {
        // vector path: a thread owns kNV consecutive 16-byte vectors; x halo comes from the lane below, g halo from the lane above
        constexpr int VEC = Io<T>::kVec, E = kNV * VEC;
        const int ipr = L / E;                                   // items per row
        const int lane = threadIdx.x & 31;
        const int total = nb * ipr;
        for (int base = 0; base < total; base += kBwdThreads) {
            const int idx = base + threadIdx.x;
            const bool act = idx < total;
            const int bl = act ? idx / ipr : 0, v = act ? idx % ipr : 0;
            const int64_t xo = (int64_t)(b0 + bl) * a.x_bs + (int64_t)d * a.x_ds + v * E;
            const int64_t go = (int64_t)(b0 + bl) * a.g_bs + (int64_t)d * a.g_ds + v * E;
            const T *xr = reinterpret_cast<const T *>(a.x) + xo;
            const T *gr = reinterpret_cast<const T *>(a.dout) + go;
            float xs[E + kMaxW - 1], g[E + kMaxW - 1];
#pragma unroll
            for (int i = 0; i < E + kMaxW - 1; ++i) { xs[i] = 0.f; g[i] = 0.f; }
            if (act) {
#pragma unroll
                for (int j = 0; j < kNV; ++j) {
                    Io<T>::ldv(xr + j * VEC, reinterpret_cast<float(&)[VEC]>(xs[kMaxW - 1 + j * VEC]));
                    Io<T>::ldv(gr + j * VEC, reinterpret_cast<float(&)[VEC]>(g[j * VEC]));
                }
            }
            const float h0 = __shfl_up_sync(0xffffffffu, xs[E + kMaxW - 4], 1);
            const float h1 = __shfl_up_sync(0xffffffffu, xs[E + kMaxW - 3], 1);
            const float h2 = __shfl_up_sync(0xffffffffu, xs[E + kMaxW - 2], 1);
            if (act && v > 0) {
                if (lane > 0) { xs[0] = h0; xs[1] = h1; xs[2] = h2; }
                else { xs[0] = Io<T>::ld(xr - 3); xs[1] = Io<T>::ld(xr - 2); xs[2] = Io<T>::ld(xr - 1); }
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                if (kSilu) {
                    float pre = bias;
#pragma unroll
                    for (int k = 0; k < kMaxW; ++k) pre = fmaf(w[k], xs[i + k], pre);
                    g[i] = gate(g[i], pre);
                }
                db += g[i];
#pragma unroll
                for (int k = 0; k < kMaxW; ++k) dw[k] = fmaf(g[i], xs[i + k], dw[k]);
            }
            // first 3 g's of the next item of the same row
            const float n0 = __shfl_down_sync(0xffffffffu, g[0], 1);
            const float n1 = __shfl_down_sync(0xffffffffu, g[1], 1);
            const float n2 = __shfl_down_sync(0xffffffffu, g[2], 1);
            if (act && v + 1 < ipr) {
                if (lane < 31 && idx + 1 < total) {
                    g[E] = n0; g[E + 1] = n1; g[E + 2] = n2;
                } else {   // warp edge: recompute the neighbour's first three g's
#pragma unroll
                    for (int j = 0; j < kMaxW - 1; ++j) {
                        float pre = bias;
                        if (kSilu) {
#pragma unroll
                            for (int k = 0; k < kMaxW; ++k) pre = fmaf(w[k], Io<T>::ld(xr + E + j - (kMaxW - 1) + k), pre);
                        }
                        g[E + j] = gate(Io<T>::ld(gr + E + j), pre);
                    }
                }
            }
            if (act) {
                T *dxr = reinterpret_cast<T *>(a.dx) + (int64_t)(b0 + bl) * a.dx_bs + (int64_t)d * a.dx_ds + v * E;
#pragma unroll
                for (int j = 0; j < kNV; ++j) {
                    float dxv[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < kMaxW; ++k) acc = fmaf(w[kMaxW - 1 - k], g[j * VEC + i + k], acc);
                        dxv[i] = acc;
                    }
                    Io<T>::stv(dxr + j * VEC, dxv);
                }
            }
        }
    } else
    for (int64_t idx = threadIdx.x; idx < (int64_t)nb * L; idx += kBwdThreads) {
        const int bl = (int)(idx / L), l = (int)(idx % L);
        const T *xr = reinterpret_cast<const T *>(a.x) + (b0 + bl) * a.x_bs + d * a.x_ds;
        const T *gr = reinterpret_cast<const T *>(a.dout) + (b0 + bl) * a.g_bs + d * a.g_ds;
        float xs[2 * kMaxW - 1];   // x[l-3 .. l+3]
#pragma unroll
        for (int i = 0; i < 2 * kMaxW - 1; ++i) {
            const int ls = l - (kMaxW - 1) + i;
            xs[i] = (ls >= 0 && ls < L) ? Io<T>::ld(xr + ls) : 0.f;
        }
        float dxv = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxW; ++j) {   // output position l + j (j = 0 is this thread's own g)
            const int lo = l + j;
            if (lo >= L) break;
            float pre = bias;
            if (kSilu) {
#pragma unroll
                for (int k = 0; k < kMaxW; ++k) pre = fmaf(w[k], xs[j + k], pre);
            }
            const float g = gate(Io<T>::ld(gr + lo), pre);
            dxv = fmaf(w[kMaxW - 1 - j], g, dxv);
            if (j == 0) {
                db += g;
#pragma unroll
                for (int k = 0; k < kMaxW; ++k) dw[k] = fmaf(g, xs[k], dw[k]);
            }
        }
        Io<T>::st(reinterpret_cast<T *>(a.dx) + (b0 + bl) * a.dx_bs + d * a.dx_ds + l, dxv);
    }
    // block reduction of dw[0..3], db
    __shared__ float red[kBwdThreads / 32][kMaxW + 1];
    float vals[kMaxW + 1] = {dw[0], dw[1], dw[2], dw[3], db};
#pragma unroll
    for (int i = 0; i < kMaxW + 1; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < kMaxW + 1; ++i) red[threadIdx.x >> 5][i] = vals[i];
    }
    __syncthreads();
    if (threadIdx.x < kMaxW + 1) {
        float s = 0.f;
        for (int wp = 0; wp < kBwdThreads / 32; ++wp) s += red[wp][threadIdx.x];
        if (threadIdx.x < kMaxW) {
            const int wi = (int)threadIdx.x - (kMaxW - W);
            if (wi >= 0) atomicAdd(a.dweight + d * W + wi, s);
        } else if (a.dbias != nullptr) {
            atomicAdd(a.dbias + d, s);
        }
    }
}

template <typename T>
int run_fwd(const ConvArgs &a, bool vec_ok, cudaStream_t stream) {
    if (vec_ok && a.perm == nullptr) {
        const int vpr = a.seqlen / Io<T>::kVec;
        const int nv = vpr % 2 == 0 ? 2 : 1;
        const int64_t per_batch = (int64_t)a.dim * (vpr / nv);
        DIMSUM_REQUIRE(per_batch < ((int64_t)1 << 31), DIMSUM_ERR_UNSUPPORTED, "causal_conv1d_fwd: dim * seqlen too large");
        dim3 grid((unsigned)((per_batch + 255) / 256), (unsigned)min(a.batch, 65535));
        auto go = [&](auto kern) { kern<<<grid, 256, 0, stream>>>(a); };
        if (a.silu) { if (nv == 2) go(conv_fwd_vec_kernel<T, 2, true>); else go(conv_fwd_vec_kernel<T, 1, true>); }
        else { if (nv == 2) go(conv_fwd_vec_kernel<T, 2, false>); else go(conv_fwd_vec_kernel<T, 1, false>); }
    } else {
        const int64_t per_batch = (int64_t)a.dim * ((a.seqlen + kGW - 1) / kGW);
        DIMSUM_REQUIRE(per_batch < ((int64_t)1 << 31) && a.batch <= 65535, DIMSUM_ERR_UNSUPPORTED,
                       "causal_conv1d_fwd: more than 65535 batch rows or dim * seqlen too large on the scalar / gather path");
        conv_fwd_scalar_kernel<T><<<dim3((unsigned)((per_batch + 255) / 256), (unsigned)a.batch), 256, 0, stream>>>(a);
    }
    return check_launch("causal_conv1d_fwd");
}

template <typename T>
int run_bwd(const ConvArgs &a, cudaStream_t stream) {
    dim3 grid(a.dim, (a.batch + kBatchPerCta - 1) / kBatchPerCta);
    auto go = [&](auto kern) { kern<<<grid, kBwdThreads, 0, stream>>>(a); };
    const int nv = !a.vec_ok ? 0 : (a.seqlen % (2 * Io<T>::kVec) == 0 ? 2 : (a.seqlen % Io<T>::kVec == 0 ? 1 : 0));
    if (a.silu) {
        if (nv == 2) go(conv_bwd_kernel<T, 2, true>); else if (nv == 1) go(conv_bwd_kernel<T, 1, true>); else go(conv_bwd_kernel<T, 0, true>);
    } else {
        if (nv == 2) go(conv_bwd_kernel<T, 2, false>); else if (nv == 1) go(conv_bwd_kernel<T, 1, false>); else go(conv_bwd_kernel<T, 0, false>);
    }
    return check_launch("causal_conv1d_bwd");
}

int check_common(const char *who, int64_t batch, int64_t dim, int64_t seqlen, int64_t width, int64_t io, int64_t wd) {
    DIMSUM_REQUIRE(batch >= 0 && dim > 0 && seqlen > 0, DIMSUM_ERR_INVALID, "%s: bad sizes", who);
    DIMSUM_REQUIRE(width >= 2 && width <= 4, DIMSUM_ERR_INVALID, "causal_conv1d only supports width between 2 and 4");
    DIMSUM_REQUIRE(io >= 0 && io <= 2 && wd >= 0 && wd <= 2, DIMSUM_ERR_INVALID, "%s: unknown dtype", who);
    DIMSUM_REQUIRE(batch * dim < (int64_t)1 << 40, DIMSUM_ERR_INVALID, "%s: tensor too large", who);
    return DIMSUM_OK;
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_causal_conv1d_fwd(const dimsum_conv_fwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "causal_conv1d_fwd: null params");
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    int rc = check_common("causal_conv1d_fwd", p->batch, p->dim, p->seqlen, p->width, p->io_dtype, p->w_dtype);
    if (rc) return rc;
    DIMSUM_REQUIRE(p->x && p->weight && p->out, DIMSUM_ERR_INVALID, "causal_conv1d_fwd: null pointer");
    if (p->batch == 0) return DIMSUM_OK;
    ConvArgs a{};
    a.x = p->x; a.weight = p->weight; a.bias = p->bias; a.out = p->out; a.perm = p->perm;
    a.x_bs = p->x_batch_stride; a.x_ds = p->x_d_stride; a.o_bs = p->out_batch_stride; a.o_ds = p->out_d_stride;
    a.w_ds = p->w_d_stride; a.w_ws = p->w_width_stride;
    a.batch = (int)p->batch; a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.width = (int)p->width;
    a.silu = p->silu != 0; a.w_dtype = (int)p->w_dtype;
    const int vec = p->io_dtype == DIMSUM_F32 ? 4 : 8;
    const bool vec_ok = p->seqlen % vec == 0 && aligned16(p->x) && aligned16(p->out) && a.x_bs % vec == 0 &&
                        a.x_ds % vec == 0 && a.o_bs % vec == 0 && a.o_ds % vec == 0;
    a.out_vec_ok = aligned16(p->out) && a.o_bs % 8 == 0 && a.o_ds % 8 == 0;
    switch (p->io_dtype) {
        case DIMSUM_F32: return run_fwd<float>(a, vec_ok, stream);
        case DIMSUM_BF16: return run_fwd<__nv_bfloat16>(a, vec_ok, stream);
        default: return run_fwd<__half>(a, vec_ok, stream);
    }
}

extern "C" int dimsum_causal_conv1d_bwd(const dimsum_conv_bwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "causal_conv1d_bwd: null params");
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    int rc = check_common("causal_conv1d_bwd", p->batch, p->dim, p->seqlen, p->width, p->io_dtype, p->w_dtype);
    if (rc) return rc;
    DIMSUM_REQUIRE(p->x && p->weight && p->dout && p->dx && p->dweight, DIMSUM_ERR_INVALID, "causal_conv1d_bwd: null pointer");
    DIMSUM_REQUIRE(p->dim <= 2147483647 && (p->batch + kBatchPerCta - 1) / kBatchPerCta <= 65535, DIMSUM_ERR_UNSUPPORTED,
                   "causal_conv1d_bwd: grid too large");
    if (p->batch == 0) return DIMSUM_OK;
    ConvArgs a{};
    a.x = p->x; a.weight = p->weight; a.bias = p->bias; a.dout = p->dout; a.dx = p->dx;
    a.dweight = p->dweight; a.dbias = p->dbias;
    a.x_bs = p->x_batch_stride; a.x_ds = p->x_d_stride; a.g_bs = p->dout_batch_stride; a.g_ds = p->dout_d_stride;
    a.dx_bs = p->dx_batch_stride; a.dx_ds = p->dx_d_stride; a.w_ds = p->w_d_stride; a.w_ws = p->w_width_stride;
    a.batch = (int)p->batch; a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.width = (int)p->width;
    a.silu = p->silu != 0; a.w_dtype = (int)p->w_dtype;
    {
        const int vec = p->io_dtype == DIMSUM_F32 ? 4 : 8;
        a.vec_ok = aligned16(p->x) && aligned16(p->dout) && aligned16(p->dx) && a.x_bs % vec == 0 && a.x_ds % vec == 0 &&
                   a.g_bs % vec == 0 && a.g_ds % vec == 0 && a.dx_bs % vec == 0 && a.dx_ds % vec == 0;
    }
    switch (p->io_dtype) {
        case DIMSUM_F32: return run_bwd<float>(a, stream);
        case DIMSUM_BF16: return run_bwd<__nv_bfloat16>(a, stream);
        default: return run_bwd<__half>(a, stream);
    }
}
