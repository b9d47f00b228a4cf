// Causal conv1d + SiLU fused IN FRONT of the x_proj contraction, on tcgen05 tensor cores (sm_100a).
//
// Replaces, for the composite the model calls (MambaInnerFn*.forward, mamba/mamba_ssm/ops/selective_scan_interface.py:836-866):
//     conv1d_out = causal_conv1d_fwd(x, w, b, silu)                       (a full read + write of (B, D, L))
//     x_dbl      = F.linear(rearrange(conv1d_out, "b d l -> (b l) d"), x_proj_weight)     (a second full read of conv1d_out)
//     B, C       = rearrange(x_dbl[:, r:r+N], "(b l) n -> b 1 n l").contiguous(), ...    (two more small passes)
// with ONE kernel that reads x once, writes u = conv1d_out once (the scan and the backward need it), and produces
// x_dbl ALREADY channel-major, (batch, dt_rank + 2 N, L): B and C are views of it in the layout the scan wants and dt is
// the right operand of the dt_proj GEMM, so both rearrange copies disappear.
//
// Design: a CTA owns one batch row x 128 tokens.  It walks the D channels in chunks of 128 bytes of K (32 fp32 or 64
// 16-bit channels).  For a chunk, every thread convolves a (VEC channels x VEC tokens) block held in registers -- the
// register-level transpose that turns the channel-major rows of x into the token-major (K-major) operand rows the tensor
// core wants -- stores its u rows to HBM with 16-byte stores and its operand rows to a SWIZZLE_128B shared tile
// (conflict-free: the 8 lanes of a quarter warp own the 8 chunks of one row).  One elected thread then issues
// tcgen05.mma (M = 128 tokens, N = dt_rank + 2 N, K = 8 / 16 per instruction) into a TMEM accumulator and commits to an
// mbarrier; the tiles are double-buffered, so the tensor core works on chunk c while the CTA produces chunk c + 1.
// fp32 I/O: kind::tf32, either one pass (TF32, what cuBLAS does under allow_tf32) or the 3xTF32 split (fp32-grade, 1e-6);
// 16-bit I/O: kind::f16 on the bf16 / fp16 values the reference's GEMM would see.  The epilogue reads the accumulator with
// tcgen05.ld (thread = token) and stores (batch, n_out, L) with coalesced rows.
#include "common.cuh"
#include "umma.cuh"

namespace dimsum {
namespace {

constexpr int kTok = 128;          // tokens per CTA == MMA M == threads per CTA
constexpr int kStages = 2;
constexpr int kMaxW = 4;

struct ConvXprojArgs {
    const void *x, *cw, *cb, *xw;
    void *u, *xdbl, *tail;
    int64_t x_bs, x_ds, u_bs, u_ds, o_bs, o_rs, t_bs, t_rs, cw_ds, cw_ws, xw_rs;
    int dim, seqlen, width, n_out, split, w_dtype, precise;
};

DEV float ld_w(const void *w, int dtype, int64_t idx) {
    if (dtype == DIMSUM_F32) return reinterpret_cast<const float *>(w)[idx];
    if (dtype == DIMSUM_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(w)[idx]);
    return __half2float(reinterpret_cast<const __half *>(w)[idx]);
}

template <typename T> struct OpTraits;
template <> struct OpTraits<float> { static constexpr bool kTf32 = true; static constexpr int kFmt = umma::kFmtTF32; };
template <> struct OpTraits<__nv_bfloat16> { static constexpr bool kTf32 = false; static constexpr int kFmt = umma::kFmtBF16; };
template <> struct OpTraits<__half> { static constexpr bool kTf32 = false; static constexpr int kFmt = umma::kFmtF16; };

// pack two fp32 values into one 32-bit word of the 16-bit storage type (low half = first)
template <typename T> DEV uint32_t pack2(float a, float b);
template <> DEV uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return reinterpret_cast<uint32_t &>(h);
}
template <> DEV uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return reinterpret_cast<uint32_t &>(h);
}

// kPrecise: 3xTF32 (fp32 I/O only)
template <typename T, bool kPrecise>
__global__ void __launch_bounds__(kTok, 2) conv_xproj_kernel(const ConvXprojArgs a) {
    constexpr int VEC = Io<T>::kVec;                    // 4 (fp32) / 8 (16-bit): elements per 16 bytes
    constexpr int KC = 128 / (int)sizeof(T);            // channels per chunk: one 128-byte operand row
    constexpr int NCQ = KC / VEC;                       // 8 sixteen-byte chunks per operand row
    constexpr int NTQ = kTok / VEC;                     // token groups per tile
    constexpr int kBlocks = NCQ * NTQ / kTok;           // (VEC x VEC) blocks per thread and chunk: 2 (fp32) / 1 (16-bit)
    constexpr int kParts = kPrecise ? 2 : 1;            // hi (+ lo) copies of each operand tile
    constexpr bool kTf32 = OpTraits<T>::kTf32;
    static_assert(!kPrecise || kTf32, "the 3xTF32 split is for fp32 operands");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar_free[kStages];              // "the tensor core has finished reading this stage"
    __shared__ uint32_t tmem_slot;
    const uint32_t a_bytes = kTok * 128;                                  // one A tile
    const uint32_t b_bytes = (uint32_t)(a.n_out / 8) * 1024;              // one B tile (n_out rows of 128 bytes)
    const uint32_t stage_bytes = kParts * (a_bytes + b_bytes);
    // 1024-byte alignment of every tile (SWIZZLE_128B): round the dynamic base up
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int l0 = blockIdx.x * kTok;
    const int L = a.seqlen;
    const uint32_t tmem_cols = a.n_out <= 32 ? 32u : a.n_out <= 64 ? 64u : a.n_out <= 128 ? 128u : 256u;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) umma::mbar_init(&bar_free[s], 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::idesc(OpTraits<T>::kFmt, kTok, a.n_out);

    const T *xb = reinterpret_cast<const T *>(a.x) + (int64_t)b * a.x_bs;
    T *ub = reinterpret_cast<T *>(a.u) + (int64_t)b * a.u_bs;
    const T *xw = reinterpret_cast<const T *>(a.xw);
    const int n_chunks = a.dim / KC;

    for (int c = 0; c < n_chunks; ++c) {
        const int s = c % kStages;
        unsigned char *A_hi = smem + s * stage_bytes;
        unsigned char *A_lo = A_hi + a_bytes;                             // only with kPrecise
        unsigned char *B_hi = A_hi + kParts * a_bytes;
        unsigned char *B_lo = B_hi + b_bytes;
        if (c >= kStages) umma::mbar_wait(&bar_free[s], ((c / kStages) - 1) & 1);     // MMAs of chunk c - kStages are done
        const int k0 = c * KC;

        // ---- B tile: x_proj_weight[:, k0 : k0 + KC], rows of 128 bytes, chunk (e, q) -> swizzled slot
        for (int id = tid; id < a.n_out * NCQ; id += kTok) {
            const int e = id / NCQ, q = id % NCQ;
            const uint4 v = *reinterpret_cast<const uint4 *>(xw + (int64_t)e * a.xw_rs + k0 + q * VEC);
            const uint32_t off = umma::sw128_off(e, q);
            *reinterpret_cast<uint4 *>(B_hi + off) = v;
            if (kPrecise) {
                const float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
                *reinterpret_cast<float4 *>(B_lo + off) =
                    make_float4(umma::tf32_lo(f[0]), umma::tf32_lo(f[1]), umma::tf32_lo(f[2]), umma::tf32_lo(f[3]));
            }
        }

        // ---- A tile: conv + SiLU of (VEC channels x VEC tokens) blocks
#pragma unroll
        for (int it = 0; it < kBlocks; ++it) {
            const int blk = tid + it * kTok;
            const int cq = blk % NCQ;                    // == lane % 8: the 8 lanes of a quarter warp own the 8 chunks of a row
            const int tq = blk / NCQ;
            const int tok0 = l0 + tq * VEC;              // first token of the block (global)
            const int ch0 = k0 + cq * VEC;
            float o[VEC][VEC];                           // [channel][token]
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int ch = ch0 + i;
                float w[kMaxW];
#pragma unroll
                for (int k = 0; k < kMaxW; ++k) {        // right-aligned taps: w[3] multiplies x[l]
                    const int wi = k - (kMaxW - a.width);
                    w[k] = wi >= 0 ? ld_w(a.cw, a.w_dtype, (int64_t)ch * a.cw_ds + wi * a.cw_ws) : 0.f;
                }
                const float bias = a.cb != nullptr ? ld_w(a.cb, a.w_dtype, ch) : 0.f;
                const T *xr = xb + (int64_t)ch * a.x_ds;
                float xv[VEC + VEC];                     // previous vector (halo in its last 3 slots) + own vector
#pragma unroll
                for (int j = 0; j < 2 * VEC; ++j) xv[j] = 0.f;
                if (tok0 < L) {
                    Io<T>::ldv(xr + tok0, reinterpret_cast<float(&)[VEC]>(xv[VEC]));
                    if (tok0 > 0) Io<T>::ldv(xr + tok0 - VEC, reinterpret_cast<float(&)[VEC]>(xv[0]));
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float acc = bias;
#pragma unroll
                    for (int k = 0; k < kMaxW; ++k) acc = fmaf(w[k], xv[VEC + j - (kMaxW - 1) + k], acc);
                    o[i][j] = silu_t<sizeof(T) == 2>(acc);
                }
                if (tok0 < L) Io<T>::stv(ub + (int64_t)ch * a.u_ds + tok0, o[i]);       // u side store, 16 bytes
            }
            // operand rows: token t = tq VEC + j holds channels ch0 .. ch0 + VEC - 1 in chunk cq
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const uint32_t off = umma::sw128_off(tq * VEC + j, cq);
                if constexpr (sizeof(T) == 4) {
                    *reinterpret_cast<float4 *>(A_hi + off) = make_float4(o[0][j], o[1][j], o[2][j], o[3][j]);
                    if (kPrecise)
                        *reinterpret_cast<float4 *>(A_lo + off) =
                            make_float4(umma::tf32_lo(o[0][j]), umma::tf32_lo(o[1][j]), umma::tf32_lo(o[2][j]), umma::tf32_lo(o[3][j]));
                } else {
                    // the GEMM of the reference sees u rounded to the storage type: round here the same way
                    *reinterpret_cast<uint4 *>(A_hi + off) = make_uint4(pack2<T>(o[0][j], o[1][j]), pack2<T>(o[2][j], o[3][j]),
                                                                       pack2<T>(o[4][j], o[5][j]), pack2<T>(o[6][j], o[7][j]));
                }
            }
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t sa_hi = umma::smem_u32(A_hi), sa_lo = umma::smem_u32(A_lo);
            const uint32_t sb_hi = umma::smem_u32(B_hi), sb_lo = umma::smem_u32(B_lo);
#pragma unroll
            for (int k = 0; k < 4; ++k) {                // 4 K steps of 32 bytes per chunk
                const uint32_t ko = k * 32;
                umma::mma<kTf32>(tmem, umma::desc_sw128(sa_hi + ko), umma::desc_sw128(sb_hi + ko), idesc, (c | k) != 0);
                if (kPrecise) {
                    umma::mma<kTf32>(tmem, umma::desc_sw128(sa_lo + ko), umma::desc_sw128(sb_hi + ko), idesc, 1u);
                    umma::mma<kTf32>(tmem, umma::desc_sw128(sa_hi + ko), umma::desc_sw128(sb_lo + ko), idesc, 1u);
                }
            }
            umma::commit(&bar_free[s]);
        }
    }
    // ---- epilogue: all MMAs done -> accumulator row (token) per thread -> x_dbl[b, e, l0 + tid]
    {
        const int last = n_chunks - 1;
        umma::mbar_wait(&bar_free[last % kStages], (last / kStages) & 1);
        umma::fence_after_sync();
        T *ob = reinterpret_cast<T *>(a.xdbl) + (int64_t)b * a.o_bs;
        T *tb = a.tail != nullptr ? reinterpret_cast<T *>(a.tail) + (int64_t)b * a.t_bs : nullptr;
        const int tok = l0 + tid;
        for (int e0 = 0; e0 < a.n_out; e0 += 8) {
            uint32_t v[8];
            umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + e0, v);
            umma::tmem_ld_wait();
            if (tok < L) {
                T *dst = (tb != nullptr && e0 >= a.split) ? tb + (int64_t)(e0 - a.split) * a.t_rs : ob + (int64_t)e0 * a.o_rs;
                const int64_t rs = (tb != nullptr && e0 >= a.split) ? a.t_rs : a.o_rs;
#pragma unroll
                for (int j = 0; j < 8; ++j) Io<T>::st(dst + j * rs + tok, __uint_as_float(v[j]));
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
    (void)lane;
}

template <typename T, bool kPrecise>
int launch(const ConvXprojArgs &a, int batch, cudaStream_t stream) {
    auto kern = conv_xproj_kernel<T, kPrecise>;
    const int parts = kPrecise ? 2 : 1;
    const int smem = kStages * parts * (kTok * 128 + (a.n_out / 8) * 1024) + 1024;     // + alignment slack
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dim3 grid((a.seqlen + kTok - 1) / kTok, batch);
    kern<<<grid, kTok, smem, stream>>>(a);
    return check_launch("conv_xproj_fwd");
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_conv_xproj_fwd(const dimsum_conv_xproj_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "conv_xproj_fwd: null params");
    if (p->batch == 0) return DIMSUM_OK;
    DIMSUM_REQUIRE(p->batch > 0 && p->dim > 0 && p->seqlen > 0, DIMSUM_ERR_INVALID, "conv_xproj_fwd: bad sizes");
    DIMSUM_REQUIRE(p->width >= 2 && p->width <= 4, DIMSUM_ERR_INVALID, "causal_conv1d only supports width between 2 and 4");
    DIMSUM_REQUIRE(p->x && p->conv_weight && p->x_proj_weight && p->u && p->x_dbl, DIMSUM_ERR_INVALID, "conv_xproj_fwd: null pointer");
    DIMSUM_REQUIRE(p->io_dtype >= 0 && p->io_dtype <= 2 && p->w_dtype >= 0 && p->w_dtype <= 2, DIMSUM_ERR_INVALID,
                   "conv_xproj_fwd: unknown dtype");
    DIMSUM_REQUIRE(p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: batch > 65535");
    const int es = p->io_dtype == DIMSUM_F32 ? 4 : 2;
    const int vec = 16 / es, kc = 128 / es;
    DIMSUM_REQUIRE(p->n_out >= 8 && p->n_out <= 256 && p->n_out % 8 == 0, DIMSUM_ERR_UNSUPPORTED,
                   "conv_xproj_fwd: x_proj rows (dt_rank + 2 dstate = %lld) must be a multiple of 8, at most 256", (long long)p->n_out);
    DIMSUM_REQUIRE(p->dim % kc == 0, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: dim must be a multiple of %d", kc);
    DIMSUM_REQUIRE(p->seqlen % vec == 0, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: seqlen must be a multiple of %d", vec);
    auto ok = [&](const void *ptr, int64_t s0, int64_t s1) { return aligned16(ptr) && s0 % vec == 0 && s1 % vec == 0; };
    DIMSUM_REQUIRE(ok(p->x, p->x_batch_stride, p->x_d_stride) && ok(p->u, p->u_batch_stride, p->u_d_stride) &&
                       ok(p->x_proj_weight, p->xw_row_stride, 0),
                   DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: x, u and x_proj_weight need 16-byte aligned rows");
    DIMSUM_REQUIRE(p->x_dbl_tail == nullptr || (p->split_rows > 0 && p->split_rows < p->n_out && p->split_rows % 8 == 0),
                   DIMSUM_ERR_INVALID, "conv_xproj_fwd: split_rows must be a multiple of 8 inside (0, n_out)");
    DIMSUM_REQUIRE(p->precision == 0 || p->io_dtype == DIMSUM_F32, DIMSUM_ERR_INVALID,
                   "conv_xproj_fwd: the 3xTF32 precision applies to fp32 I/O only");

    ConvXprojArgs a;
    a.x = p->x; a.cw = p->conv_weight; a.cb = p->conv_bias; a.xw = p->x_proj_weight; a.u = p->u; a.xdbl = p->x_dbl;
    a.tail = p->x_dbl_tail; a.t_bs = p->tail_batch_stride; a.t_rs = p->tail_row_stride; a.split = (int)p->split_rows;
    a.x_bs = p->x_batch_stride; a.x_ds = p->x_d_stride; a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride;
    a.o_bs = p->x_dbl_batch_stride; a.o_rs = p->x_dbl_row_stride; a.cw_ds = p->w_d_stride; a.cw_ws = p->w_width_stride;
    a.xw_rs = p->xw_row_stride;
    a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.width = (int)p->width; a.n_out = (int)p->n_out;
    a.w_dtype = (int)p->w_dtype; a.precise = (int)p->precision;
    switch (p->io_dtype) {
        case DIMSUM_F32:
            return p->precision ? launch<float, true>(a, (int)p->batch, stream) : launch<float, false>(a, (int)p->batch, stream);
        case DIMSUM_BF16: return launch<__nv_bfloat16, false>(a, (int)p->batch, stream);
        default: return launch<__half, false>(a, (int)p->batch, stream);
    }
}
