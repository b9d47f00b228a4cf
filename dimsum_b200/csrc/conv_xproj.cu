// Causal conv1d + SiLU fused IN FRONT of the x_proj contraction, on TMA + tcgen05 tensor cores (sm_100a).
//
// Replaces, for the composite the model calls (MambaInnerFn*.forward, mamba/mamba_ssm/ops/selective_scan_interface.py:836-866):
//     conv1d_out = causal_conv1d_fwd(x, w, b, silu)                       (a full read + write of (B, D, L))
//     x_dbl      = F.linear(rearrange(conv1d_out, "b d l -> (b l) d"), x_proj_weight)     (a second full read of conv1d_out)
//     B, C       = rearrange(x_dbl[:, r:r+N], "(b l) n -> b 1 n l").contiguous(), ...    (two more small passes)
// with ONE kernel that reads x once, writes u = conv1d_out once (the scan and the backward need it), and produces
// x_dbl ALREADY channel-major: dt as the (rank, B L) operand of the dt_proj GEMM, B and C in the layout the scan reads.
//
// Design (one CTA per SM-resident slot = one batch row x 128 tokens, warp-specialised):
//   * warp 4, one lane: TMA producer.  Per chunk of K (128 bytes of channels: 32 fp32 / 64 16-bit) it issues
//     cp.async.bulk.tensor loads of the x tile (KC channel rows x (128 + halo) tokens; the 3-D tensor map zero-fills the
//     causal left edge and the ragged right edge), of the matching x_proj_weight tile (written 128-byte-swizzled, the
//     tensor core's operand layout) and plain bulk copies of the conv taps / bias, all completing on one mbarrier per
//     ring slot.  The ring is 3-4 chunks deep, so HBM latency never reaches the compute warps.
//   * warps 0-7: two compute groups of 128 threads that take alternate chunks (8 warps keep the SM's schedulers busy; one
//     CTA per SM keeps the grid at 6.9 waves of 148).  Every thread convolves a (VEC channels x VEC tokens) block read from
//     the ring with conflict-free 16-byte loads -- the register-level transpose from channel-major x rows to the token-major
//     (K-major) operand rows -- stores its u rows to HBM (8 lanes = 128 contiguous bytes) and its operand rows to the
//     group's padded no-swizzle K-major tile (SBO 144 bytes: conflict-free 16-byte stores).  After a 128-thread named
//     barrier one thread of the group issues tcgen05.mma (M = 128 tokens, N = dt_rank + 2 N, K = 32 bytes per instruction)
//     into the GROUP'S OWN TMEM accumulator (two issuing threads are not ordered, so they must not share one) and commits to
//     the mbarriers that free the group's operand tile and the ring slot; the tensor core runs while the next chunks are
//     produced.  The epilogue adds the two accumulators.
//   * fp32 I/O: kind::tf32, one pass (TF32, what cuBLAS does under allow_tf32) or the 3xTF32 split (fp32-grade, ~1e-6; the
//     low part of x_proj_weight arrives through a second tensor map); 16-bit I/O: kind::f16 on the bf16 / fp16 values the
//     reference's GEMM would see.  Epilogue: tcgen05.ld (thread = token) -> coalesced rows of dt / B / C.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace dimsum {
namespace {

constexpr int kTok = 128;          // tokens per CTA == MMA M == compute threads per CTA
constexpr int kGroups = 2;         // compute warp groups of 128 threads; group g owns chunks g, g + 2, ... and its own accumulator
constexpr int kThreads = kGroups * kTok + 32;   // + the TMA producer warp
constexpr int kMaxW = 4;
constexpr int kSboA = 144;         // bytes between 8-token groups of the A operand tile (128 + 16: conflict-free stores)
constexpr int kLboA = (kTok / 8) * kSboA;   // bytes between 16-byte K chunks
constexpr int kABytes = 8 * kLboA;          // one A tile: 8 chunks x 16 token groups x 144 = 18432

struct ConvXprojArgs {
    const void *cw, *cb;
    void *u, *xdbl, *tail;
    int64_t u_bs, u_ds, o_bs, o_rs, t_bs, t_rs;
    int dim, seqlen, width, n_out, split;
};

template <typename T> struct OpTraits;
template <> struct OpTraits<float> { static constexpr bool kTf32 = true; static constexpr int kFmt = umma::kFmtTF32; };
template <> struct OpTraits<__nv_bfloat16> { static constexpr bool kTf32 = false; static constexpr int kFmt = umma::kFmtBF16; };
template <> struct OpTraits<__half> { static constexpr bool kTf32 = false; static constexpr int kFmt = umma::kFmtF16; };

template <typename T> DEV uint32_t pack2(float a, float b);
template <> DEV uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return reinterpret_cast<uint32_t &>(h);
}
template <> DEV uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return reinterpret_cast<uint32_t &>(h);
}

// no-swizzle K-major matrix descriptor (validated with padded strides by tools/microbench/umma_probe.cu)
DEV uint64_t desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(umma::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
DEV void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(umma::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
DEV void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(umma::smem_u32(dst)), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}
DEV void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ring-slot layout (bytes): x tile | x_proj_weight tile (hi) | (lo) | conv taps | conv bias   -- every part 128-byte aligned,
// the weight tiles 1024-byte aligned (SWIZZLE_128B)
template <typename T, bool kPrecise>
struct Slot {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int KC = 128 / (int)sizeof(T);
    static constexpr int kPitch = (kTok + VEC) * (int)sizeof(T);          // bytes per x row: halo vector + 128 tokens
    static constexpr int kXBytes = KC * kPitch;                            // 16896 (fp32) / 17408 (16-bit)
    static constexpr int kXPad = (kXBytes + 1023) / 1024 * 1024;
    __host__ __device__ static int w_bytes(int n_out) { return (n_out / 8) * 1024; }
    static constexpr int kCwBytes = (KC * 20 + 1023) / 1024 * 1024;       // conv taps (KC x 16 bytes) + bias (KC x 4 bytes)
    __host__ __device__ static int bytes(int n_out) { return kXPad + (kPrecise ? 2 : 1) * w_bytes(n_out) + kCwBytes; }
};

template <typename T, bool kPrecise, int kNS>
__global__ void __launch_bounds__(kThreads, 1) conv_xproj_kernel(const ConvXprojArgs a, const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_w,
                                                                 const __grid_constant__ CUtensorMap map_wlo) {
    using S = Slot<T, kPrecise>;
    constexpr int VEC = S::VEC, KC = S::KC;
    constexpr int NCQ = KC / VEC;                       // 8 sixteen-byte chunks per operand row
    constexpr int kBlocks = NCQ * (kTok / VEC) / kTok;  // (VEC x VEC) blocks per thread and chunk: 2 (fp32) / 1 (16-bit)
    constexpr int kParts = kPrecise ? 2 : 1;
    constexpr bool kTf32 = OpTraits<T>::kTf32;
    static_assert(!kPrecise || kTf32, "the 3xTF32 split is for fp32 operands");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar_full[kNS], bar_empty[kNS], bar_afree[kGroups];
    __shared__ uint32_t tmem_slot;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int w_bytes = S::w_bytes(a.n_out);
    const int slot_bytes = S::bytes(a.n_out);
    unsigned char *a_tiles = smem + kNS * slot_bytes;   // [kGroups][kParts][kABytes]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int l0 = blockIdx.x * kTok;
    const int L = a.seqlen;
    const int n_chunks = a.dim / KC;
    const uint32_t acc_cols = (uint32_t)a.n_out;                         // columns per group accumulator
    const uint32_t need_cols = kGroups * acc_cols;
    const uint32_t tmem_cols = need_cols <= 32 ? 32u : need_cols <= 64 ? 64u : need_cols <= 128 ? 128u : need_cols <= 256 ? 256u : 512u;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kNS; ++s) { umma::mbar_init(&bar_full[s], 1); umma::mbar_init(&bar_empty[s], 1); }
#pragma unroll
        for (int g = 0; g < kGroups; ++g) umma::mbar_init(&bar_afree[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == kGroups * 4) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(S::kXBytes + kParts * w_bytes + KC * 16 + KC * 4);
            const float *cw = reinterpret_cast<const float *>(a.cw);
            const float *cb = reinterpret_cast<const float *>(a.cb);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % kNS;
                if (c >= kNS) umma::mbar_wait(&bar_empty[s], ((c / kNS) - 1) & 1);      // the MMAs of chunk c - kNS are done
                unsigned char *slot = smem + s * slot_bytes;
                const int k0 = c * KC;
                mbar_expect_tx(&bar_full[s], tx);
                tma_load_3d(slot, &map_x, &bar_full[s], l0 - VEC, k0, b);               // tokens l0 - VEC .. l0 + 127 (zero-filled edges)
                tma_load_2d(slot + S::kXPad, &map_w, &bar_full[s], k0, 0);
                if (kPrecise) tma_load_2d(slot + S::kXPad + w_bytes, &map_wlo, &bar_full[s], k0, 0);
                unsigned char *cws = slot + S::kXPad + kParts * w_bytes;
                bulk_load(cws, cw + (int64_t)k0 * 4, KC * 16, &bar_full[s]);
                bulk_load(cws + KC * 16, cb + k0, KC * 4, &bar_full[s]);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------ conv + operand tiles + MMA issue
        const int grp = warp >> 2, gwarp = warp & 3, gtid = tid & (kTok - 1);
        const uint32_t idesc = umma::idesc(OpTraits<T>::kFmt, kTok, a.n_out);
        const uint32_t acc = tmem + grp * acc_cols;
        T *ub = reinterpret_cast<T *>(a.u) + (int64_t)b * a.u_bs;
        unsigned char *A_hi = a_tiles + grp * kParts * kABytes;
        unsigned char *A_lo = A_hi + kABytes;
        int mine = 0;                                    // chunks this group has issued so far
        for (int c = grp; c < n_chunks; c += kGroups, ++mine) {
            const int s = c % kNS;
            unsigned char *slot = smem + s * slot_bytes;
            const float *cws = reinterpret_cast<const float *>(slot + S::kXPad + kParts * w_bytes);
            const float *cbs = cws + KC * 4;
            const int k0 = c * KC;
            umma::mbar_wait(&bar_full[s], (c / kNS) & 1);                               // the ring slot has landed
            if (mine >= 1) umma::mbar_wait(&bar_afree[grp], (mine - 1) & 1);            // the group's previous MMAs have read its A tile
#pragma unroll
            for (int it = 0; it < kBlocks; ++it) {
                // lane % 8 = token group (8 lanes: 128 contiguous bytes of a channel row), lane / 8 (+ 4 it) = chunk of K
                const int tq = sizeof(T) == 4 ? (lane & 7) + 8 * gwarp : (lane & 7) + 8 * (gwarp & 1);
                const int cq = sizeof(T) == 4 ? (lane >> 3) + 4 * it : (lane >> 3) + 4 * (gwarp >> 1);
                const int tok0 = l0 + tq * VEC;
                float o[VEC][VEC];                       // [channel][token]
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const int r = cq * VEC + i;          // channel row inside the chunk
                    const float4 w4 = *reinterpret_cast<const float4 *>(cws + r * 4);
                    const float w[kMaxW] = {w4.x, w4.y, w4.z, w4.w};          // width 4: w[3] multiplies x[l]
                    const float bias = cbs[r];
                    const T *xr = reinterpret_cast<const T *>(slot + r * S::kPitch) + tq * VEC;
                    float xv[2 * VEC];                   // previous vector (halo in its last 3 slots) + own vector
                    Io<T>::ldv(xr, reinterpret_cast<float(&)[VEC]>(xv[0]));
                    Io<T>::ldv(xr + VEC, reinterpret_cast<float(&)[VEC]>(xv[VEC]));
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        float acc_ = bias;
#pragma unroll
                        for (int k = 0; k < kMaxW; ++k) acc_ = fmaf(w[k], xv[VEC + j - (kMaxW - 1) + k], acc_);
                        o[i][j] = silu_t<sizeof(T) == 2>(acc_);
                    }
                    if (tok0 < L) Io<T>::stv(ub + (int64_t)(k0 + r) * a.u_ds + tok0, o[i]);       // u side store, 16 bytes
                }
                // operand rows: token t = tq VEC + j holds channels cq VEC .. + VEC - 1 in chunk cq
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const int t = tq * VEC + j;
                    const uint32_t off = (uint32_t)((t & 7) * 16 + (t >> 3) * kSboA + cq * kLboA);
                    if constexpr (sizeof(T) == 4) {
                        *reinterpret_cast<float4 *>(A_hi + off) = make_float4(o[0][j], o[1][j], o[2][j], o[3][j]);
                        if (kPrecise)
                            *reinterpret_cast<float4 *>(A_lo + off) = make_float4(umma::tf32_lo(o[0][j]), umma::tf32_lo(o[1][j]),
                                                                                  umma::tf32_lo(o[2][j]), umma::tf32_lo(o[3][j]));
                    } else {
                        // the GEMM of the reference sees u rounded to the storage type: round here the same way
                        *reinterpret_cast<uint4 *>(A_hi + off) = make_uint4(pack2<T>(o[0][j], o[1][j]), pack2<T>(o[2][j], o[3][j]),
                                                                           pack2<T>(o[4][j], o[5][j]), pack2<T>(o[6][j], o[7][j]));
                    }
                }
            }
            umma::fence_smem_to_async();
            named_bar_sync(1 + grp, kTok);
            if (gtid == 0) {
                umma::fence_after_sync();
                const uint32_t sa_hi = umma::smem_u32(A_hi), sa_lo = umma::smem_u32(A_lo);
                const uint32_t sb_hi = umma::smem_u32(slot + S::kXPad), sb_lo = sb_hi + (uint32_t)w_bytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {            // 4 K steps of 32 bytes (two 16-byte chunks) per chunk
                    const uint64_t da_hi = desc_plain(sa_hi + k * 2 * kLboA, kLboA, kSboA);
                    const uint64_t db_hi = umma::desc_sw128(sb_hi + k * 32);
                    umma::mma<kTf32>(acc, da_hi, db_hi, idesc, (mine | k) != 0);
                    if (kPrecise) {
                        umma::mma<kTf32>(acc, desc_plain(sa_lo + k * 2 * kLboA, kLboA, kSboA), db_hi, idesc, 1u);
                        umma::mma<kTf32>(acc, da_hi, umma::desc_sw128(sb_lo + k * 32), idesc, 1u);
                    }
                }
                umma::commit(&bar_afree[grp]);           // the group's A tile may be overwritten
                umma::commit(&bar_empty[s]);             // the ring slot may be refilled (chunk c + kNS)
            }
        }
        // ---- epilogue (group 0): both groups' MMAs done -> sum of the two accumulator rows (thread = token) -> coalesced rows
        if (grp == 0) {
            int ng[kGroups];                                             // chunks issued by each group
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                ng[g] = (n_chunks - g + kGroups - 1) / kGroups;
                if (ng[g] > 0) umma::mbar_wait(&bar_afree[g], (ng[g] - 1) & 1);
            }
            umma::fence_after_sync();
            T *ob = reinterpret_cast<T *>(a.xdbl) + (int64_t)b * a.o_bs;
            T *tb = a.tail != nullptr ? reinterpret_cast<T *>(a.tail) + (int64_t)b * a.t_bs : nullptr;
            const int tok = l0 + tid;
            for (int e0 = 0; e0 < a.n_out; e0 += 8) {
                uint32_t v[kGroups][8];
#pragma unroll
                for (int g = 0; g < kGroups; ++g)
                    if (ng[g] > 0) umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + g * acc_cols + e0, v[g]);
                umma::tmem_ld_wait();
                if (tok < L) {
                    const bool in_tail = tb != nullptr && e0 >= a.split;
                    T *dst = in_tail ? tb + (int64_t)(e0 - a.split) * a.t_rs : ob + (int64_t)e0 * a.o_rs;
                    const int64_t rs = in_tail ? a.t_rs : a.o_rs;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float sum = __uint_as_float(v[0][j]);
#pragma unroll
                        for (int g = 1; g < kGroups; ++g) sum += ng[g] > 0 ? __uint_as_float(v[g][j]) : 0.f;
                        Io<T>::st(dst + j * rs + tok, sum);
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

CUtensorMapDataType tm_dtype(int io_dtype) {
    return io_dtype == DIMSUM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : io_dtype == DIMSUM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

template <typename T, bool kPrecise, int kNS>
int launch(const ConvXprojArgs &a, int batch, const CUtensorMap &mx, const CUtensorMap &mw, const CUtensorMap &mwl, cudaStream_t stream) {
    auto kern = conv_xproj_kernel<T, kPrecise, kNS>;
    const int smem = kNS * Slot<T, kPrecise>::bytes(a.n_out) + kGroups * (kPrecise ? 2 : 1) * kABytes + 1024;
    if (smem > 227 * 1024) return fail(DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: n_out = %d needs %d bytes of shared memory", a.n_out, smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dim3 grid((a.seqlen + kTok - 1) / kTok, batch);
    kern<<<grid, kThreads, smem, stream>>>(a, mx, mw, mwl);
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
    DIMSUM_REQUIRE(p->width == 4, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: only the conv width 4 of the model is fused");
    DIMSUM_REQUIRE(p->x && p->conv_weight && p->conv_bias && p->x_proj_weight && p->u && p->x_dbl, DIMSUM_ERR_INVALID,
                   "conv_xproj_fwd: null pointer (the fused kernel needs a conv bias)");
    DIMSUM_REQUIRE(p->io_dtype >= 0 && p->io_dtype <= 2, DIMSUM_ERR_INVALID, "conv_xproj_fwd: unknown dtype");
    DIMSUM_REQUIRE(p->w_dtype == DIMSUM_F32 && p->w_width_stride == 1 && p->w_d_stride == p->width &&
                       aligned16(p->conv_weight) && aligned16(p->conv_bias),
                   DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: conv weight must be contiguous fp32 (dim, 4), 16-byte aligned");
    DIMSUM_REQUIRE(p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: batch > 65535");
    const int es = p->io_dtype == DIMSUM_F32 ? 4 : 2;
    const int vec = 16 / es, kc = 128 / es;
    DIMSUM_REQUIRE(p->n_out >= 8 && p->n_out <= 128 && p->n_out % 8 == 0 && kGroups * p->n_out <= 512, DIMSUM_ERR_UNSUPPORTED,
                   "conv_xproj_fwd: x_proj rows (dt_rank + 2 dstate = %lld) must be a multiple of 8, at most 128", (long long)p->n_out);
    DIMSUM_REQUIRE(p->dim % kc == 0, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: dim must be a multiple of %d", kc);
    DIMSUM_REQUIRE(p->seqlen % vec == 0, DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: seqlen must be a multiple of %d", vec);
    auto ok = [&](const void *ptr, int64_t s0, int64_t s1) { return aligned16(ptr) && s0 % vec == 0 && s1 % vec == 0; };
    DIMSUM_REQUIRE(ok(p->x, p->x_batch_stride, p->x_d_stride) && ok(p->u, p->u_batch_stride, p->u_d_stride) &&
                       ok(p->x_proj_weight, p->xw_row_stride, 0),
                   DIMSUM_ERR_UNSUPPORTED, "conv_xproj_fwd: x, u and x_proj_weight need 16-byte aligned rows");
    DIMSUM_REQUIRE(p->x_dbl_tail == nullptr || (p->split_rows > 0 && p->split_rows < p->n_out && p->split_rows % 8 == 0),
                   DIMSUM_ERR_INVALID, "conv_xproj_fwd: split_rows must be a multiple of 8 inside (0, n_out)");
    DIMSUM_REQUIRE(p->precision == 0 || (p->io_dtype == DIMSUM_F32 && p->x_proj_weight_lo != nullptr && aligned16(p->x_proj_weight_lo)),
                   DIMSUM_ERR_INVALID, "conv_xproj_fwd: the 3xTF32 precision needs fp32 I/O and x_proj_weight_lo");
    EncodeTiledFn enc = encode_tiled();
    DIMSUM_REQUIRE(enc != nullptr, DIMSUM_ERR_CUDA, "conv_xproj_fwd: cuTensorMapEncodeTiled is not available from this driver");

    // x: (batch, dim, seqlen) -> box (128 + halo tokens, KC channels, 1 row); out-of-range tokens read as zero
    CUtensorMap mx, mw, mwl;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)p->seqlen, (cuuint64_t)p->dim, (cuuint64_t)p->batch};
        const cuuint64_t strides[2] = {(cuuint64_t)p->x_d_stride * es, (cuuint64_t)p->x_batch_stride * es};
        const cuuint32_t box[3] = {(cuuint32_t)(kTok + vec), (cuuint32_t)kc, 1u};
        const cuuint32_t ones[3] = {1u, 1u, 1u};
        const CUresult r = enc(&mx, tm_dtype((int)p->io_dtype), 3, const_cast<void *>(p->x), dims, strides, box, ones,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DIMSUM_REQUIRE(r == CUDA_SUCCESS, DIMSUM_ERR_CUDA, "conv_xproj_fwd: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    auto weight_map = [&](CUtensorMap *m, const void *w) {
        const cuuint64_t dims[2] = {(cuuint64_t)p->dim, (cuuint64_t)p->n_out};
        const cuuint64_t strides[1] = {(cuuint64_t)p->xw_row_stride * es};
        const cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)p->n_out};
        const cuuint32_t ones[2] = {1u, 1u};
        return enc(m, tm_dtype((int)p->io_dtype), 2, const_cast<void *>(w), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    {
        CUresult r = weight_map(&mw, p->x_proj_weight);
        DIMSUM_REQUIRE(r == CUDA_SUCCESS, DIMSUM_ERR_CUDA, "conv_xproj_fwd: cuTensorMapEncodeTiled(x_proj_weight) failed with %d", (int)r);
        r = weight_map(&mwl, p->precision ? p->x_proj_weight_lo : p->x_proj_weight);
        DIMSUM_REQUIRE(r == CUDA_SUCCESS, DIMSUM_ERR_CUDA, "conv_xproj_fwd: cuTensorMapEncodeTiled(x_proj_weight_lo) failed with %d", (int)r);
    }

    ConvXprojArgs a;
    a.cw = p->conv_weight; a.cb = p->conv_bias; a.u = p->u; a.xdbl = p->x_dbl;
    a.tail = p->x_dbl_tail; a.t_bs = p->tail_batch_stride; a.t_rs = p->tail_row_stride; a.split = (int)p->split_rows;
    a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride;
    a.o_bs = p->x_dbl_batch_stride; a.o_rs = p->x_dbl_row_stride;
    a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.width = (int)p->width; a.n_out = (int)p->n_out;
    const int B = (int)p->batch;
    switch (p->io_dtype) {
        case DIMSUM_F32:
            return p->precision ? launch<float, true, 3>(a, B, mx, mw, mwl, stream) : launch<float, false, 4>(a, B, mx, mw, mwl, stream);
        case DIMSUM_BF16: return launch<__nv_bfloat16, false, 4>(a, B, mx, mw, mwl, stream);
        default: return launch<__half, false, 4>(a, B, mx, mw, mwl, stream);
    }
}
