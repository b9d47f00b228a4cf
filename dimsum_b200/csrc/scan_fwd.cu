// Selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303).
// Design (B200-first, not a port):
//   * one CTA = one batch row x a slab of 128 channels; one THREAD owns one channel row and walks the
//     sequence with all <=16 states in registers as 8 packed f32x2 pairs (FFMA2/FMUL2) -- no cross-lane
//     scan, so 4 FP ops + 1 exp per state-step instead of the 6 of a lane-split scan;
//   * the sequence is processed in LC-step chunks.  Per chunk the CTA (a) loads u/delta with coalesced
//     128-bit loads, applies bias+softplus, and parks fp32 delta/u in padded shared memory, together with
//     the B/C tile transposed to [l][n] (read later as warp-wide broadcasts, so B/C are fetched once per
//     128 channels instead of once per channel as in the reference), (b) runs the per-thread recurrence
//     reading its row with conflict-free LDS.128, (c) writes y through the same padded tile and gates it
//     with silu(z) on the coalesced way out;
//   * part of the 16 exps per step can run as a polynomial on the FMA pipe (kPoly pairs), because at
//     16 MUFU lanes/clk/SM the exp unit, not HBM, is the first limiter of this kernel.
#include "common.cuh"

namespace dimsum {
namespace {

constexpr int kRows = 128;   // threads per CTA == channel rows per CTA
constexpr int kNS = 16;      // padded state count
constexpr int kBCPitch = 20; // words; 16 states + pad -> conflict-free transposed STS.128 and aligned LDS.128

struct ScanFwdArgs {
    const void *u, *delta, *z, *B, *C;
    const float *A, *D, *delta_bias;
    void *out, *out_z;
    float *x;
    const int32_t *perm;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, out_bs, out_ds, oz_bs, oz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int dim, seqlen, dstate, n_groups, n_chunks, chunk_len, softplus;
};

template <int LC>
struct ScanSmem {
    float dl[kRows][LC + 4];
    float uy[kRows][LC + 4];
    float Bs[LC][kBCPitch];
    float Cs[LC][kBCPitch];
    float bias[kRows];
};

// x[b][d][chunk][2n + slot]: slot 0 = h after 16 steps of the 32-step chunk, slot 1 = h after the chunk.
DEV void store_state(const ScanFwdArgs &a, int b, int d, int chunk, int slot, const float2 (&h2)[kNS / 2]) {
    float *xp = a.x + (((int64_t)b * a.dim + d) * a.n_chunks + chunk) * (2 * a.dstate) + slot;
#pragma unroll
    for (int p = 0; p < kNS / 2; ++p) {
        if (2 * p < a.dstate) xp[4 * p] = h2[p].x;
        if (2 * p + 1 < a.dstate) xp[4 * p + 2] = h2[p].y;
    }
}

template <typename T, int LC, bool kVecIO, bool kHasZ, int kPoly, int kPolyDeg>
__global__ void __launch_bounds__(kRows, 4) scan_fwd_kernel(const ScanFwdArgs a) {
    constexpr int VEC = Io<T>::kVec;
    constexpr int VPR = LC / VEC;       // 16-byte vectors per row chunk
    constexpr int RPP = kRows / VPR;    // rows covered per pass of the coalesced mapping
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem<LC> &s = *reinterpret_cast<ScanSmem<LC> *>(smem_raw);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int dpg = a.dim / a.n_groups;                 // channels per group
    const int slabs_per_group = (dpg + kRows - 1) / kRows;
    const int g = blockIdx.x / slabs_per_group;
    const int d0 = g * dpg + (blockIdx.x % slabs_per_group) * kRows;
    const int nrows = min(kRows, (g + 1) * dpg - d0);   // valid rows of this slab
    const int L = a.seqlen;

    const T *u = reinterpret_cast<const T *>(a.u) + b * a.u_bs + (int64_t)d0 * a.u_ds;
    const T *dl = reinterpret_cast<const T *>(a.delta) + b * a.dl_bs + (int64_t)d0 * a.dl_ds;
    const T *Bg = reinterpret_cast<const T *>(a.B) + b * a.B_bs + g * a.B_gs;
    const T *Cg = reinterpret_cast<const T *>(a.C) + b * a.C_bs + g * a.C_gs;

    // per-row constants of the scanning thread
    const bool row_ok = tid < nrows;
    float2 A2[kNS / 2];
    float Dv = 0.f;
    {
        const float *Arow = a.A + (int64_t)(d0 + tid) * a.A_ds;
#pragma unroll
        for (int p = 0; p < kNS / 2; ++p) {
            float a0 = (row_ok && 2 * p < a.dstate) ? Arow[(2 * p) * a.A_ns] : 0.f;
            float a1 = (row_ok && 2 * p + 1 < a.dstate) ? Arow[(2 * p + 1) * a.A_ns] : 0.f;
            A2[p] = make_float2(a0 * kLog2e, a1 * kLog2e);
        }
        if (row_ok && a.D != nullptr) Dv = a.D[d0 + tid];
        s.bias[tid] = (row_ok && a.delta_bias != nullptr) ? a.delta_bias[d0 + tid] : 0.f;
    }
    float2 h2[kNS / 2];
#pragma unroll
    for (int p = 0; p < kNS / 2; ++p) h2[p] = make_float2(0.f, 0.f);
    __syncthreads();

    const int n_lc = (L + LC - 1) / LC;
    for (int c = 0; c < n_lc; ++c) {
        const int l0 = c * LC;
        // ---------------------------------------------------------------- (a) prep: global -> smem
        if (kVecIO) {
            const int vcol = tid % VPR, rsub = tid / VPR;
            const bool col_ok = l0 + vcol * VEC < L;
#pragma unroll
            for (int pass = 0; pass < VPR; ++pass) {
                const int r = pass * RPP + rsub;
                float uv[VEC], dv[VEC];
                if (r < nrows && col_ok) {
                    Io<T>::ldv(u + (int64_t)r * a.u_ds + l0 + vcol * VEC, uv);
                    Io<T>::ldv(dl + (int64_t)r * a.dl_ds + l0 + vcol * VEC, dv);
                    const float bias = s.bias[r];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float d = dv[i] + bias;
                        dv[i] = a.softplus ? softplus_f(d) : d;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { uv[i] = 0.f; dv[i] = 0.f; }
                }
#pragma unroll
                for (int i = 0; i < VEC; i += 4) {
                    *reinterpret_cast<float4 *>(&s.dl[r][vcol * VEC + i]) = make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]);
                    *reinterpret_cast<float4 *>(&s.uy[r][vcol * VEC + i]) = make_float4(uv[i], uv[i + 1], uv[i + 2], uv[i + 3]);
                }
            }
        } else {
            for (int idx = tid; idx < kRows * LC; idx += kRows) {
                const int r = idx / LC, col = idx % LC;
                float uv = 0.f, dv = 0.f;
                if (r < nrows && l0 + col < L) {
                    uv = Io<T>::ld(u + (int64_t)r * a.u_ds + l0 + col);
                    float d = Io<T>::ld(dl + (int64_t)r * a.dl_ds + l0 + col) + s.bias[r];
                    dv = a.softplus ? softplus_f(d) : d;
                }
                s.dl[r][col] = dv;
                s.uy[r][col] = uv;
            }
        }
        {   // B / C tile -> [l][n]; thread = (l = tid % LC, state quad = tid / LC) (+ strided for LC < 32)
            for (int idx = tid; idx < LC * (kNS / 4); idx += kRows) {
                const int l = idx % LC, nq = idx / LC;
                float bv[4], cv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = nq * 4 + i;
                    const bool ok = n < a.dstate && l0 + l < L;
                    bv[i] = ok ? Io<T>::ld(Bg + n * a.B_ns + l0 + l) : 0.f;
                    cv[i] = ok ? Io<T>::ld(Cg + n * a.C_ns + l0 + l) : 0.f;
                }
                *reinterpret_cast<float4 *>(&s.Bs[l][nq * 4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
                *reinterpret_cast<float4 *>(&s.Cs[l][nq * 4]) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- (b) recurrence, thread == row
#pragma unroll 1
        for (int j = 0; j < LC; j += 4) {
            const float4 d4 = *reinterpret_cast<const float4 *>(&s.dl[tid][j]);
            const float4 u4 = *reinterpret_cast<const float4 *>(&s.uy[tid][j]);
            const float ds[4] = {d4.x, d4.y, d4.z, d4.w};
            const float us[4] = {u4.x, u4.y, u4.z, u4.w};
            float ys[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float dlt = ds[k];
                const float du = dlt * us[k];
                float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < kNS / 4; ++q) {
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s.Bs[j + k][q * 4]);
                    const float4 Cq = *reinterpret_cast<const float4 *>(&s.Cs[j + k][q * 4]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = q * 2 + e;
                        const float2 Bp = e ? make_float2(Bq.z, Bq.w) : make_float2(Bq.x, Bq.y);
                        const float2 Cp = e ? make_float2(Cq.z, Cq.w) : make_float2(Cq.x, Cq.y);
                        float2 t = mul2(splat2(dlt), A2[p]);
                        float2 dec;
                        if (p < kPoly) {
                            dec = ex2_poly2<kPolyDeg>(t);
                        } else {
                            dec = make_float2(ex2_mufu(t.x), ex2_mufu(t.y));
                        }
                        const float2 drive = mul2(splat2(du), Bp);
                        h2[p] = fma2(dec, h2[p], drive);
                        y2 = fma2(Cp, h2[p], y2);
                    }
                }
                ys[k] = fmaf(Dv, us[k], y2.x + y2.y);
            }
            *reinterpret_cast<float4 *>(&s.uy[tid][j]) = make_float4(ys[0], ys[1], ys[2], ys[3]);
            if (LC == 32 && j == 12 && a.x != nullptr && row_ok) store_state(a, b, d0 + tid, c, 0, h2);
        }
        // checkpoint for the backward / last_state: h after the chunk (slot 1; slot 0 was written mid-chunk)
        if (a.x != nullptr && row_ok) store_state(a, b, d0 + tid, c, 1, h2);
        __syncthreads();

        // ---------------------------------------------------------------- (c) epilogue: smem -> global
        T *out = reinterpret_cast<T *>(a.out);
        T *oz = reinterpret_cast<T *>(a.out_z);
        const T *z = reinterpret_cast<const T *>(a.z);
        if (kVecIO && a.perm == nullptr) {
            const int vcol = tid % VPR, rsub = tid / VPR;
            if (l0 + vcol * VEC < L) {
#pragma unroll
                for (int pass = 0; pass < VPR; ++pass) {
                    const int r = pass * RPP + rsub;
                    if (r >= nrows) continue;
                    float yv[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; i += 4) {
                        const float4 y4 = *reinterpret_cast<const float4 *>(&s.uy[r][vcol * VEC + i]);
                        yv[i] = y4.x; yv[i + 1] = y4.y; yv[i + 2] = y4.z; yv[i + 3] = y4.w;
                    }
                    const int64_t col = l0 + vcol * VEC;
                    if (out != nullptr) Io<T>::stv(out + b * a.out_bs + (int64_t)(d0 + r) * a.out_ds + col, yv);
                    if (kHasZ) {
                        float zv[VEC];
                        Io<T>::ldv(z + b * a.z_bs + (int64_t)(d0 + r) * a.z_ds + col, zv);
#pragma unroll
                        for (int i = 0; i < VEC; ++i) yv[i] *= silu_f(zv[i]);
                        Io<T>::stv(oz + b * a.oz_bs + (int64_t)(d0 + r) * a.oz_ds + col, yv);
                    }
                }
            }
        } else {
            for (int idx = tid; idx < kRows * LC; idx += kRows) {
                const int r = idx / LC, col = idx % LC;
                if (r >= nrows || l0 + col >= L) continue;
                float y = s.uy[r][col];
                if (out != nullptr) Io<T>::st(out + b * a.out_bs + (int64_t)(d0 + r) * a.out_ds + l0 + col, y);
                if (kHasZ) {
                    const int tok = a.perm != nullptr ? a.perm[l0 + col] : l0 + col;
                    y *= silu_f(Io<T>::ld(z + b * a.z_bs + (int64_t)(d0 + r) * a.z_ds + tok));
                    Io<T>::st(oz + b * a.oz_bs + (int64_t)(d0 + r) * a.oz_ds + tok, y);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int LC, bool kVecIO, bool kHasZ, int kPoly, int kPolyDeg>
int launch(const ScanFwdArgs &a, int batch, cudaStream_t stream) {
    auto kern = scan_fwd_kernel<T, LC, kVecIO, kHasZ, kPoly, kPolyDeg>;
    const int smem = (int)sizeof(ScanSmem<LC>);
    static bool configured = false;   // per instantiation
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    const int dpg = a.dim / a.n_groups;
    dim3 grid(a.n_groups * ((dpg + kRows - 1) / kRows), batch);
    kern<<<grid, kRows, smem, stream>>>(a);
    return check_launch("selective_scan_fwd");
}

template <typename T, int kPoly, int kPolyDeg>
int dispatch(const ScanFwdArgs &a, int batch, bool vec_ok, cudaStream_t stream) {
    const bool has_z = a.z != nullptr;
    if (vec_ok) {
        return has_z ? launch<T, 32, true, true, kPoly, kPolyDeg>(a, batch, stream)
                     : launch<T, 32, true, false, kPoly, kPolyDeg>(a, batch, stream);
    }
    return has_z ? launch<T, 32, false, true, kPoly, kPolyDeg>(a, batch, stream)
                 : launch<T, 32, false, false, kPoly, kPolyDeg>(a, batch, stream);
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_selective_scan_fwd(const dimsum_scan_fwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "selective_scan_fwd: null params");
    DIMSUM_REQUIRE(p->batch >= 0 && p->dim > 0 && p->seqlen > 0 && p->dstate > 0, DIMSUM_ERR_INVALID,
                   "selective_scan_fwd: bad sizes batch=%lld dim=%lld seqlen=%lld dstate=%lld", (long long)p->batch,
                   (long long)p->dim, (long long)p->seqlen, (long long)p->dstate);
    DIMSUM_REQUIRE(p->dstate <= 256, DIMSUM_ERR_INVALID, "selective_scan only supports state dimension <= 256");
    DIMSUM_REQUIRE(p->dstate <= kNS, DIMSUM_ERR_UNSUPPORTED,
                   "selective_scan_fwd: dstate=%lld > 16 is not implemented in the B200 kernels", (long long)p->dstate);
    DIMSUM_REQUIRE(p->n_groups >= 1 && p->dim % p->n_groups == 0, DIMSUM_ERR_INVALID,
                   "selective_scan_fwd: dim %lld not divisible by n_groups %lld", (long long)p->dim, (long long)p->n_groups);
    DIMSUM_REQUIRE(p->u && p->delta && p->A && p->B && p->C, DIMSUM_ERR_INVALID, "selective_scan_fwd: null input pointer");
    DIMSUM_REQUIRE((p->z == nullptr) == (p->out_z == nullptr), DIMSUM_ERR_INVALID,
                   "selective_scan_fwd: z and out_z must be given together");
    DIMSUM_REQUIRE(p->out != nullptr || p->out_z != nullptr, DIMSUM_ERR_INVALID, "selective_scan_fwd: no output buffer");
    DIMSUM_REQUIRE(p->perm == nullptr || p->z != nullptr, DIMSUM_ERR_INVALID, "selective_scan_fwd: perm needs z/out_z");
    if (p->x != nullptr) {
        DIMSUM_REQUIRE(p->chunk_len == 32, DIMSUM_ERR_INVALID, "selective_scan_fwd: chunk_len must be 32");
        DIMSUM_REQUIRE(p->n_chunks == (p->seqlen + p->chunk_len - 1) / p->chunk_len, DIMSUM_ERR_INVALID,
                       "selective_scan_fwd: n_chunks does not match seqlen / chunk_len");
    }
    DIMSUM_REQUIRE(p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "selective_scan_fwd: batch > 65535");
    if (p->batch == 0) return DIMSUM_OK;

    ScanFwdArgs a;
    a.u = p->u; a.delta = p->delta; a.z = p->z; a.B = p->B; a.C = p->C;
    a.A = reinterpret_cast<const float *>(p->A);
    a.D = reinterpret_cast<const float *>(p->D);
    a.delta_bias = reinterpret_cast<const float *>(p->delta_bias);
    a.out = p->out; a.out_z = p->out_z; a.x = reinterpret_cast<float *>(p->x); a.perm = p->perm;
    a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride;
    a.dl_bs = p->delta_batch_stride; a.dl_ds = p->delta_d_stride;
    a.z_bs = p->z_batch_stride; a.z_ds = p->z_d_stride;
    a.out_bs = p->out_batch_stride; a.out_ds = p->out_d_stride;
    a.oz_bs = p->out_z_batch_stride; a.oz_ds = p->out_z_d_stride;
    a.A_ds = p->A_d_stride; a.A_ns = p->A_dstate_stride;
    a.B_bs = p->B_batch_stride; a.B_gs = p->B_group_stride; a.B_ns = p->B_dstate_stride;
    a.C_bs = p->C_batch_stride; a.C_gs = p->C_group_stride; a.C_ns = p->C_dstate_stride;
    a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.dstate = (int)p->dstate; a.n_groups = (int)p->n_groups;
    a.n_chunks = (int)p->n_chunks; a.chunk_len = (int)(p->chunk_len > 0 ? p->chunk_len : 2048);
    a.softplus = p->delta_softplus != 0;

    const int esz = p->io_dtype == DIMSUM_F32 ? 4 : 2;
    const int vec = 16 / esz;
    auto strides_ok = [&](int64_t bs, int64_t ds) { return bs % vec == 0 && ds % vec == 0; };
    bool vec_ok = p->seqlen % vec == 0 && aligned16(p->u) && aligned16(p->delta) && strides_ok(a.u_bs, a.u_ds) &&
                  strides_ok(a.dl_bs, a.dl_ds);
    if (p->out) vec_ok = vec_ok && aligned16(p->out) && strides_ok(a.out_bs, a.out_ds);
    if (p->z) vec_ok = vec_ok && aligned16(p->z) && aligned16(p->out_z) && strides_ok(a.z_bs, a.z_ds) && strides_ok(a.oz_bs, a.oz_ds);

    switch (p->io_dtype) {
        case DIMSUM_F32: return dispatch<float, 0, 5>(a, (int)p->batch, vec_ok, stream);
        case DIMSUM_BF16: return dispatch<__nv_bfloat16, 0, 3>(a, (int)p->batch, vec_ok, stream);
        case DIMSUM_F16: return dispatch<__half, 0, 3>(a, (int)p->batch, vec_ok, stream);
        default: return fail(DIMSUM_ERR_INVALID, "selective_scan_fwd: unknown io_dtype %lld", (long long)p->io_dtype);
    }
}
