// Selective-scan backward for sm_100a.
//
// Replaces selective_scan_bwd_kernel (mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489).
//
// Maths per (row, state n), a_l = exp(delta_l A_n), b_l = delta_l u_l B_{l,n}, h_l = a_l h_{l-1} + b_l:
//     g_l      = C_{l,n} dy_l + a_{l+1} g_{l+1}                 (reverse recurrence, g_L = 0)
//     e_l      = g_l a_l h_{l-1}
//     dA_n    += e_l delta_l            ddelta_l += e_l A_n + g_l u_l B_{l,n}        du_l += g_l delta_l B_{l,n}
//     dB_{l,n} += g_l delta_l u_l  (sum over rows)               dC_{l,n} += dy_l h_l (sum over rows)
// with dy = dout*silu(z), dz = dout*out*silu'(z), du += D dy, dD += dy u, and the softplus chain rule on ddelta.
//
// Mapping: a thread owns ONE ROW x ONE PACKED STATE PAIR (f32x2), a warp = 4 rows x 8 pairs, a CTA = 32 rows.
// The sequence is walked backwards in 16-step sub-chunks: the forward's checkpoints (x, every 16 steps) restart
// the recurrence, the 16 h's and a's of the sub-chunk live in registers, the reverse sweep consumes them.
// Cross-thread sums use "transposed" butterflies (each lane ends up owning one fully reduced value):
// 3 SHFL per step for ddelta/du over the 8 pairs, 3 SHFL per step for dB/dC over the 4 rows of the warp; the 8
// warps then write disjoint shared slices (no shared atomics) which are summed and sent to global memory as one
// fp32 atomic per (l, n) per CTA -- 1/32 of the reference's atomic traffic.
#include "common.cuh"

namespace dimsum {
namespace {

constexpr int kRowsB = 32;         // rows per CTA
constexpr int kPairs = 8;          // 16 padded states
constexpr int kThreadsB = kRowsB * kPairs;
constexpr int kSub = 16;           // steps per sub-chunk
constexpr int kPitch = kSub + 4;   // padded row pitch of the [row][l] tiles (words)

struct ScanBwdArgs {
    const void *u, *delta, *z, *B, *C, *dout, *out;
    const float *A, *D, *delta_bias, *x;
    void *du, *ddelta, *dz, *out_z;
    float *dA, *dB, *dC, *dD, *ddelta_bias;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, g_bs, g_ds;
    int64_t du_bs, du_ds, dd_bs, dd_ds, dz_bs, dz_ds, oz_bs, oz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, dB_bs, dB_gs, dB_ns, dC_bs, dC_gs, dC_ns;
    int dim, seqlen, dstate, n_groups, n_chunks, softplus;
};

struct BwdSmem {
    float u[kRowsB][kPitch];
    float dl[kRowsB][kPitch];
    float dy[kRowsB][kPitch];
    float sig[kRowsB][kPitch];      // d softplus / d raw delta
    float ddl[kRowsB][kPitch];      // outputs
    float du[kRowsB][kPitch];
    float Bs[kSub][kPitch];
    float Cs[kSub][kPitch];
    float part[kThreadsB / 32][kSub][33];   // per-warp dB/dC partial sums: [l][0..15] = dB_n, [l][16..31] = dC_n (+1 pad)
};

template <typename T, bool kHasZ>
__global__ void __launch_bounds__(kThreadsB, 2) scan_bwd_kernel(const ScanBwdArgs a) {
    __shared__ __align__(16) BwdSmem s;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int row = tid >> 3, p = tid & 7;
    const int b = blockIdx.y;
    const int dpg = a.dim / a.n_groups;
    const int slabs_per_group = (dpg + kRowsB - 1) / kRowsB;
    const int g = blockIdx.x / slabs_per_group;
    const int d0 = g * dpg + (blockIdx.x % slabs_per_group) * kRowsB;
    const int nrows = min(kRowsB, (g + 1) * dpg - d0);
    const int L = a.seqlen;
    const bool row_ok = row < nrows;
    const int d = d0 + row;

    const T *u = reinterpret_cast<const T *>(a.u) + b * a.u_bs + (int64_t)d0 * a.u_ds;
    const T *dl = reinterpret_cast<const T *>(a.delta) + b * a.dl_bs + (int64_t)d0 * a.dl_ds;
    const T *go = reinterpret_cast<const T *>(a.dout) + b * a.g_bs + (int64_t)d0 * a.g_ds;
    const T *Bg = reinterpret_cast<const T *>(a.B) + b * a.B_bs + g * a.B_gs;
    const T *Cg = reinterpret_cast<const T *>(a.C) + b * a.C_bs + g * a.C_gs;

    // per-thread constants: A (natural and log2e-scaled) for this state pair
    float2 An = make_float2(0.f, 0.f);
    if (row_ok) {
        const float *Arow = a.A + (int64_t)d * a.A_ds;
        if (2 * p < a.dstate) An.x = Arow[(2 * p) * a.A_ns];
        if (2 * p + 1 < a.dstate) An.y = Arow[(2 * p + 1) * a.A_ns];
    }
    const float2 A2 = make_float2(An.x * kLog2e, An.y * kLog2e);
    const float Dv = (row_ok && a.D != nullptr) ? a.D[d] : 0.f;
    float2 dA = make_float2(0.f, 0.f);
    float2 carry = make_float2(0.f, 0.f);     // a_{l+1} g_{l+1} entering the current step from the right
    float dD_acc = 0.f, dbias_acc = 0.f;

    const int n_sub = (L + kSub - 1) / kSub;
    for (int k = n_sub - 1; k >= 0; --k) {
        const int l0 = k * kSub;
        // ------------------------------------------------------------ (a) coalesced loads + elementwise prep
        for (int idx = tid; idx < kRowsB * kSub; idx += kThreadsB) {
            const int r = idx / kSub, col = idx % kSub;
            float uv = 0.f, dv = 0.f, dyv = 0.f, sg = 0.f;
            if (r < nrows && l0 + col < L) {
                const int l = l0 + col;
                uv = Io<T>::ld(u + (int64_t)r * a.u_ds + l);
                const float raw = Io<T>::ld(dl + (int64_t)r * a.dl_ds + l) + (a.delta_bias != nullptr ? a.delta_bias[d0 + r] : 0.f);
                if (a.softplus) {
                    dv = softplus_f(raw);
                    sg = raw <= 20.f ? sigmoid_f(raw) : 1.f;
                } else {
                    dv = raw;
                    sg = 1.f;
                }
                dyv = Io<T>::ld(go + (int64_t)r * a.g_ds + l);
                if (kHasZ) {
                    const float zv = Io<T>::ld(reinterpret_cast<const T *>(a.z) + b * a.z_bs + (int64_t)(d0 + r) * a.z_ds + l);
                    const float ov = Io<T>::ld(reinterpret_cast<const T *>(a.out) + b * a.o_bs + (int64_t)(d0 + r) * a.o_ds + l);
                    const float sz = sigmoid_f(zv);
                    const float silu = zv * sz;
                    // d/dz [z sigmoid(z)] = sigmoid(z) (1 + z (1 - sigmoid(z)))
                    Io<T>::st(reinterpret_cast<T *>(a.dz) + b * a.dz_bs + (int64_t)(d0 + r) * a.dz_ds + l,
                              dyv * ov * sz * fmaf(zv, 1.f - sz, 1.f));
                    if (a.out_z != nullptr)
                        Io<T>::st(reinterpret_cast<T *>(a.out_z) + b * a.oz_bs + (int64_t)(d0 + r) * a.oz_ds + l, ov * silu);
                    dyv *= silu;
                }
            }
            s.u[r][col] = uv; s.dl[r][col] = dv; s.dy[r][col] = dyv; s.sig[r][col] = sg;
        }
        {
            const int l = tid % kSub, n = tid / kSub;    // 256 threads = 16 l x 16 n
            const bool ok = n < a.dstate && l0 + l < L;
            s.Bs[l][n] = ok ? Io<T>::ld(Bg + n * a.B_ns + l0 + l) : 0.f;
            s.Cs[l][n] = ok ? Io<T>::ld(Cg + n * a.C_ns + l0 + l) : 0.f;
        }
        __syncthreads();

        // ------------------------------------------------------------ (b) recompute the sub-chunk forward
        float2 hist[kSub + 1], dec[kSub];
        hist[0] = make_float2(0.f, 0.f);
        if (k > 0 && row_ok) {
            const int steps = l0;                                   // state after `steps` steps
            const int ck = (steps % 32 == 0) ? steps / 32 - 1 : steps / 32;
            const int slot = (steps % 32 == 0) ? 1 : 0;
            const float *xp = a.x + (((int64_t)b * a.dim + d) * a.n_chunks + ck) * (2 * a.dstate) + slot;
            if (2 * p < a.dstate) hist[0].x = xp[4 * p];
            if (2 * p + 1 < a.dstate) hist[0].y = xp[4 * p + 2];
        }
#pragma unroll
        for (int i = 0; i < kSub; ++i) {
            const float dlt = s.dl[row][i];
            const float2 Bp = *reinterpret_cast<const float2 *>(&s.Bs[i][2 * p]);
            const float2 t = mul2(splat2(dlt), A2);
            dec[i] = make_float2(ex2_mufu(t.x), ex2_mufu(t.y));
            hist[i + 1] = fma2(dec[i], hist[i], mul2(splat2(dlt * s.u[row][i]), Bp));
        }
        // ------------------------------------------------------------ (c) reverse sweep
#pragma unroll
        for (int i = kSub - 1; i >= 0; --i) {
            const float dlt = s.dl[row][i], uv = s.u[row][i], dyv = s.dy[row][i];
            const float2 Bp = *reinterpret_cast<const float2 *>(&s.Bs[i][2 * p]);
            const float2 Cp = *reinterpret_cast<const float2 *>(&s.Cs[i][2 * p]);
            const float2 gl = fma2(Cp, splat2(dyv), carry);
            carry = mul2(dec[i], gl);
            const float2 e = mul2(carry, hist[i]);                 // g a h_{l-1}
            dA = fma2(e, splat2(dlt), dA);
            const float2 gB = mul2(gl, Bp);
            const float2 dd2 = fma2(e, An, mul2(gB, splat2(uv)));
            float v_dd = dd2.x + dd2.y;                            // partial ddelta_l over this pair
            float v_du = (gB.x + gB.y) * dlt;                      // partial du_l
            float2 dBp = mul2(gl, splat2(dlt * uv));
            float2 dCp = mul2(hist[i + 1], splat2(dyv));
            // --- pairs butterfly (lane bits 0..2): bit2 = 0 lanes end with ddelta, bit2 = 1 lanes with du
            {
                const bool hi = lane & 4;
                float keep = hi ? v_du : v_dd, send = hi ? v_dd : v_du;
                keep += __shfl_xor_sync(0xffffffffu, send, 4);
                keep += __shfl_xor_sync(0xffffffffu, keep, 2);
                keep += __shfl_xor_sync(0xffffffffu, keep, 1);
                if (p == 0) {
                    const float ddraw = keep * s.sig[row][i];
                    s.ddl[row][i] = ddraw;
                    dbias_acc += ddraw;
                    dD_acc = fmaf(dyv, uv, dD_acc);
                } else if (p == 4) {
                    s.du[row][i] = fmaf(Dv, dyv, keep);
                }
            }
            // --- rows butterfly (lane bits 3..4): each lane ends with one of {dB.x, dB.y, dC.x, dC.y} over 4 rows
            {
                const bool hi16 = lane & 16, hi8 = lane & 8;
                float k0 = hi16 ? dCp.x : dBp.x, k1 = hi16 ? dCp.y : dBp.y;
                const float s0 = hi16 ? dBp.x : dCp.x, s1 = hi16 ? dBp.y : dCp.y;
                k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                float r = hi8 ? k1 : k0;
                const float sd = hi8 ? k0 : k1;
                r += __shfl_xor_sync(0xffffffffu, sd, 8);
                s.part[warp][i][(hi16 ? 16 : 0) + 2 * p + (hi8 ? 1 : 0)] = r;
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ (d) coalesced stores + dB/dC flush
        for (int idx = tid; idx < kRowsB * kSub; idx += kThreadsB) {
            const int r = idx / kSub, col = idx % kSub;
            if (r < nrows && l0 + col < L) {
                const int l = l0 + col;
                Io<T>::st(reinterpret_cast<T *>(a.ddelta) + b * a.dd_bs + (int64_t)(d0 + r) * a.dd_ds + l, s.ddl[r][col]);
                Io<T>::st(reinterpret_cast<T *>(a.du) + b * a.du_bs + (int64_t)(d0 + r) * a.du_ds + l, s.du[r][col]);
            }
        }
        for (int idx = tid; idx < kSub * 32; idx += kThreadsB) {
            const int l = idx % kSub, v = idx / kSub;              // v < 16: dB_n, v >= 16: dC_n ; l fastest for the atomics
            const int n = v & 15;
            if (n < a.dstate && l0 + l < L) {
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < kThreadsB / 32; ++w) sum += s.part[w][l][v];
                float *dst = v < 16 ? a.dB + b * a.dB_bs + g * a.dB_gs + n * a.dB_ns + l0 + l
                                    : a.dC + b * a.dC_bs + g * a.dC_gs + n * a.dC_ns + l0 + l;
                atomicAdd(dst, sum);
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ per-row parameter gradients
    if (row_ok) {
        if (2 * p < a.dstate) atomicAdd(a.dA + (int64_t)d * a.dstate + 2 * p, dA.x);
        if (2 * p + 1 < a.dstate) atomicAdd(a.dA + (int64_t)d * a.dstate + 2 * p + 1, dA.y);
        if (p == 0) {
            if (a.dD != nullptr) atomicAdd(a.dD + d, dD_acc);
            if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, dbias_acc);
        }
    }
}

template <typename T>
int run(const ScanBwdArgs &a, int batch, cudaStream_t stream) {
    const int dpg = a.dim / a.n_groups;
    dim3 grid(a.n_groups * ((dpg + kRowsB - 1) / kRowsB), batch);
    if (a.z != nullptr) {
        scan_bwd_kernel<T, true><<<grid, kThreadsB, 0, stream>>>(a);
    } else {
        scan_bwd_kernel<T, false><<<grid, kThreadsB, 0, stream>>>(a);
    }
    return check_launch("selective_scan_bwd");
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_selective_scan_bwd(const dimsum_scan_bwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "selective_scan_bwd: null params");
    DIMSUM_REQUIRE(p->batch >= 0 && p->dim > 0 && p->seqlen > 0 && p->dstate > 0, DIMSUM_ERR_INVALID, "selective_scan_bwd: bad sizes");
    DIMSUM_REQUIRE(p->dstate <= 256, DIMSUM_ERR_INVALID, "selective_scan only supports state dimension <= 256");
    DIMSUM_REQUIRE(p->dstate <= 2 * kPairs, DIMSUM_ERR_UNSUPPORTED,
                   "selective_scan_bwd: dstate=%lld > 16 is not implemented in the B200 kernels", (long long)p->dstate);
    DIMSUM_REQUIRE(p->n_groups >= 1 && p->dim % p->n_groups == 0, DIMSUM_ERR_INVALID, "selective_scan_bwd: dim %% n_groups != 0");
    DIMSUM_REQUIRE(p->u && p->delta && p->A && p->B && p->C && p->dout && p->x, DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: null input pointer");
    DIMSUM_REQUIRE(p->du && p->ddelta && p->dA && p->dB && p->dC, DIMSUM_ERR_INVALID, "selective_scan_bwd: null output pointer");
    DIMSUM_REQUIRE((p->z == nullptr) || (p->out != nullptr && p->dz != nullptr), DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: z needs out and dz");
    DIMSUM_REQUIRE(p->chunk_len == 32 && p->n_chunks == (p->seqlen + 31) / 32, DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: x must hold 32-step chunks (n_chunks = ceil(seqlen/32))");
    DIMSUM_REQUIRE(p->batch <= 65535, DIMSUM_ERR_UNSUPPORTED, "selective_scan_bwd: batch > 65535");
    if (p->batch == 0) return DIMSUM_OK;

    ScanBwdArgs a;
    a.u = p->u; a.delta = p->delta; a.z = p->z; a.B = p->B; a.C = p->C; a.dout = p->dout; a.out = p->out;
    a.A = reinterpret_cast<const float *>(p->A); a.D = reinterpret_cast<const float *>(p->D);
    a.delta_bias = reinterpret_cast<const float *>(p->delta_bias); a.x = reinterpret_cast<const float *>(p->x);
    a.du = p->du; a.ddelta = p->ddelta; a.dz = p->dz; a.out_z = p->out_z_recompute;
    a.dA = p->dA; a.dB = p->dB; a.dC = p->dC; a.dD = p->dD; a.ddelta_bias = p->ddelta_bias;
    a.u_bs = p->u_batch_stride; a.u_ds = p->u_d_stride; a.dl_bs = p->delta_batch_stride; a.dl_ds = p->delta_d_stride;
    a.z_bs = p->z_batch_stride; a.z_ds = p->z_d_stride; a.o_bs = p->out_batch_stride; a.o_ds = p->out_d_stride;
    a.g_bs = p->dout_batch_stride; a.g_ds = p->dout_d_stride; a.du_bs = p->du_batch_stride; a.du_ds = p->du_d_stride;
    a.dd_bs = p->ddelta_batch_stride; a.dd_ds = p->ddelta_d_stride; a.dz_bs = p->dz_batch_stride; a.dz_ds = p->dz_d_stride;
    a.oz_bs = p->out_z_batch_stride; a.oz_ds = p->out_z_d_stride;
    a.A_ds = p->A_d_stride; a.A_ns = p->A_dstate_stride;
    a.B_bs = p->B_batch_stride; a.B_gs = p->B_group_stride; a.B_ns = p->B_dstate_stride;
    a.C_bs = p->C_batch_stride; a.C_gs = p->C_group_stride; a.C_ns = p->C_dstate_stride;
    a.dB_bs = p->dB_batch_stride; a.dB_gs = p->dB_group_stride; a.dB_ns = p->dB_dstate_stride;
    a.dC_bs = p->dC_batch_stride; a.dC_gs = p->dC_group_stride; a.dC_ns = p->dC_dstate_stride;
    a.dim = (int)p->dim; a.seqlen = (int)p->seqlen; a.dstate = (int)p->dstate; a.n_groups = (int)p->n_groups;
    a.n_chunks = (int)p->n_chunks; a.softplus = p->delta_softplus != 0;
    switch (p->io_dtype) {
        case DIMSUM_F32: return run<float>(a, (int)p->batch, stream);
        case DIMSUM_BF16: return run<__nv_bfloat16>(a, (int)p->batch, stream);
        case DIMSUM_F16: return run<__half>(a, (int)p->batch, stream);
        default: return fail(DIMSUM_ERR_INVALID, "selective_scan_bwd: unknown io_dtype");
    }
}
