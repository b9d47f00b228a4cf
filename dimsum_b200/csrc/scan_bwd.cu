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
// Mapping (third layout of this kernel; the first two put rows on lanes and paid for it in the dB/dC sums):
//   * a lane owns TWO STATES of a PAIR OF ROWS: every recurrence value is a packed f32x2 (row 2p, row 2p+1), so the
//     per-row scalars (delta, delta*u, dy) arrive as natural pairs from interleaved shared tiles and only B / C need a
//     splat (a register move);
//   * 8 lanes (16 states) form a row group that walks its 2 row pairs one after the other, so dB / dC of the 4 rows
//     accumulate in registers and meet the other 15 groups of the CTA once per 4 steps through shared memory;
//   * what has to cross lanes per step is only the two 16-state dot products (sum_n e A, sum_n g B): 8 packed values per
//     lane and 4 steps go through a padded transposing tile, lane j ends up with total j.
// A CTA owns 32 rows (64 threads; 64 rows / 128 threads in the round-1 shape) and walks the sequence backwards in 16-step chunks restarted from the forward's 16-step
// checkpoints (x): one forward sweep parks the state every 4 steps, then the four 4-step mini-chunks are replayed
// (h and a history in registers) and swept in reverse.  Global memory sees ONE fp32 atomic per (l, n) per 64 rows, and
// u / delta / dout / z / out / du / ddelta / dz move as coalesced 16-byte vectors.
#include <cstdlib>

#include "common.cuh"

namespace dimsum {
namespace {

// threads per CTA: 64 (32 rows; 128 registers, 8 CTAs = 16 warps per SM; the default) or 128 (64 rows; 3 CTAs = 12 warps per
// SM at 168 registers, round 1)
constexpr int kLn = 8;                     // lanes per row group (2 states each)
template <int KT> struct Cta {
    static constexpr int kT = KT;
    static constexpr int kGroups = KT / kLn;       // 16 (8) row groups
    static constexpr int kPairs = 2 * kGroups;     // 32 (16) row pairs = 64 (32) rows per CTA
    static constexpr int kRowsB = 2 * kPairs;
    static constexpr int kMinCtas = KT == 128 ? 3 : 7;
};
constexpr int kSub = 16;                   // steps per chunk (== checkpoint spacing of the forward)
constexpr int kMini = 4;                   // steps whose history lives in registers
constexpr int kMinis = kSub / kMini;
constexpr int kTileG = 2 * kMini + 1;      // float4 units per row group inside a slab (2 row pairs x 4 steps + 1: neighbouring
                                           // groups start 4 banks apart, so their 8-byte broadcast loads do not collide)
// float4 units per mini-chunk slab of a [mini][row pair][step] tile; slab stride = 1 mod 8 makes the 16-byte stores of the prep
// pass conflict-free
template <int KT> __device__ __host__ constexpr int tile_m() { return Cta<KT>::kGroups * kTileG + 1; }
// element (mini m, row pair p, step i) of a row-pair tile
template <int KT>
__device__ __forceinline__ constexpr int tile_at(int m, int p, int i) { return m * tile_m<KT>() + (p >> 1) * kTileG + (p & 1) * kMini + i; }
constexpr bool kShflReduce = false;        // true: per-step 16-state dot products by recursive-halving warp shuffles instead of the shared
                                           // transposing tile.  Measured: 3.04 -> 2.99 ms fp32, 2.92 -> 2.86 ms bf16 at 256 x
                                           // 2048 x 256, 0.242 -> 0.264 ms bf16 at the training shape: shuffles ride the same LSU pipe; not enabled
constexpr int kStLane = 20;                // words per lane record of the transposing tile (8 packed values + pad)
constexpr int kStGroup = kLn * kStLane + 16;   // words per group: odd multiple of 16 -> the two groups of a half warp use disjoint banks

struct ScanBwdArgs {
    const void *u, *delta, *z, *B, *C, *dout, *out;
    const float *A, *D, *delta_bias, *x;
    void *du, *ddelta, *dz, *out_z;
    float *dA, *dB, *dC, *dD, *ddelta_bias;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, g_bs, g_ds;
    int64_t du_bs, du_ds, dd_bs, dd_ds, dz_bs, dz_ds, oz_bs, oz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, dB_bs, dB_gs, dB_ns, dC_bs, dC_gs, dC_ns;
    int dim, seqlen, dstate, n_groups, n_chunks, softplus, vec_io, vec_bc;
};

template <int KT>
struct BwdSmem {
    static constexpr int kPairs = Cta<KT>::kPairs, kGroups = Cta<KT>::kGroups;
    // row-pair tiles, element (mini m, row pair p, step i) at [tile_at<KT>(m, p, i)]
    float4 X[kMinis * tile_m<KT>()];      // (delta r0, delta r1, delta*u r0, delta*u r1); later (ddelta r0, r1, du r0, r1)
    float4 Y[kMinis * tile_m<KT>()];      // (dy r0, dy r1, D dy r0, D dy r1)
    float4 Z[kMinis * tile_m<KT>()];      // (ln2 s' r0, ln2 s' r1, u s' r0, u s' r1) with s' = d softplus / d raw delta
    float2 Bd[kSub][kLn];                 // (B_n0, B_n0+1) per lane
    float2 Cd[kSub][kLn];
    float4 hs[kPairs][kSub / kMini - 1][kLn];   // state at the start of mini-chunks 0..2: (n0 r0, n0 r1, n0+1 r0, n0+1 r1)
    float st[kGroups * kStGroup];         // transposing tile of the row groups; reused for the dB/dC hand-over
};

template <typename T, bool kHasZ, int KT>
__global__ void __launch_bounds__(KT, Cta<KT>::kMinCtas) scan_bwd_kernel(const ScanBwdArgs a) {
    constexpr int kT = KT, kGroups = Cta<KT>::kGroups, kPairs = Cta<KT>::kPairs, kRowsB = Cta<KT>::kRowsB;
    constexpr int VEC = kMini;                           // global I/O in 4-element vectors (16 bytes fp32, 8 bytes bf16/fp16)
    constexpr int VPR = kSub / VEC;                      // vectors per row chunk: the 128 threads are 32 row pairs x 4 vectors
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<KT> &s = *reinterpret_cast<BwdSmem<KT> *>(smem_raw);
    const int tid = threadIdx.x;
    const int grp = tid >> 3, ln = tid & 7;
    const int n0 = 2 * ln;                               // first of this lane's two states
    const int b = blockIdx.y;
    const int dpg = a.dim / a.n_groups;
    const int slabs_per_group = (dpg + kRowsB - 1) / kRowsB;
    const int g = blockIdx.x / slabs_per_group;
    const int d0 = g * dpg + (blockIdx.x % slabs_per_group) * kRowsB;
    const int nrows = min(kRowsB, (g + 1) * dpg - d0);
    const int L = a.seqlen;

    const T *u = reinterpret_cast<const T *>(a.u) + b * a.u_bs + (int64_t)d0 * a.u_ds;
    const T *dl = reinterpret_cast<const T *>(a.delta) + b * a.dl_bs + (int64_t)d0 * a.dl_ds;
    const T *go = reinterpret_cast<const T *>(a.dout) + b * a.g_bs + (int64_t)d0 * a.g_ds;
    const T *zz = kHasZ ? reinterpret_cast<const T *>(a.z) + b * a.z_bs + (int64_t)d0 * a.z_ds : nullptr;
    const T *oo = kHasZ ? reinterpret_cast<const T *>(a.out) + b * a.o_bs + (int64_t)d0 * a.o_ds : nullptr;
    T *dzp = kHasZ ? reinterpret_cast<T *>(a.dz) + b * a.dz_bs + (int64_t)d0 * a.dz_ds : nullptr;
    T *ozp = (kHasZ && a.out_z != nullptr) ? reinterpret_cast<T *>(a.out_z) + b * a.oz_bs + (int64_t)d0 * a.oz_ds : nullptr;
    T *dup = reinterpret_cast<T *>(a.du) + b * a.du_bs + (int64_t)d0 * a.du_ds;
    T *ddp = reinterpret_cast<T *>(a.ddelta) + b * a.dd_bs + (int64_t)d0 * a.dd_ds;
    const T *Bg = reinterpret_cast<const T *>(a.B) + b * a.B_bs + g * a.B_gs;
    const T *Cg = reinterpret_cast<const T *>(a.C) + b * a.C_bs + g * a.C_gs;

    // per-thread constants: A (log2e-scaled) of this lane's 2 states for the 2 x 2 rows of its group
    float2 A2[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int si = 0; si < 2; ++si) {
            const int r = grp * 4 + 2 * q, n = n0 + si;
            const float a0 = (r < nrows && n < a.dstate) ? a.A[(int64_t)(d0 + r) * a.A_ds + n * a.A_ns] : 0.f;
            const float a1 = (r + 1 < nrows && n < a.dstate) ? a.A[(int64_t)(d0 + r + 1) * a.A_ds + n * a.A_ns] : 0.f;
            A2[q][si] = make_float2(a0 * kLog2e, a1 * kLog2e);
        }
    }
    float2 dA[2][2], carry[2][2];              // carry = a_{l+1} g_{l+1} entering the current step from the right
    float2 dbias_acc[2];                       // meaningful on the even lanes (they finish the per-step sums)
    float dD_prep[2] = {0.f, 0.f};             // sum_l dy u of the two rows this thread loads in the prep pass
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        dbias_acc[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int si = 0; si < 2; ++si) dA[q][si] = carry[q][si] = make_float2(0.f, 0.f);
    }
    float *const st_grp = s.st + grp * kStGroup;
    // prep / store item of this thread: (row pair, 16-byte vector of steps), vectors fastest -> VPR neighbouring lanes
    // cover one row's 64-byte chunk
    const bool io_thread = tid < kPairs * VPR;
    const int io_rp = tid / VPR, io_col = (tid % VPR) * VEC;
    float Dio[2] = {0.f, 0.f}, bias_io[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int r = 2 * io_rp + c;
        if (io_thread && r < nrows) {
            if (a.D != nullptr) Dio[c] = a.D[d0 + r];
            if (a.delta_bias != nullptr) bias_io[c] = a.delta_bias[d0 + r];
        }
    }

    // one forward step of this lane's 2 states x 2 rows; leaves the decays in dec
    auto fwd_step = [&](const float4 v, const float2 Bv, int q, float2 (&h)[2], float2 (&dec)[2]) {
        const float2 dlt = make_float2(v.x, v.y), dtu = make_float2(v.z, v.w);
        const float2 Bs[2] = {splat2(Bv.x), splat2(Bv.y)};
#pragma unroll
        for (int si = 0; si < 2; ++si) {
            const float2 t = mul2(dlt, A2[q][si]);
            dec[si] = make_float2(ex2_mufu(t.x), ex2_mufu(t.y));
            h[si] = fma2(dec[si], h[si], mul2(dtu, Bs[si]));
        }
    };

    const int n_sub = (L + kSub - 1) / kSub;
    for (int k = n_sub - 1; k >= 0; --k) {
        const int l0 = k * kSub;
        // ------------------------------------------------------------ (a) coalesced loads + elementwise prep
        if (io_thread) {
            const int rp = io_rp, col = io_col, l = l0 + col;
            float dv[2][VEC], tu[2][VEC], gv[2][VEC], dg[2][VEC], sk[2][VEC], su[2][VEC];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int r = 2 * rp + c;
#pragma unroll
                for (int i = 0; i < VEC; ++i) { dv[c][i] = tu[c][i] = gv[c][i] = dg[c][i] = sk[c][i] = su[c][i] = 0.f; }
                if (r < nrows && l < L) {
                    const bool full = a.vec_io && l + VEC <= L;
                    float uv[VEC], zv[VEC], ov[VEC], raw[VEC];
                    if (full) {
                        Io<T>::ld4(u + (int64_t)r * a.u_ds + l, uv);
                        Io<T>::ld4(dl + (int64_t)r * a.dl_ds + l, raw);
                        Io<T>::ld4(go + (int64_t)r * a.g_ds + l, gv[c]);
                        if (kHasZ) {
                            Io<T>::ld4(zz + (int64_t)r * a.z_ds + l, zv);
                            Io<T>::ld4(oo + (int64_t)r * a.o_ds + l, ov);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const bool ok = l + i < L;
                            uv[i] = ok ? Io<T>::ld(u + (int64_t)r * a.u_ds + l + i) : 0.f;
                            raw[i] = ok ? Io<T>::ld(dl + (int64_t)r * a.dl_ds + l + i) : 0.f;
                            gv[c][i] = ok ? Io<T>::ld(go + (int64_t)r * a.g_ds + l + i) : 0.f;
                            if (kHasZ) {
                                zv[i] = ok ? Io<T>::ld(zz + (int64_t)r * a.z_ds + l + i) : 0.f;
                                ov[i] = ok ? Io<T>::ld(oo + (int64_t)r * a.o_ds + l + i) : 0.f;
                            }
                        }
                    }
                    float dzv[VEC], ozv[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float x = raw[i] + bias_io[c];
                        float sgm = 1.f;
                        if (a.softplus) {
                            dv[c][i] = softplus_f(x);
                            sgm = x <= 20.f ? sigmoid_f(x) : 1.f;
                        } else {
                            dv[c][i] = x;
                        }
                        if (kHasZ) {
                            const float sz = sigmoid_f(zv[i]);
                            const float silu = zv[i] * sz;
                            dzv[i] = gv[c][i] * ov[i] * sz * fmaf(zv[i], 1.f - sz, 1.f);   // d/dz [z sigmoid(z)]
                            ozv[i] = ov[i] * silu;
                            gv[c][i] *= silu;
                        }
                        if (l + i >= L) { dv[c][i] = 0.f; uv[i] = 0.f; gv[c][i] = 0.f; sgm = 0.f; }
                        tu[c][i] = dv[c][i] * uv[i];
                        dg[c][i] = Dio[c] * gv[c][i];
                        sk[c][i] = kLn2 * sgm;
                        su[c][i] = uv[i] * sgm;
                        dD_prep[c] = fmaf(gv[c][i], uv[i], dD_prep[c]);
                    }
                    if (kHasZ) {
                        if (full) {
                            Io<T>::st4(dzp + (int64_t)r * a.dz_ds + l, dzv);
                            if (ozp != nullptr) Io<T>::st4(ozp + (int64_t)r * a.oz_ds + l, ozv);
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) {
                                if (l + i < L) {
                                    Io<T>::st(dzp + (int64_t)r * a.dz_ds + l + i, dzv[i]);
                                    if (ozp != nullptr) Io<T>::st(ozp + (int64_t)r * a.oz_ds + l + i, ozv[i]);
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int e = tile_at<KT>(col / kMini + i / kMini, rp, i % kMini);
                s.X[e] = make_float4(dv[0][i], dv[1][i], tu[0][i], tu[1][i]);
                s.Y[e] = make_float4(gv[0][i], gv[1][i], dg[0][i], dg[1][i]);
                s.Z[e] = make_float4(sk[0][i], sk[1][i], su[0][i], su[1][i]);
            }
        }
        // B / C tiles [step][state]: a lane's 8-byte load is its two states (the splats over the row pair are register moves:
        // the LSU pipe is the busy one, the issue slots are not)
        for (int idx = tid; idx < 2 * 16 * VPR; idx += kT) {
            const int which = idx / (16 * VPR), rem = idx % (16 * VPR);
            const int n = rem % 16, col = (rem / 16) * VEC;
            const T *src = (which ? Cg + n * a.C_ns : Bg + n * a.B_ns) + l0 + col;
            float bv[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) bv[i] = 0.f;
            if (n < a.dstate && l0 + col < L) {
                if (a.vec_bc && l0 + col + VEC <= L) {
                    Io<T>::ld4(src, bv);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) bv[i] = (l0 + col + i < L) ? Io<T>::ld(src + i) : 0.f;
                }
            }
            float *tile = which ? &s.Cd[0][0].x : &s.Bd[0][0].x;
#pragma unroll
            for (int i = 0; i < VEC; ++i) tile[(col + i) * 2 * kLn + n] = bv[i];
        }
        __syncthreads();

        // ------------------------------------------------------------ (b) one forward sweep: mini-chunk start states
        float2 h3[2][2];                                        // state at the start of the last mini-chunk
        {
            float2 h[2][2], dtmp[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                h[q][0] = h[q][1] = make_float2(0.f, 0.f);
                if (k > 0) {                                    // state after l0 steps: planar checkpoint record
                    const int ck = (l0 - 1) / 32, half = (l0 % 32 == 16) ? 0 : 1;
                    const int r = grp * 4 + 2 * q;
                    const float *xp = a.x + (((int64_t)b * a.dim + d0 + r) * a.n_chunks + ck) * (2 * a.dstate) + half * a.dstate;
                    const int64_t next = (int64_t)a.n_chunks * 2 * a.dstate;
#pragma unroll
                    for (int si = 0; si < 2; ++si) {
                        const int n = n0 + si;
                        if (n < a.dstate) {
                            if (r < nrows) h[q][si].x = xp[n];
                            if (r + 1 < nrows) h[q][si].y = xp[next + n];
                        }
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < kMinis - 1; ++m) {
#pragma unroll
                for (int q = 0; q < 2; ++q) s.hs[grp * 2 + q][m][ln] = make_float4(h[q][0].x, h[q][0].y, h[q][1].x, h[q][1].y);
#pragma unroll
                for (int i = 0; i < kMini; ++i) {
                    const float2 Bv = s.Bd[m * kMini + i][ln];           // one load serves both row pairs
#pragma unroll
                    for (int q = 0; q < 2; ++q) fwd_step(s.X[tile_at<KT>(m, grp * 2 + q, i)], Bv, q, h[q], dtmp);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) { h3[q][0] = h[q][0]; h3[q][1] = h[q][1]; }
        }

        // ------------------------------------------------------------ (c) mini-chunks in reverse
#pragma unroll 1
        for (int m = kMinis - 1; m >= 0; --m) {
            float accB[kMini][2], accC[kMini][2];               // dB / dC of this group's 4 rows
            float2 Bm[kMini], Cm[kMini];                        // B / C of the 4 steps: loaded once, used by both row pairs
#pragma unroll
            for (int i = 0; i < kMini; ++i) {
                accB[i][0] = accB[i][1] = accC[i][0] = accC[i][1] = 0.f;
                Bm[i] = s.Bd[m * kMini + i][ln];
                Cm[i] = s.Cd[m * kMini + i][ln];
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e0 = tile_at<KT>(m, grp * 2 + q, 0);
                float2 hist[kMini + 1][2], dec[kMini][2];
                float4 xv[kMini];
                if (m == kMinis - 1) {
                    hist[0][0] = h3[q][0];
                    hist[0][1] = h3[q][1];
                } else {
                    const float4 c = s.hs[grp * 2 + q][m][ln];
                    hist[0][0] = make_float2(c.x, c.y);
                    hist[0][1] = make_float2(c.z, c.w);
                }
#pragma unroll
                for (int i = 0; i < kMini; ++i) {               // replay the 4 steps, keeping h and a
                    hist[i + 1][0] = hist[i][0];
                    hist[i + 1][1] = hist[i][1];
                    xv[i] = s.X[e0 + i];
                    fwd_step(xv[i], Bm[i], q, hist[i + 1], dec[i]);
                }
                float4 pq_prev = make_float4(0.f, 0.f, 0.f, 0.f), pq_a = pq_prev, pq_b = pq_prev;
#pragma unroll
                for (int i = kMini - 1; i >= 0; --i) {
                    const float2 dyv = *reinterpret_cast<const float2 *>(&s.Y[e0 + i]);
                    const float2 dlt = make_float2(xv[i].x, xv[i].y), dtu = make_float2(xv[i].z, xv[i].w);
                    const float2 Bs[2] = {splat2(Bm[i].x), splat2(Bm[i].y)};
                    const float2 Cs[2] = {splat2(Cm[i].x), splat2(Cm[i].y)};
                    float2 P = make_float2(0.f, 0.f), Q = make_float2(0.f, 0.f);
#pragma unroll
                    for (int si = 0; si < 2; ++si) {
                        const float2 gl = fma2(Cs[si], dyv, carry[q][si]);
                        carry[q][si] = mul2(dec[i][si], gl);
                        const float2 e = mul2(carry[q][si], hist[i][si]);     // g a h_{l-1}
                        dA[q][si] = fma2(e, dlt, dA[q][si]);
                        P = fma2(e, A2[q][si], P);
                        Q = fma2(gl, Bs[si], Q);
                        accB[i][si] = fmaf(gl.y, dtu.y, fmaf(gl.x, dtu.x, accB[i][si]));
                        accC[i][si] = fmaf(hist[i + 1][si].y, dyv.y, fmaf(hist[i + 1][si].x, dyv.x, accC[i][si]));
                    }
                    if (!kShflReduce) {
                        *reinterpret_cast<float4 *>(st_grp + ln * kStLane + 4 * i) = make_float4(P.x, P.y, Q.x, Q.y);
                    } else {
                        // recursive halving over the 8 lanes of the group, folded into the step loop so that at most two
                        // (P, Q) sets are alive: steps 3|2 and 1|0 are split over lane bit 2 as soon as both exist
                        const float4 cur = make_float4(P.x, P.y, Q.x, Q.y);
                        if (i & 1) {
                            pq_prev = cur;
                        } else {
                            const bool hi = (ln & 4) != 0;                    // hi lanes keep step i, low lanes step i + 1
                            const float4 send = hi ? pq_prev : cur, keep = hi ? cur : pq_prev;
                            const float4 got = make_float4(__shfl_xor_sync(0xffffffffu, send.x, 4), __shfl_xor_sync(0xffffffffu, send.y, 4),
                                                           __shfl_xor_sync(0xffffffffu, send.z, 4), __shfl_xor_sync(0xffffffffu, send.w, 4));
                            const float4 sum = make_float4(keep.x + got.x, keep.y + got.y, keep.z + got.z, keep.w + got.w);
                            if (i == 2) pq_a = sum; else pq_b = sum;
                        }
                    }
                }
                float2 tot;
                if (!kShflReduce) {
                    __syncwarp();
                    // transpose: lane j = 2 i + w collects sum_n of value w (0: e A, 1: g B) of step i, both rows
                    tot = *reinterpret_cast<const float2 *>(st_grp + 2 * ln);
#pragma unroll
                    for (int o = 1; o < kLn; ++o) tot = add2(tot, *reinterpret_cast<const float2 *>(st_grp + o * kStLane + 2 * ln));
                } else {
                    const bool hi1 = (ln & 2) != 0;                           // lane bit 1: steps 1|0 (set b) vs 3|2 (set a)
                    const float4 send = hi1 ? pq_a : pq_b, keep = hi1 ? pq_b : pq_a;
                    const float4 k2 = make_float4(keep.x + __shfl_xor_sync(0xffffffffu, send.x, 2), keep.y + __shfl_xor_sync(0xffffffffu, send.y, 2),
                                                  keep.z + __shfl_xor_sync(0xffffffffu, send.z, 2), keep.w + __shfl_xor_sync(0xffffffffu, send.w, 2));
                    const bool odd = (ln & 1) != 0;                           // lane bit 0: Q (odd) vs P (even)
                    const float sx = odd ? k2.x : k2.z, sy = odd ? k2.y : k2.w;
                    tot = make_float2((odd ? k2.z : k2.x) + __shfl_xor_sync(0xffffffffu, sx, 1),
                                      (odd ? k2.w : k2.y) + __shfl_xor_sync(0xffffffffu, sy, 1));
                }
                const float qx = __shfl_down_sync(0xffffffffu, tot.x, 1), qy = __shfl_down_sync(0xffffffffu, tot.y, 1);
                if ((ln & 1) == 0) {
                    // ddelta = (ln2 sum_n e A2 + u sum_n g B) s' ; du = delta sum_n g B + D dy
                    // (shuffle reduction: lane bits 1, 2 select step 3 - (2 bit1 + bit2); tile reduction: step = lane / 2)
                    const int e = e0 + (kShflReduce ? 3 - (2 * ((ln >> 1) & 1) + ((ln >> 2) & 1)) : (ln >> 1));
                    const float4 xr = s.X[e], yr = s.Y[e], zr = s.Z[e];
                    const float dd0 = fmaf(zr.z, qx, tot.x * zr.x), dd1 = fmaf(zr.w, qy, tot.y * zr.y);
                    s.X[e] = make_float4(dd0, dd1, fmaf(xr.x, qx, yr.z), fmaf(xr.y, qy, yr.w));
                    dbias_acc[q].x += dd0;
                    dbias_acc[q].y += dd1;
                }
                if (!kShflReduce) __syncwarp();
            }
            // dB / dC of the 4 steps: each group leaves its 64 sums in its own tile, 128 threads add the 16 groups
            // (thread = (B|C, state, step), steps fastest so that 4 lanes hit 16 contiguous bytes), one atomic each
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int si = c & 1;
                const float4 f = c < 2 ? make_float4(accB[0][si], accB[1][si], accB[2][si], accB[3][si])
                                       : make_float4(accC[0][si], accC[1][si], accC[2][si], accC[3][si]);
                *reinterpret_cast<float4 *>(st_grp + (c * kLn + ln) * 4) = f;
            }
            __syncthreads();
#pragma unroll
            for (int t = tid; t < 128; t += kT) {                // 128 sums = (B|C, state, step)
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int gg = 0; gg < kGroups; gg += 4) {
                    s0 += s.st[gg * kStGroup + t];
                    s1 += s.st[(gg + 1) * kStGroup + t];
                    s2 += s.st[(gg + 2) * kStGroup + t];
                    s3 += s.st[(gg + 3) * kStGroup + t];
                }
                const int c = t >> 5, n = 2 * ((t >> 2) & 7) + (c & 1), l = l0 + m * kMini + (t & 3);
                if (n < a.dstate && l < L) {
                    float *dst = c < 2 ? a.dB + b * a.dB_bs + g * a.dB_gs + n * a.dB_ns + l
                                       : a.dC + b * a.dC_bs + g * a.dC_gs + n * a.dC_ns + l;
                    atomicAdd(dst, (s0 + s1) + (s2 + s3));
                }
            }
            __syncthreads();
        }

        // ------------------------------------------------------------ (d) coalesced stores of ddelta / du
        if (io_thread) {
            const int rp = io_rp, col = io_col, l = l0 + col;
            float dd[2][VEC], du_[2][VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float4 o = s.X[tile_at<KT>(col / kMini + i / kMini, rp, i % kMini)];
                dd[0][i] = o.x; dd[1][i] = o.y; du_[0][i] = o.z; du_[1][i] = o.w;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int r = 2 * rp + c;
                if (r < nrows && l < L) {
                    if (a.vec_io && l + VEC <= L) {
                        Io<T>::st4(ddp + (int64_t)r * a.dd_ds + l, dd[c]);
                        Io<T>::st4(dup + (int64_t)r * a.du_ds + l, du_[c]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            if (l + i < L) {
                                Io<T>::st(ddp + (int64_t)r * a.dd_ds + l + i, dd[c][i]);
                                Io<T>::st(dup + (int64_t)r * a.du_ds + l + i, du_[c][i]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ per-row parameter gradients
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int r = grp * 4 + 2 * q;
#pragma unroll
        for (int si = 0; si < 2; ++si) {
            const int n = n0 + si;
            if (n < a.dstate) {
                if (r < nrows) atomicAdd(a.dA + (int64_t)(d0 + r) * a.dstate + n, dA[q][si].x);
                if (r + 1 < nrows) atomicAdd(a.dA + (int64_t)(d0 + r + 1) * a.dstate + n, dA[q][si].y);
            }
        }
        if (a.ddelta_bias != nullptr) {                   // the four even lanes each saw one step of every mini-chunk
            float2 db = dbias_acc[q];
#pragma unroll
            for (int o = 2; o <= 4; o <<= 1) {
                db.x += __shfl_xor_sync(0xffffffffu, db.x, o);
                db.y += __shfl_xor_sync(0xffffffffu, db.y, o);
            }
            if (ln == 0) {
                if (r < nrows) atomicAdd(a.ddelta_bias + d0 + r, db.x);
                if (r + 1 < nrows) atomicAdd(a.ddelta_bias + d0 + r + 1, db.y);
            }
        }
    }
    if (a.dD != nullptr && tid < kPairs * VPR) {          // warp-uniform: whole warps take part in the shuffles
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v = dD_prep[c];
#pragma unroll
            for (int o = 1; o < VPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int r = 2 * io_rp + c;
            if (tid % VPR == 0 && r < nrows) atomicAdd(a.dD + d0 + r, v);
        }
    }
}

template <typename T, int KT>
int run_width(const ScanBwdArgs &a, int batch, cudaStream_t stream) {
    const int dpg = a.dim / a.n_groups;
    dim3 grid(a.n_groups * ((dpg + Cta<KT>::kRowsB - 1) / Cta<KT>::kRowsB), batch);
    const int smem = (int)sizeof(BwdSmem<KT>);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<grid, KT, smem, stream>>>(a);
    };
    if (a.z != nullptr) go(scan_bwd_kernel<T, true, KT>); else go(scan_bwd_kernel<T, false, KT>);
    return check_launch("selective_scan_bwd");
}

template <typename T>
int run(const ScanBwdArgs &a, int batch, cudaStream_t stream) {
    // 64-thread CTAs (32 rows each; under __launch_bounds__(64, 7) ptxas settles on 128 registers, so 8 CTAs = 16 warps are
    // resident per SM) beat the 128-thread CTAs of round 1 (3 per SM = 12 warps at 168 registers) at every shape -- the kernel
    // is latency-bound, more resident warps and finer CTAs help:
    // 3.23 -> 3.04 ms fp32 / 3.16 -> 2.92 ms bf16 at 256 x 2048 x 256, 0.34 -> 0.26 ms at the 32-latent training shape, whose
    // grid also fits one wave now (1024 CTAs instead of 512 on 444 slots).  Asking for 8 CTAs per SM in the launch bound gave a
    // slower schedule (3.34 ms) and so did 32-thread CTAs (3.33 ms).  DIMSUM_SCAN_BWD_THREADS=128 runs the round-1 shape.
    static const int forced = [] { const char *e = getenv("DIMSUM_SCAN_BWD_THREADS"); return e ? atoi(e) : 0; }();
    return forced == 128 ? run_width<T, 128>(a, batch, stream) : run_width<T, 64>(a, batch, stream);
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_selective_scan_bwd(const dimsum_scan_bwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "selective_scan_bwd: null params");
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
    DIMSUM_REQUIRE(p->batch >= 0 && p->dim > 0 && p->seqlen > 0 && p->dstate > 0, DIMSUM_ERR_INVALID, "selective_scan_bwd: bad sizes");
    DIMSUM_REQUIRE(p->dstate <= 256, DIMSUM_ERR_INVALID, "selective_scan only supports state dimension <= 256");
    DIMSUM_REQUIRE(p->dstate <= 16, DIMSUM_ERR_UNSUPPORTED,
                   "selective_scan_bwd: dstate=%lld > 16 is not implemented in the B200 kernels", (long long)p->dstate);
    DIMSUM_REQUIRE(p->n_groups >= 1 && p->dim % p->n_groups == 0, DIMSUM_ERR_INVALID, "selective_scan_bwd: dim %% n_groups != 0");
    DIMSUM_REQUIRE(p->u && p->delta && p->A && p->B && p->C && p->dout && p->x, DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: null input pointer");
    DIMSUM_REQUIRE(p->du && p->ddelta && p->dA && p->dB && p->dC, DIMSUM_ERR_INVALID, "selective_scan_bwd: null output pointer");
    DIMSUM_REQUIRE((p->z == nullptr) || (p->out != nullptr && p->dz != nullptr), DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: z needs out and dz");
    DIMSUM_REQUIRE(p->chunk_len == 32 && p->n_chunks == (p->seqlen + 31) / 32 + 1, DIMSUM_ERR_INVALID,
                   "selective_scan_bwd: x must hold ceil(seqlen/32) checkpoint records plus the final-state record");
    DIMSUM_REQUIRE(p->io_dtype >= 0 && p->io_dtype <= 2, DIMSUM_ERR_INVALID, "selective_scan_bwd: unknown io_dtype");
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

    const int vec = 4;                                    // elements per vector access; 16 bytes fp32, 8 bytes for 16-bit types
    const uintptr_t amask = p->io_dtype == DIMSUM_F32 ? 15u : 7u;
    auto al = [&](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & amask) == 0; };
    auto ok = [&](const void *ptr, int64_t bs, int64_t ds) { return ptr == nullptr || (al(ptr) && bs % vec == 0 && ds % vec == 0); };
    a.vec_io = ok(p->u, a.u_bs, a.u_ds) && ok(p->delta, a.dl_bs, a.dl_ds) && ok(p->dout, a.g_bs, a.g_ds) &&
               ok(p->z, a.z_bs, a.z_ds) && ok(p->out, a.o_bs, a.o_ds) && ok(p->dz, a.dz_bs, a.dz_ds) &&
               ok(p->out_z_recompute, a.oz_bs, a.oz_ds) && ok(p->du, a.du_bs, a.du_ds) && ok(p->ddelta, a.dd_bs, a.dd_ds);
    a.vec_bc = al(p->B) && al(p->C) && a.B_bs % vec == 0 && a.B_gs % vec == 0 && a.B_ns % vec == 0 &&
               a.C_bs % vec == 0 && a.C_gs % vec == 0 && a.C_ns % vec == 0;
    switch (p->io_dtype) {
        case DIMSUM_F32: return run<float>(a, (int)p->batch, stream);
        case DIMSUM_BF16: return run<__nv_bfloat16>(a, (int)p->batch, stream);
        default: return run<__half>(a, (int)p->batch, stream);
    }
}
