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
// Mapping: TWO threads per channel row, each owning 8 states as 4 packed f32x2 pairs; a warp = 16 rows, a CTA = 64 rows.
// The sequence is walked backwards in 16-step chunks restarted from the forward's 16-step checkpoints (x).  Inside a
// chunk the thread first runs the recurrence forward once, parking the state every 4 steps in shared memory, then
// handles the four 4-step mini-chunks in reverse: recompute 4 steps (h and a history: 36 packed registers), sweep them
// backwards.  Per step the two halves of a row meet with two xor-1 shuffles (ddelta, du); the 32 dB / dC partials of a
// warp's 16 rows are summed through a private padded shared tile (4 STS.128 + 4 LDS.128 + 2 shuffle rounds) and
// accumulated per CTA, so global memory sees ONE fp32 atomic per (l, n) per 64 rows -- 1/64 of the reference's atomic
// traffic -- and all of u / delta / dout / z / out / du / ddelta / dz move as coalesced 16-byte vectors.
#include "common.cuh"

namespace dimsum {
namespace {

constexpr int kRowsB = 64;                 // rows per CTA
constexpr int kThreadsB = 2 * kRowsB;      // two threads per row
constexpr int kSub = 16;                   // steps per chunk (== checkpoint spacing of the forward)
constexpr int kMini = 4;                   // steps whose history lives in registers
constexpr int kPitch = kSub + 4;           // padded row pitch of the [row][l] tiles (words)
constexpr int kCkPitch = 4 * 16 + 4;       // [row][4 mini-chunk starts][16 states] (+pad)
constexpr int kRedPitch = 36;              // [16 rows][32 partials] (+pad)
constexpr int kHalf = 4;                   // packed pairs per thread (8 states)

struct ScanBwdArgs {
    const void *u, *delta, *z, *B, *C, *dout, *out;
    const float *A, *D, *delta_bias, *x;
    void *du, *ddelta, *dz, *out_z;
    float *dA, *dB, *dC, *dD, *ddelta_bias;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, g_bs, g_ds;
    int64_t du_bs, du_ds, dd_bs, dd_ds, dz_bs, dz_ds, oz_bs, oz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, dB_bs, dB_gs, dB_ns, dC_bs, dC_gs, dC_ns;
    int dim, seqlen, dstate, n_groups, n_chunks, softplus, vec_io;
};

struct BwdSmem {
    float u[kRowsB][kPitch];
    float dl[kRowsB][kPitch];
    float dy[kRowsB][kPitch];
    float sig[kRowsB][kPitch];               // d softplus / d raw delta
    float ddl[kRowsB][kPitch];               // outputs
    float du[kRowsB][kPitch];
    float Bs[kSub][kPitch];
    float Cs[kSub][kPitch];
    float ck[kRowsB][kCkPitch];              // state at the start of each mini-chunk
    float red[kThreadsB / 32][2][16][kRedPitch];   // per-warp dB/dC reduction tile, double-buffered by step parity
    float acc[kSub][32 + 1];                 // CTA sums for the chunk: [l][0..15] = dB_n, [l][16..31] = dC_n
};

template <typename T, bool kHasZ>
__global__ void __launch_bounds__(kThreadsB, 3) scan_bwd_kernel(const ScanBwdArgs a) {
    constexpr int VEC = Io<T>::kVec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem &s = *reinterpret_cast<BwdSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int row = tid >> 1, hf = tid & 1;          // row of the slab, which half of the states
    const int rw = lane >> 1;                        // row inside the warp (0..15)
    const int b = blockIdx.y;
    const int dpg = a.dim / a.n_groups;
    const int slabs_per_group = (dpg + kRowsB - 1) / kRowsB;
    const int g = blockIdx.x / slabs_per_group;
    const int d0 = g * dpg + (blockIdx.x % slabs_per_group) * kRowsB;
    const int nrows = min(kRowsB, (g + 1) * dpg - d0);
    const int L = a.seqlen;
    const bool row_ok = row < nrows;
    const int d = d0 + row;
    const int n0 = hf * 2 * kHalf;                   // first state of this thread

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

    // per-thread constants: A (log2e-scaled) for this thread's 4 state pairs
    float2 A2[kHalf];
#pragma unroll
    for (int p = 0; p < kHalf; ++p) {
        const float *Arow = a.A + (int64_t)d * a.A_ds;
        const int n = n0 + 2 * p;
        const float a0 = (row_ok && n < a.dstate) ? Arow[n * a.A_ns] : 0.f;
        const float a1 = (row_ok && n + 1 < a.dstate) ? Arow[(n + 1) * a.A_ns] : 0.f;
        A2[p] = make_float2(a0 * kLog2e, a1 * kLog2e);
    }
    const float Dv = (row_ok && a.D != nullptr) ? a.D[d] : 0.f;
    float2 dA[kHalf], carry[kHalf];            // carry = a_{l+1} g_{l+1} entering the current step from the right
#pragma unroll
    for (int p = 0; p < kHalf; ++p) { dA[p] = make_float2(0.f, 0.f); carry[p] = make_float2(0.f, 0.f); }
    float dD_acc = 0.f, dbias_acc = 0.f;

    // one forward step of this thread's 8 states; returns the decays
    auto fwd_step = [&](int i, float2 (&h)[kHalf], float2 (&dec)[kHalf]) {
        const float dlt = s.dl[row][i];
        const float du_ = dlt * s.u[row][i];
        const float4 B0 = *reinterpret_cast<const float4 *>(&s.Bs[i][n0]);
        const float4 B1 = *reinterpret_cast<const float4 *>(&s.Bs[i][n0 + 4]);
        const float2 Bp[kHalf] = {make_float2(B0.x, B0.y), make_float2(B0.z, B0.w), make_float2(B1.x, B1.y), make_float2(B1.z, B1.w)};
#pragma unroll
        for (int p = 0; p < kHalf; ++p) {
            const float2 t = mul2(splat2(dlt), A2[p]);
            dec[p] = make_float2(ex2_mufu(t.x), ex2_mufu(t.y));
            h[p] = fma2(dec[p], h[p], mul2(splat2(du_), Bp[p]));
        }
    };

    const int n_sub = (L + kSub - 1) / kSub;
    for (int k = n_sub - 1; k >= 0; --k) {
        const int l0 = k * kSub;
        // ------------------------------------------------------------ (a) coalesced loads + elementwise prep
        {
            constexpr int VPR = kSub / VEC;                      // vectors per row chunk
            for (int idx = tid; idx < kRowsB * VPR; idx += kThreadsB) {
                const int r = idx / VPR, v = idx % VPR;
                const int col = v * VEC, l = l0 + col;
                float uv[VEC], dv[VEC], gv[VEC], sg[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) { uv[i] = 0.f; dv[i] = 0.f; gv[i] = 0.f; sg[i] = 0.f; }
                if (r < nrows && l < L) {
                    const bool full = a.vec_io && l + VEC <= L;
                    float zv[VEC], ov[VEC], raw[VEC];
                    if (full) {
                        Io<T>::ldv(u + (int64_t)r * a.u_ds + l, uv);
                        Io<T>::ldv(dl + (int64_t)r * a.dl_ds + l, raw);
                        Io<T>::ldv(go + (int64_t)r * a.g_ds + l, gv);
                        if (kHasZ) {
                            Io<T>::ldv(zz + (int64_t)r * a.z_ds + l, zv);
                            Io<T>::ldv(oo + (int64_t)r * a.o_ds + l, ov);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const bool ok = l + i < L;
                            uv[i] = ok ? Io<T>::ld(u + (int64_t)r * a.u_ds + l + i) : 0.f;
                            raw[i] = ok ? Io<T>::ld(dl + (int64_t)r * a.dl_ds + l + i) : 0.f;
                            gv[i] = ok ? Io<T>::ld(go + (int64_t)r * a.g_ds + l + i) : 0.f;
                            if (kHasZ) {
                                zv[i] = ok ? Io<T>::ld(zz + (int64_t)r * a.z_ds + l + i) : 0.f;
                                ov[i] = ok ? Io<T>::ld(oo + (int64_t)r * a.o_ds + l + i) : 0.f;
                            }
                        }
                    }
                    const float bias = a.delta_bias != nullptr ? a.delta_bias[d0 + r] : 0.f;
                    float dzv[VEC], ozv[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float x = raw[i] + bias;
                        if (a.softplus) {
                            dv[i] = softplus_f(x);
                            sg[i] = x <= 20.f ? sigmoid_f(x) : 1.f;
                        } else {
                            dv[i] = x;
                            sg[i] = 1.f;
                        }
                        if (kHasZ) {
                            const float sz = sigmoid_f(zv[i]);
                            const float silu = zv[i] * sz;
                            dzv[i] = gv[i] * ov[i] * sz * fmaf(zv[i], 1.f - sz, 1.f);   // d/dz [z sigmoid(z)]
                            ozv[i] = ov[i] * silu;
                            gv[i] *= silu;
                        }
                        if (l + i >= L) { dv[i] = 0.f; uv[i] = 0.f; gv[i] = 0.f; sg[i] = 0.f; }
                    }
                    if (kHasZ) {
                        if (full) {
                            Io<T>::stv(dzp + (int64_t)r * a.dz_ds + l, dzv);
                            if (ozp != nullptr) Io<T>::stv(ozp + (int64_t)r * a.oz_ds + l, ozv);
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
#pragma unroll
                for (int i = 0; i < VEC; i += 4) {
                    *reinterpret_cast<float4 *>(&s.u[r][col + i]) = make_float4(uv[i], uv[i + 1], uv[i + 2], uv[i + 3]);
                    *reinterpret_cast<float4 *>(&s.dl[r][col + i]) = make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]);
                    *reinterpret_cast<float4 *>(&s.dy[r][col + i]) = make_float4(gv[i], gv[i + 1], gv[i + 2], gv[i + 3]);
                    *reinterpret_cast<float4 *>(&s.sig[r][col + i]) = make_float4(sg[i], sg[i + 1], sg[i + 2], sg[i + 3]);
                }
            }
            for (int idx = tid; idx < kSub * 16; idx += kThreadsB) {     // B / C tile [l][n]
                const int l = idx % kSub, n = idx / kSub;
                const bool ok = n < a.dstate && l0 + l < L;
                s.Bs[l][n] = ok ? Io<T>::ld(Bg + n * a.B_ns + l0 + l) : 0.f;
                s.Cs[l][n] = ok ? Io<T>::ld(Cg + n * a.C_ns + l0 + l) : 0.f;
            }
            for (int idx = tid; idx < kSub * 32; idx += kThreadsB) s.acc[idx / 32][idx % 32] = 0.f;
        }
        __syncthreads();

        // ------------------------------------------------------------ (b) one forward pass: mini-chunk start states
        float2 hist[kMini + 1][kHalf], dec[kMini][kHalf];
#pragma unroll
        for (int p = 0; p < kHalf; ++p) hist[0][p] = make_float2(0.f, 0.f);
        if (k > 0 && row_ok) {                                  // state after l0 steps: planar checkpoint record
            const int ck = (l0 - 1) / 32, half = (l0 % 32 == 16) ? 0 : 1;
            const float *xp = a.x + (((int64_t)b * a.dim + d) * a.n_chunks + ck) * (2 * a.dstate) + half * a.dstate;
#pragma unroll
            for (int p = 0; p < kHalf; ++p) {
                const int n = n0 + 2 * p;
                if (n < a.dstate) hist[0][p].x = xp[n];
                if (n + 1 < a.dstate) hist[0][p].y = xp[n + 1];
            }
        }
        {
            float2 h[kHalf], dtmp[kHalf];
#pragma unroll
            for (int p = 0; p < kHalf; ++p) h[p] = hist[0][p];
#pragma unroll
            for (int m = 0; m < kSub / kMini - 1; ++m) {          // mini-chunks 0..2: only their end state is kept
                float *ckp = &s.ck[row][m * 16 + n0];
                *reinterpret_cast<float4 *>(ckp) = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
                *reinterpret_cast<float4 *>(ckp + 4) = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
#pragma unroll
                for (int i = 0; i < kMini; ++i) fwd_step(m * kMini + i, h, dtmp);
            }
            float *ckp = &s.ck[row][3 * 16 + n0];
            *reinterpret_cast<float4 *>(ckp) = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
            *reinterpret_cast<float4 *>(ckp + 4) = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
        }

        // ------------------------------------------------------------ (c) mini-chunks in reverse
#pragma unroll 1
        for (int m = kSub / kMini - 1; m >= 0; --m) {
            {   // recompute the 4 steps, keeping h and a
                const float *ckp = &s.ck[row][m * 16 + n0];
                const float4 c0 = *reinterpret_cast<const float4 *>(ckp), c1 = *reinterpret_cast<const float4 *>(ckp + 4);
                hist[0][0] = make_float2(c0.x, c0.y); hist[0][1] = make_float2(c0.z, c0.w);
                hist[0][2] = make_float2(c1.x, c1.y); hist[0][3] = make_float2(c1.z, c1.w);
#pragma unroll
                for (int i = 0; i < kMini; ++i) {
#pragma unroll
                    for (int p = 0; p < kHalf; ++p) hist[i + 1][p] = hist[i][p];
                    fwd_step(m * kMini + i, hist[i + 1], dec[i]);
                }
            }
#pragma unroll
            for (int i = kMini - 1; i >= 0; --i) {
                const int li = m * kMini + i;
                const float dlt = s.dl[row][li], uv = s.u[row][li], dyv = s.dy[row][li];
                const float4 B0 = *reinterpret_cast<const float4 *>(&s.Bs[li][n0]);
                const float4 B1 = *reinterpret_cast<const float4 *>(&s.Bs[li][n0 + 4]);
                const float4 C0 = *reinterpret_cast<const float4 *>(&s.Cs[li][n0]);
                const float4 C1 = *reinterpret_cast<const float4 *>(&s.Cs[li][n0 + 4]);
                const float2 Bp[kHalf] = {make_float2(B0.x, B0.y), make_float2(B0.z, B0.w), make_float2(B1.x, B1.y), make_float2(B1.z, B1.w)};
                const float2 Cp[kHalf] = {make_float2(C0.x, C0.y), make_float2(C0.z, C0.w), make_float2(C1.x, C1.y), make_float2(C1.z, C1.w)};
                float2 accE = make_float2(0.f, 0.f), accG = make_float2(0.f, 0.f);
                float2 dBp[kHalf], dCp[kHalf];
                const float du_ = dlt * uv;
#pragma unroll
                for (int p = 0; p < kHalf; ++p) {
                    const float2 gl = fma2(Cp[p], splat2(dyv), carry[p]);
                    carry[p] = mul2(dec[i][p], gl);
                    const float2 e = mul2(carry[p], hist[i][p]);            // g a h_{l-1}
                    dA[p] = fma2(e, splat2(dlt), dA[p]);
                    accE = fma2(e, A2[p], accE);
                    accG = fma2(gl, Bp[p], accG);
                    dBp[p] = mul2(gl, splat2(du_));
                    dCp[p] = mul2(hist[i + 1][p], splat2(dyv));
                }
                // the two halves of the row meet: ddelta = sum_n e A + u sum_n g B ; du = delta sum_n g B + D dy
                float E = (accE.x + accE.y) * kLn2, G = accG.x + accG.y;     // A = A2 * ln 2
                E += __shfl_xor_sync(0xffffffffu, E, 1);
                G += __shfl_xor_sync(0xffffffffu, G, 1);
                if (hf == 0) {
                    const float ddraw = fmaf(uv, G, E) * s.sig[row][li];
                    s.ddl[row][li] = ddraw;
                    dbias_acc += ddraw;
                    dD_acc = fmaf(dyv, uv, dD_acc);
                } else {
                    s.du[row][li] = fmaf(dlt, G, Dv * dyv);
                }
                // dB / dC of the warp's 16 rows: private tile (double-buffered by step parity, so one __syncwarp per step
                // and the read-back latency overlaps the next step's arithmetic), column sums, two shuffle rounds, CTA sums
                float(*red)[kRedPitch] = s.red[warp][li & 1];
                float *rp = &red[rw][n0];
                *reinterpret_cast<float4 *>(rp) = make_float4(dBp[0].x, dBp[0].y, dBp[1].x, dBp[1].y);
                *reinterpret_cast<float4 *>(rp + 4) = make_float4(dBp[2].x, dBp[2].y, dBp[3].x, dBp[3].y);
                *reinterpret_cast<float4 *>(rp + 16) = make_float4(dCp[0].x, dCp[0].y, dCp[1].x, dCp[1].y);
                *reinterpret_cast<float4 *>(rp + 20) = make_float4(dCp[2].x, dCp[2].y, dCp[3].x, dCp[3].y);
                __syncwarp();
                {
                    const int rg = lane >> 3, cg = lane & 7;                 // rows 4 rg .. 4 rg + 3, columns 4 cg .. 4 cg + 3
                    const float4 w0 = *reinterpret_cast<const float4 *>(&red[4 * rg][4 * cg]);
                    const float4 w1 = *reinterpret_cast<const float4 *>(&red[4 * rg + 1][4 * cg]);
                    const float4 w2 = *reinterpret_cast<const float4 *>(&red[4 * rg + 2][4 * cg]);
                    const float4 w3 = *reinterpret_cast<const float4 *>(&red[4 * rg + 3][4 * cg]);
                    float4 v = make_float4((w0.x + w1.x) + (w2.x + w3.x), (w0.y + w1.y) + (w2.y + w3.y),
                                           (w0.z + w1.z) + (w2.z + w3.z), (w0.w + w1.w) + (w2.w + w3.w));
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
                        v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
                    }
                    if (rg == 0) {
                        atomicAdd(&s.acc[li][4 * cg], v.x);
                        atomicAdd(&s.acc[li][4 * cg + 1], v.y);
                        atomicAdd(&s.acc[li][4 * cg + 2], v.z);
                        atomicAdd(&s.acc[li][4 * cg + 3], v.w);
                    }
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ (d) coalesced stores + dB/dC flush
        {
            constexpr int VPR = kSub / VEC;
            for (int idx = tid; idx < kRowsB * VPR; idx += kThreadsB) {
                const int r = idx / VPR, v = idx % VPR;
                const int col = v * VEC, l = l0 + col;
                if (r < nrows && l < L) {
                    float dd[VEC], du_[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { dd[i] = s.ddl[r][col + i]; du_[i] = s.du[r][col + i]; }
                    if (a.vec_io && l + VEC <= L) {
                        Io<T>::stv(ddp + (int64_t)r * a.dd_ds + l, dd);
                        Io<T>::stv(dup + (int64_t)r * a.du_ds + l, du_);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            if (l + i < L) {
                                Io<T>::st(ddp + (int64_t)r * a.dd_ds + l + i, dd[i]);
                                Io<T>::st(dup + (int64_t)r * a.du_ds + l + i, du_[i]);
                            }
                        }
                    }
                }
            }
            for (int idx = tid; idx < kSub * 32; idx += kThreadsB) {
                const int l = idx % kSub, v = idx / kSub;          // l fastest: consecutive lanes hit consecutive addresses
                const int n = v & 15;
                if (n < a.dstate && l0 + l < L) {
                    float *dst = v < 16 ? a.dB + b * a.dB_bs + g * a.dB_gs + n * a.dB_ns + l0 + l
                                        : a.dC + b * a.dC_bs + g * a.dC_gs + n * a.dC_ns + l0 + l;
                    atomicAdd(dst, s.acc[l][v]);
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------ per-row parameter gradients
    if (row_ok) {
#pragma unroll
        for (int p = 0; p < kHalf; ++p) {
            const int n = n0 + 2 * p;
            if (n < a.dstate) atomicAdd(a.dA + (int64_t)d * a.dstate + n, dA[p].x);
            if (n + 1 < a.dstate) atomicAdd(a.dA + (int64_t)d * a.dstate + n + 1, dA[p].y);
        }
        if (hf == 0) {
            if (a.dD != nullptr) atomicAdd(a.dD + d, dD_acc);
            if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, dbias_acc);
        }
    }
}

template <typename T>
int run(const ScanBwdArgs &a, int batch, cudaStream_t stream) {
    const int dpg = a.dim / a.n_groups;
    dim3 grid(a.n_groups * ((dpg + kRowsB - 1) / kRowsB), batch);
    const int smem = (int)sizeof(BwdSmem);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<grid, kThreadsB, smem, stream>>>(a);
    };
    if (a.z != nullptr) go(scan_bwd_kernel<T, true>); else go(scan_bwd_kernel<T, false>);
    return check_launch("selective_scan_bwd");
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

    const int vec = p->io_dtype == DIMSUM_F32 ? 4 : 8;
    auto ok = [&](const void *ptr, int64_t bs, int64_t ds) { return ptr == nullptr || (aligned16(ptr) && bs % vec == 0 && ds % vec == 0); };
    a.vec_io = ok(p->u, a.u_bs, a.u_ds) && ok(p->delta, a.dl_bs, a.dl_ds) && ok(p->dout, a.g_bs, a.g_ds) &&
               ok(p->z, a.z_bs, a.z_ds) && ok(p->out, a.o_bs, a.o_ds) && ok(p->dz, a.dz_bs, a.dz_ds) &&
               ok(p->out_z_recompute, a.oz_bs, a.oz_ds) && ok(p->du, a.du_bs, a.du_ds) && ok(p->ddelta, a.dd_bs, a.dd_ds);
    switch (p->io_dtype) {
        case DIMSUM_F32: return run<float>(a, (int)p->batch, stream);
        case DIMSUM_BF16: return run<__nv_bfloat16>(a, (int)p->batch, stream);
        default: return run<__half>(a, (int)p->batch, stream);
    }
}
