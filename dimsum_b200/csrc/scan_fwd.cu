// Selective-scan forward for sm_100a.
//
// Replaces selective_scan_fwd_kernel (mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303).
// Design (B200-first, not a port):
//   * one CTA = one batch row x a slab of 128 channels; one THREAD owns one channel row and walks the sequence with all
//     <=16 states in registers as 8 packed f32x2 pairs (FFMA2 / FMUL2) -- no cross-lane scan, so 4 FP ops + 1 exp per
//     state-step instead of the 6 of a lane-split scan;
//   * the sequence is cut into chunks of 64 bytes per row (16 fp32 / 32 bf16 steps).  The raw u / delta / z tiles of
//     chunk c+1 stream into shared memory with cp.async (16-byte, fully coalesced, no registers) WHILE chunk c is being
//     scanned: a 2-stage software pipeline, so HBM latency is hidden behind the exp-bound recurrence even at 12 warps/SM;
//   * tiles keep the storage dtype; their 64-byte rows are XOR-swizzled in 16-byte chunks, which makes the per-thread
//     128-bit row reads bank-conflict free without padding; bias + softplus, the D skip and the silu(z) gate are applied by the scanning thread itself, which overwrites
//     its u slot with the gated output, so the epilogue is a pure coalesced 128-bit copy-out;
//   * the B / C tile is fetched one chunk ahead through registers and parked transposed as [l][n] fp32 (pitch 20), read
//     as warp-wide 128-bit broadcasts: B and C are fetched once per 128 channels, not once per channel as in the reference;
//   * at 16 SFU lanes/clk/SM the exp unit -- not HBM -- is the first limiter of this kernel (20 MUFU results per element).
//     When the caller vouches that every row of A is an arithmetic progression, A[d][n] = (n+1) A[d][0] -- true for the
//     S4D-real initialisation A = -(1..N) of mamba_simple.py:514-521 and checked on the host -- the 16 decays of a step
//     are the powers r, r^2, .. r^16 of ONE exp (kArith): 1 MUFU + 8 packed multiplies instead of 16 MUFU.  Moving
//     part of the exps to a polynomial on the FMA pipe was measured and is slower (profiles/r1_scan_fwd_experiments.md).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace dimsum {
namespace {

// threads per CTA == channel rows per CTA: 128, or 64 for single-wave grids that balance better in finer pieces
constexpr int kNS = 16;             // padded state count
constexpr int kBCPitch = 20;        // words: 16 states + pad -> conflict-free transposed STS.128, aligned LDS.128
constexpr int kRowBytes = 64;       // payload bytes per tile row per chunk
constexpr int kRowPitch = 64;       // bytes: no padding; the four 16-byte chunks of a row are XOR-swizzled instead (swz)

struct ScanFwdArgs {
    const void *u, *delta, *z, *B, *C;
    const float *A, *D, *delta_bias;
    void *out, *out_z;
    float *x;
    const int32_t *perm;
    int64_t u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, out_bs, out_ds, oz_bs, oz_ds;
    int64_t A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    int dim, seqlen, dstate, n_groups, n_chunks, softplus, vec_io;
};

template <int LC, int kRows>
struct ScanSmem {
    unsigned char u[2][kRows][kRowPitch];
    unsigned char dl[2][kRows][kRowPitch];
    unsigned char z[2][kRows][kRowPitch];
    float Bs[2][LC][kBCPitch];
    float Cs[2][LC][kBCPitch];
};

// Tile rows are 64 bytes = four 16-byte chunks.  Chunk c of row r lives at chunk position c ^ ((r >> 1) & 3): with one
// row per lane, the 8 lanes of a quarter-warp then touch 8 distinct 16-byte bank groups (conflict-free LDS/STS.128),
// and the row-major fill / copy-out pattern (4 lanes per row) stays conflict-free too -- with 20 % less shared memory
// than an 80-byte padded pitch, which is what lets 4 CTAs share an SM.
DEV int swz(int r, int chunk) { return (chunk ^ ((r >> 1) & 3)) * 16; }
template <typename T>
DEV T *tile_elem(unsigned char *row_base, int r, int col) {      // address of element `col` of tile row r
    const int byte = col * (int)sizeof(T);
    return reinterpret_cast<T *>(row_base + swz(r, byte >> 4) + (byte & 15));
}

DEV void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Checkpoints for the backward: record c of x[b][d][c][0 .. 2N) holds, PLANAR, h after step 32c+16 in [0, N) and h after
// step 32c+32 in [N, 2N) -- 64 contiguous bytes per store, i.e. full sectors.  (half = 0 / 1.)
DEV void store_state(const ScanFwdArgs &a, int b, int d, int chunk, int half, const float2 (&h2)[kNS / 2]) {
    float *xp = a.x + (((int64_t)b * a.dim + d) * a.n_chunks + chunk) * (2 * a.dstate) + half * a.dstate;
    if (a.dstate == kNS) {
#pragma unroll
        for (int p = 0; p < kNS / 2; p += 2)
            *reinterpret_cast<float4 *>(xp + 2 * p) = make_float4(h2[p].x, h2[p].y, h2[p + 1].x, h2[p + 1].y);
    } else {
#pragma unroll
        for (int p = 0; p < kNS / 2; ++p) {
            if (2 * p < a.dstate) xp[2 * p] = h2[p].x;
            if (2 * p + 1 < a.dstate) xp[2 * p + 1] = h2[p].y;
        }
    }
}
// The last record (index n_chunks - 1) keeps the reference's interleaved convention so that
// last_state = x[:, :, -1, 1::2] (selective_scan_interface.py:39) still reads the final state.
DEV void store_last_state(const ScanFwdArgs &a, int b, int d, const float2 (&h2)[kNS / 2]) {
    float *xp = a.x + (((int64_t)b * a.dim + d) * a.n_chunks + (a.n_chunks - 1)) * (2 * a.dstate);
#pragma unroll
    for (int p = 0; p < kNS / 2; ++p) {
        if (2 * p < a.dstate) { xp[4 * p] = 0.f; xp[4 * p + 1] = h2[p].x; }
        if (2 * p + 1 < a.dstate) { xp[4 * p + 2] = 0.f; xp[4 * p + 3] = h2[p].y; }
    }
}

// (ex2.approx.f16x2 decays for 16-bit I/O were measured and rejected: two MUFU.EX2.F16 per packed op on sm_100a, slower, and
// 3e-2 off on the reference test distribution -- profiles/r2_scan_fwd_experiments.md.)
template <typename T, bool kHasZ, bool kSoftplus, bool kArith, int kRows>
__global__ void __launch_bounds__(kRows, (sizeof(T) == 4 ? 4 : 3) * (128 / kRows)) scan_fwd_kernel(const ScanFwdArgs a) {
    constexpr int VEC = Io<T>::kVec;                 // elements per 16 bytes
    constexpr int LC = kRowBytes / (int)sizeof(T);   // steps per chunk: 16 (fp32) / 32 (16-bit)
    constexpr int VPR = kRowBytes / 16;              // 16-byte vectors per tile row = 4
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem<LC, kRows> &s = *reinterpret_cast<ScanSmem<LC, kRows> *>(smem_raw);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int dpg = a.dim / a.n_groups;
    const int slabs_per_group = (dpg + kRows - 1) / kRows;
    const int g = blockIdx.x / slabs_per_group;
    const int d0 = g * dpg + (blockIdx.x % slabs_per_group) * kRows;
    const int nrows = min(kRows, (g + 1) * dpg - d0);
    const int L = a.seqlen;
    const bool row_ok = tid < nrows;
    const bool gather = a.perm != nullptr;

    const T *u = reinterpret_cast<const T *>(a.u) + b * a.u_bs + (int64_t)d0 * a.u_ds;
    const T *dl = reinterpret_cast<const T *>(a.delta) + b * a.dl_bs + (int64_t)d0 * a.dl_ds;
    const T *z = kHasZ ? reinterpret_cast<const T *>(a.z) + b * a.z_bs + (int64_t)d0 * a.z_ds : nullptr;
    const T *Bg = reinterpret_cast<const T *>(a.B) + b * a.B_bs + g * a.B_gs;
    const T *Cg = reinterpret_cast<const T *>(a.C) + b * a.C_bs + g * a.C_gs;
    T *out = a.out != nullptr ? reinterpret_cast<T *>(a.out) + b * a.out_bs + (int64_t)d0 * a.out_ds : nullptr;
    T *oz = kHasZ ? reinterpret_cast<T *>(a.out_z) + b * a.oz_bs + (int64_t)d0 * a.oz_ds : nullptr;

    // per-row constants of the scanning thread
    float2 A2[kNS / 2];
    float Dv = 0.f, bias = 0.f;
    {
        const float *Arow = a.A + (int64_t)(d0 + tid) * a.A_ds;
#pragma unroll
        for (int p = 0; p < kNS / 2; ++p) {
            const float a0 = (row_ok && 2 * p < a.dstate) ? Arow[(2 * p) * a.A_ns] : 0.f;
            const float a1 = (row_ok && 2 * p + 1 < a.dstate) ? Arow[(2 * p + 1) * a.A_ns] : 0.f;
            A2[p] = make_float2(a0 * kLog2e, a1 * kLog2e);
        }
        if (row_ok && a.D != nullptr) Dv = a.D[d0 + tid];
        if (row_ok && a.delta_bias != nullptr) bias = a.delta_bias[d0 + tid];
    }
    float2 h2[kNS / 2];
#pragma unroll
    for (int p = 0; p < kNS / 2; ++p) h2[p] = make_float2(0.f, 0.f);

    const int n_lc = (L + LC - 1) / LC;

    // ---- producers ---------------------------------------------------------------------------------------------
    // tile fill for chunk c into stage st: 16-byte cp.async when everything is aligned, scalar copies otherwise
    auto fill_tiles = [&](int c, int st) {
        const int l0 = c * LC;
        if (a.vec_io) {
#pragma unroll
            for (int i = 0; i < VPR; ++i) {
                const int idx = tid + i * kRows;          // 0 .. kRows*VPR
                const int r = idx / VPR, v = idx % VPR;
                const int col = l0 + v * VEC;
                if (r < nrows && col < L) {
                    cp_async16(&s.u[st][r][swz(r, v)], u + (int64_t)r * a.u_ds + col);
                    cp_async16(&s.dl[st][r][swz(r, v)], dl + (int64_t)r * a.dl_ds + col);
                    if (kHasZ && !gather) cp_async16(&s.z[st][r][swz(r, v)], z + (int64_t)r * a.z_ds + col);
                }
            }
            if (kHasZ && gather) {
                for (int idx = tid; idx < kRows * LC; idx += kRows) {
                    const int r = idx / LC, col = idx % LC;
                    if (r < nrows && l0 + col < L)
                        *tile_elem<T>(&s.z[st][r][0], r, col) = z[(int64_t)r * a.z_ds + a.perm[l0 + col]];
                }
            }
        } else {
            for (int idx = tid; idx < kRows * LC; idx += kRows) {
                const int r = idx / LC, col = idx % LC;
                if (r < nrows && l0 + col < L) {
                    *tile_elem<T>(&s.u[st][r][0], r, col) = u[(int64_t)r * a.u_ds + l0 + col];
                    *tile_elem<T>(&s.dl[st][r][0], r, col) = dl[(int64_t)r * a.dl_ds + l0 + col];
                    if (kHasZ) {
                        const int tok = gather ? a.perm[l0 + col] : l0 + col;
                        *tile_elem<T>(&s.z[st][r][0], r, col) = z[(int64_t)r * a.z_ds + tok];
                    }
                }
            }
        }
        cp_async_commit();
    };
    // B / C chunk through registers: item = (l = item % LC, state quad = item / LC), LC * 4 items over kRows threads
    constexpr int kBCItems = LC * (kNS / 4);
    constexpr int kBCIter = (kBCItems + kRows - 1) / kRows;
    float bq[kBCIter][4], cq[kBCIter][4];
    auto load_bc = [&](int c) {
#pragma unroll
        for (int it = 0; it < kBCIter; ++it) {
            const int item = tid + it * kRows;
            const int l = c * LC + item % LC, q = item / LC;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = q * 4 + i;
                const bool ok = item < kBCItems && n < a.dstate && l < L;
                bq[it][i] = ok ? Io<T>::ld(Bg + n * a.B_ns + l) : 0.f;
                cq[it][i] = ok ? Io<T>::ld(Cg + n * a.C_ns + l) : 0.f;
            }
        }
    };
    auto store_bc = [&](int st) {
#pragma unroll
        for (int it = 0; it < kBCIter; ++it) {
            const int item = tid + it * kRows;
            if (item < kBCItems) {
                *reinterpret_cast<float4 *>(&s.Bs[st][item % LC][(item / LC) * 4]) = make_float4(bq[it][0], bq[it][1], bq[it][2], bq[it][3]);
                *reinterpret_cast<float4 *>(&s.Cs[st][item % LC][(item / LC) * 4]) = make_float4(cq[it][0], cq[it][1], cq[it][2], cq[it][3]);
            }
        }
    };

    fill_tiles(0, 0);
    load_bc(0);
    store_bc(0);

    for (int c = 0; c < n_lc; ++c) {
        const int st = c & 1;
        const int l0 = c * LC;
        cp_async_wait_all();
        __syncthreads();                                   // chunk c landed for everybody; epilogue(c-1) finished
        const bool more = c + 1 < n_lc;
        if (more) {
            fill_tiles(c + 1, st ^ 1);                     // streams in behind the recurrence below
            load_bc(c + 1);
        }

        // ------------------------------------------------------------------------------ recurrence, thread == row
        // One branch-free basic block per 16-byte group of steps, so ptxas can interleave the independent softplus /
        // exp / FMA chains of the VEC steps; the ragged tail (l >= L) takes a masked copy of the same body.
        auto group = [&](int j, auto masked) {
            constexpr bool kMask = decltype(masked)::value;
            float dv[VEC], uv[VEC], zv[VEC], yv[VEC], gv[VEC];
            Io<T>::ldv(tile_elem<T>(&s.dl[st][tid][0], tid, j), dv);
            Io<T>::ldv(tile_elem<T>(&s.u[st][tid][0], tid, j), uv);
            if (kHasZ) Io<T>::ldv(tile_elem<T>(&s.z[st][tid][0], tid, j), zv);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                dv[k] += bias;
                if (kSoftplus) dv[k] = softplus_t<sizeof(T) == 2>(dv[k]);
                if (kMask && l0 + j + k >= L) {            // identity step past the end keeps the final state exact
                    dv[k] = 0.f;                           // (tile columns past L are never filled: do not trust them)
                    uv[k] = 0.f;
                    if (kHasZ) zv[k] = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float dlt = dv[k];
                const float du = dlt * uv[k];
                float2 dec[kNS / 2];
                if (kArith) {
                    // A[n] = (n + 1) A[0]: decays are r^(n+1) with r = 2^(delta A2[0])
                    const float r = ex2_mufu(dlt * A2[0].x);
                    const float r2 = r * r;
                    dec[0] = make_float2(r, r2);
#pragma unroll
                    for (int p = 1; p < kNS / 2; ++p) dec[p] = mul2(dec[p - 1], splat2(r2));
                } else {
#pragma unroll
                    for (int p = 0; p < kNS / 2; ++p) {
                        const float2 t = mul2(splat2(dlt), A2[p]);
                        dec[p] = make_float2(ex2_mufu(t.x), ex2_mufu(t.y));
                    }
                }
                float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < kNS / 4; ++q) {
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s.Bs[st][j + k][q * 4]);
                    const float4 Cq = *reinterpret_cast<const float4 *>(&s.Cs[st][j + k][q * 4]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = q * 2 + e;
                        const float2 Bp = e ? make_float2(Bq.z, Bq.w) : make_float2(Bq.x, Bq.y);
                        const float2 Cp = e ? make_float2(Cq.z, Cq.w) : make_float2(Cq.x, Cq.y);
                        h2[p] = fma2(dec[p], h2[p], mul2(splat2(du), Bp));
                        if (e) yb = fma2(Cp, h2[p], yb); else ya = fma2(Cp, h2[p], ya);
                    }
                }
                const float2 y2 = add2(ya, yb);
                yv[k] = fmaf(Dv, uv[k], y2.x + y2.y);
                gv[k] = kHasZ ? yv[k] * silu_t<sizeof(T) == 2>(zv[k]) : yv[k];
            }
            // in place: gated output over u, pre-gate output (training only) over delta
            Io<T>::stv(tile_elem<T>(&s.u[st][tid][0], tid, j), gv);
            if (kHasZ && out != nullptr) Io<T>::stv(tile_elem<T>(&s.dl[st][tid][0], tid, j), yv);
        };
        const bool tail = l0 + LC > L;
#pragma unroll 1
        for (int half = 0; half < LC / 16; ++half) {
            if (!tail) {
#pragma unroll 1
                for (int j = half * 16; j < half * 16 + 16; j += VEC) group(j, std::false_type{});
            } else {
#pragma unroll 1
                for (int j = half * 16; j < half * 16 + 16; j += VEC) group(j, std::true_type{});
            }
            if (a.x != nullptr && row_ok) {                // 16-step checkpoints for the backward
                const int done = l0 + half * 16 + 16;
                store_state(a, b, d0 + tid, (done - 1) / 32, (done % 32 == 0) ? 1 : 0, h2);
            }
        }
        if (more) store_bc(st ^ 1);
        __syncthreads();                                   // outputs of chunk c visible

        // ------------------------------------------------------------------------------ copy-out
        // no z: the (ungated) result goes to `out`; with z: gated -> out_z and, when training, pre-gate -> out
        T *dst_main = kHasZ ? oz : out;
        const int64_t main_ds = kHasZ ? a.oz_ds : a.out_ds;
        if (a.vec_io && !gather) {
#pragma unroll
            for (int i = 0; i < VPR; ++i) {
                const int idx = tid + i * kRows;
                const int r = idx / VPR, v = idx % VPR;
                const int col = l0 + v * VEC;
                if (r < nrows && col < L) {
                    *reinterpret_cast<uint4 *>(dst_main + (int64_t)r * main_ds + col) =
                        *reinterpret_cast<const uint4 *>(&s.u[st][r][swz(r, v)]);
                    if (kHasZ && out != nullptr)
                        *reinterpret_cast<uint4 *>(out + (int64_t)r * a.out_ds + col) =
                            *reinterpret_cast<const uint4 *>(&s.dl[st][r][swz(r, v)]);
                }
            }
        } else {
            for (int idx = tid; idx < kRows * LC; idx += kRows) {
                const int r = idx / LC, col = idx % LC;
                if (r < nrows && l0 + col < L) {
                    const int tok = (gather && kHasZ) ? a.perm[l0 + col] : l0 + col;
                    dst_main[(int64_t)r * main_ds + tok] = *tile_elem<T>(&s.u[st][r][0], r, col);
                    if (kHasZ && out != nullptr)
                        out[(int64_t)r * a.out_ds + l0 + col] = *tile_elem<T>(&s.dl[st][r][0], r, col);
                }
            }
        }
    }
    if (a.x != nullptr && row_ok) store_last_state(a, b, d0 + tid, h2);
}

template <typename T, bool kHasZ, bool kSoftplus, bool kArith, int kRows>
int launch_rows(const ScanFwdArgs &a, int batch, cudaStream_t stream) {
    constexpr int LC = kRowBytes / (int)sizeof(T);
    auto kern = scan_fwd_kernel<T, kHasZ, kSoftplus, kArith, kRows>;
    const int smem = (int)sizeof(ScanSmem<LC, kRows>);
    // per instantiation, one bit per device (the attribute is per device); atomic because autograd calls in from several threads
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const int dpg = a.dim / a.n_groups;
    dim3 grid(a.n_groups * ((dpg + kRows - 1) / kRows), batch);
    kern<<<grid, kRows, smem, stream>>>(a);
    return check_launch("selective_scan_fwd");
}

// Fraction of the machine a grid of n CTAs keeps busy when `per_sm` of them fit on each of `sms` SMs: a single wave is
// limited by the fullest SM, several waves by the last, partial one.
inline double wave_efficiency(long long n, int sms, int per_sm) {
    if (n <= (long long)sms * per_sm) {
        const double per = (double)n / sms;
        return per / (double)((n + sms - 1) / sms);
    }
    const double waves = (double)n / ((double)sms * per_sm);
    return waves / (double)(long long)(waves + 0.999999);
}

template <typename T, bool kHasZ, bool kSoftplus, bool kArith>
int launch(const ScanFwdArgs &a, int batch, cudaStream_t stream) {
    // 128-row CTAs share one B/C tile among 128 channels; 64-row CTAs cost twice the B/C staging per channel but cut a
    // single-wave grid into finer pieces (64 latents x 1024 channels: 512 CTAs leave SMs with 3 or 4 -> 86 % busy; 1024
    // half-height CTAs leave them with 6 or 7 -> 99 %).  DIMSUM_SCAN_ROWS=128|64 pins the choice.
    static std::atomic<int> sms_cached{0};
    static std::atomic<int> forced_cached{-1};
    int sms = sms_cached.load(std::memory_order_relaxed);
    if (sms == 0) {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        const char *e = getenv("DIMSUM_SCAN_ROWS");
        forced_cached.store(e ? atoi(e) : 0, std::memory_order_relaxed);
        sms = v > 0 ? v : 148;
        sms_cached.store(sms, std::memory_order_relaxed);
    }
    const int forced = forced_cached.load(std::memory_order_relaxed);
    const int dpg = a.dim / a.n_groups;
    const int per128 = sizeof(T) == 4 ? 4 : 3;
    const long long n128 = (long long)a.n_groups * ((dpg + 127) / 128) * batch, n64 = (long long)a.n_groups * ((dpg + 63) / 64) * batch;
    bool half = wave_efficiency(n64, sms, 2 * per128) > wave_efficiency(n128, sms, per128) + 0.04;
    if (forced == 64) half = true;
    if (forced == 128) half = false;
    return half ? launch_rows<T, kHasZ, kSoftplus, kArith, 64>(a, batch, stream)
                : launch_rows<T, kHasZ, kSoftplus, kArith, 128>(a, batch, stream);
}

template <typename T, bool kArith>
int dispatch2(const ScanFwdArgs &a, int batch, cudaStream_t stream) {
    if (a.z != nullptr) {
        return a.softplus ? launch<T, true, true, kArith>(a, batch, stream) : launch<T, true, false, kArith>(a, batch, stream);
    }
    return a.softplus ? launch<T, false, true, kArith>(a, batch, stream) : launch<T, false, false, kArith>(a, batch, stream);
}

template <typename T>
int dispatch(const ScanFwdArgs &a, int batch, bool arith, cudaStream_t stream) {
    return arith ? dispatch2<T, true>(a, batch, stream) : dispatch2<T, false>(a, batch, stream);
}

}  // namespace
}  // namespace dimsum

using namespace dimsum;

extern "C" int dimsum_selective_scan_fwd(const dimsum_scan_fwd_params *p, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DIMSUM_REQUIRE(p != nullptr, DIMSUM_ERR_INVALID, "selective_scan_fwd: null params");
    if (p != nullptr && p->batch == 0) return DIMSUM_OK;   // empty tensors may carry null pointers
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
    DIMSUM_REQUIRE(p->io_dtype >= 0 && p->io_dtype <= 2, DIMSUM_ERR_INVALID, "selective_scan_fwd: unknown io_dtype %lld",
                   (long long)p->io_dtype);
    if (p->x != nullptr) {
        DIMSUM_REQUIRE(p->chunk_len == 32, DIMSUM_ERR_INVALID, "selective_scan_fwd: chunk_len must be 32");
        DIMSUM_REQUIRE(p->n_chunks == (p->seqlen + 31) / 32 + 1, DIMSUM_ERR_INVALID,
                       "selective_scan_fwd: n_chunks must be ceil(seqlen / 32) + 1");
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
    a.n_chunks = (int)p->n_chunks;
    a.softplus = p->delta_softplus != 0;

    const int esz = p->io_dtype == DIMSUM_F32 ? 4 : 2;
    const int vec = 16 / esz;
    auto strides_ok = [&](int64_t bs, int64_t ds) { return bs % vec == 0 && ds % vec == 0; };
    bool vec_ok = p->seqlen % vec == 0 && aligned16(p->u) && aligned16(p->delta) && strides_ok(a.u_bs, a.u_ds) &&
                  strides_ok(a.dl_bs, a.dl_ds);
    if (p->out) vec_ok = vec_ok && aligned16(p->out) && strides_ok(a.out_bs, a.out_ds);
    if (p->z) vec_ok = vec_ok && aligned16(p->z) && aligned16(p->out_z) && strides_ok(a.z_bs, a.z_ds) && strides_ok(a.oz_bs, a.oz_ds);
    a.vec_io = vec_ok;

    // the arithmetic-progression shortcut needs all 16 states (padding states have A = 0, which breaks the progression)
    const bool arith = p->a_is_arithmetic != 0 && p->dstate == kNS;
    switch (p->io_dtype) {
        case DIMSUM_F32: return dispatch<float>(a, (int)p->batch, arith, stream);
        case DIMSUM_BF16: return dispatch<__nv_bfloat16>(a, (int)p->batch, arith, stream);
        default: return dispatch<__half>(a, (int)p->batch, arith, stream);
    }
}
