// tcgen05 / TMEM bring-up probe (B200, sm_100a).  Checks, against a CPU fp64 product, the shared-memory matrix
// descriptors (no-swizzle "interleave" layouts, K-major and MN-major), the instruction descriptor and the TMEM read-back
// this repo's fused kernels use, for kind::tf32 and kind::f16 (bf16), plus the 3xTF32 split (hi = raw fp32, the tensor
// core truncates; lo = x - trunc(x)) that gives fp32-grade accuracy.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define DEV __device__ __forceinline__

DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors -------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64)
DEV uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)layout_type << 61;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a_format [7,10), b_format [10,13), a_major [15],
// b_major [16], N>>3 [17,23), M>>4 [24,29)
__host__ __device__ inline uint32_t make_idesc(int fmt, int a_mn_major, int b_mn_major, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
enum { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };

DEV void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
DEV void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
DEV void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEV void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
DEV void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
DEV void ld_tmem_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- test kernel ---------------------------------------------------------------------------------------------------------
struct Cfg {
    int fmt;          // FMT_TF32 / FMT_BF16
    int a_mn, b_mn;   // 1 = MN-major, 0 = K-major
    int M, N, K;      // M = 128, N multiple of 8 (<= 256), K multiple of the per-instruction K
    int pad;          // extra bytes added to the 8-row / MN-group stride (SBO) -- to test padded, bank-conflict-free strides
    int split;        // 1: 3xTF32 (A = Ah + Al, B = Bh + Bl; D = Ah Bh + Al Bh + Ah Bl)
    int sw128;        // 1: both operands K-major with the 128-byte swizzle (rows of 128 B, 8-row groups of 1024 B, 16-byte chunk
                      //    index XOR row % 8; descriptor layout_type 2, SBO 1024); K * element size must be 128
    int lbo_pad;      // extra bytes on the K-chunk stride (LBO) of the no-swizzle K-major layout
};

// byte offset of element (mn, k) inside an operand tile (element size es, kk = K extent of the whole tile, nn = MN extent)
__host__ __device__ inline uint32_t tile_off(int mn_major, int es, int mn, int k, int nn, int kk, int pad, uint32_t *lbo, uint32_t *sbo,
                                             int sw128 = 0, int lbo_pad = 0) {
    const int T = 16 / es;                       // elements per 16 bytes
    if (sw128) {
        if (lbo) { *lbo = 16; *sbo = 1024; }
        const int chunk = (k * es) / 16;
        return (mn / 8) * 1024 + (mn % 8) * 128 + ((chunk ^ (mn % 8)) * 16) + (k * es) % 16;
    }
    if (!mn_major) {                             // K-major: core = 8 rows x 16 B; chunks of a row group are LBO apart, row groups SBO apart
        const uint32_t SBO = 128 + pad, LBO = (nn / 8) * SBO + lbo_pad;
        if (lbo) { *lbo = LBO; *sbo = SBO; }
        return (mn % 8) * 16 + (mn / 8) * SBO + (k / T) * LBO + (k % T) * es;
    }
    // MN-major: core = 8 k-rows x 16 B (T elements along MN); MN groups SBO apart, groups of 8 k LBO apart
    const uint32_t SBO = 128 + pad, LBO = (nn / T) * SBO;
    if (lbo) { *lbo = LBO; *sbo = SBO; }
    return (mn % T) * es + (mn / T) * SBO + (k % 8) * 16 + (k / 8) * LBO;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, const float *A, const float *B, float *D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int es = c.fmt == FMT_TF32 ? 4 : 2;
    const int kper = 32 / es;                               // K per instruction: 8 (tf32) / 16 (bf16)
    uint32_t lboA, sboA, lboB, sboB;
    tile_off(c.a_mn, es, 0, 0, c.M, c.K, c.pad, &lboA, &sboA, c.sw128, c.lbo_pad);
    tile_off(c.b_mn, es, 0, 0, c.N, c.K, c.pad, &lboB, &sboB, c.sw128, c.lbo_pad);
    const uint32_t szA = c.sw128 ? (c.M / 8) * 1024 : tile_off(c.a_mn, es, c.M - 1, c.K - 1, c.M, c.K, c.pad, nullptr, nullptr, 0, c.lbo_pad) + es;
    const uint32_t szB = c.sw128 ? (c.N / 8) * 1024 : tile_off(c.b_mn, es, c.N - 1, c.K - 1, c.N, c.K, c.pad, nullptr, nullptr, 0, c.lbo_pad) + es;
    const uint32_t offAh = 0, offAl = (szA + 1023) & ~1023u, offBh = 2 * offAl, offBl = offBh + ((szB + 1023) & ~1023u);

    auto put = [&](uint32_t base, int mn_major, int mn, int k, int nn, float v, bool lo) {
        const uint32_t o = base + tile_off(mn_major, es, mn, k, nn, c.K, c.pad, nullptr, nullptr, c.sw128, c.lbo_pad);
        if (es == 4) {
            float w = v;
            if (lo) w = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            *reinterpret_cast<float *>(smem + o) = w;
        } else {
            *reinterpret_cast<__nv_bfloat16 *>(smem + o) = __float2bfloat16_rn(v);
        }
    };
    for (int i = tid; i < c.M * c.K; i += 128) {
        put(offAh, c.a_mn, i / c.K, i % c.K, c.M, A[i], false);
        if (c.split) put(offAl, c.a_mn, i / c.K, i % c.K, c.M, A[i], true);
    }
    for (int i = tid; i < c.N * c.K; i += 128) {
        put(offBh, c.b_mn, i / c.K, i % c.K, c.N, B[i], false);
        if (c.split) put(offBl, c.b_mn, i / c.K, i % c.K, c.N, B[i], true);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc(c.fmt, c.a_mn, c.b_mn, c.M, c.N);
        const uint32_t sbase = smem_u32(smem);
        // bytes to advance the start address per instruction along K
        const uint32_t advA = c.sw128 ? 32 : c.a_mn ? (kper / 8) * lboA : (kper * es / 16) * lboA;
        const uint32_t advB = c.sw128 ? 32 : c.b_mn ? (kper / 8) * lboB : (kper * es / 16) * lboB;
        uint32_t acc = 0;
        for (int pass = 0; pass < (c.split ? 3 : 1); ++pass) {
            const uint32_t oa = pass == 1 ? offAl : offAh, ob = pass == 2 ? offBl : offBh;
            for (int k = 0; k < c.K / kper; ++k) {
                const uint64_t da = make_desc(sbase + oa + k * advA, lboA, sboA, c.sw128 ? 2 : 0);
                const uint64_t db = make_desc(sbase + ob + k * advB, lboB, sboB, c.sw128 ? 2 : 0);
                if (c.fmt == FMT_TF32) mma_tf32(tmem, da, db, idesc, acc); else mma_f16(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread t of warp w owns TMEM lane 32 w + t == row m of D
    for (int n0 = 0; n0 < c.N; n0 += 32) {
        uint32_t v[32];
        ld_tmem_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        for (int j = 0; j < 32 && n0 + j < c.N; ++j) D[(size_t)tid * c.N + n0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x7fff + ((u >> 16) & 1);
    u &= 0xffff0000u;
    memcpy(&x, &u, 4);
    return x;
}
static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

static int run(const char *name, Cfg c) {
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K), D((size_t)c.M * c.N, -7.f);
    srand(1234);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_exact = 0, max_model = 0, ref_max = 0;
    for (int m = 0; m < c.M; ++m)
        for (int n = 0; n < c.N; ++n) {
            double exact = 0, model = 0;
            for (int k = 0; k < c.K; ++k) {
                const float a = A[(size_t)m * c.K + k], b = B[(size_t)n * c.K + k];
                exact += (double)a * b;
                if (c.fmt == FMT_TF32) model += (double)tf32_trunc(a) * tf32_trunc(b);
                else model += (double)bf16_round(a) * bf16_round(b);
            }
            const double got = D[(size_t)m * c.N + n];
            max_exact = fmax(max_exact, fabs(got - exact));
            max_model = fmax(max_model, fabs(got - model));
            ref_max = fmax(ref_max, fabs(exact));
        }
    printf("%-44s max|D-exact|/max|D| = %.3e   max|D-operand-rounded model|/max|D| = %.3e\n", name, max_exact / ref_max, max_model / ref_max);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

int main() {
    int bad = 0;
    //                                         fmt      a_mn b_mn  M    N    K  pad split
    bad += run("tf32  A K-major   B K-major  128x64x32", {FMT_TF32, 0, 0, 128, 64, 32, 0, 0});
    bad += run("tf32  A MN-major  B K-major  128x64x32", {FMT_TF32, 1, 0, 128, 64, 32, 0, 0});
    bad += run("tf32  A MN-major  B K-major  pad16     ", {FMT_TF32, 1, 0, 128, 64, 32, 16, 0});
    bad += run("tf32  A K-major   B MN-major 128x64x32", {FMT_TF32, 0, 1, 128, 64, 32, 0, 0});
    bad += run("tf32  A K-major   B MN-major 128x16x32", {FMT_TF32, 0, 1, 128, 16, 32, 0, 0});
    bad += run("tf32  A K-major   B K-major  128x256x64", {FMT_TF32, 0, 0, 128, 256, 64, 0, 0});
    bad += run("tf32  3xTF32 A MN  B K       128x64x64", {FMT_TF32, 1, 0, 128, 64, 64, 0, 1});
    bad += run("tf32  3xTF32 A K   B K       128x256x64", {FMT_TF32, 0, 0, 128, 256, 64, 0, 1});
    bad += run("tf32  3xTF32 A K   B MN      128x64x128", {FMT_TF32, 0, 1, 128, 64, 128, 0, 1});
    bad += run("bf16  A K-major   B K-major  128x64x64", {FMT_BF16, 0, 0, 128, 64, 64, 0, 0});
    bad += run("bf16  A MN-major  B K-major  128x64x64", {FMT_BF16, 1, 0, 128, 64, 64, 0, 0});
    bad += run("bf16  A K-major   B MN-major 128x64x64", {FMT_BF16, 0, 1, 128, 64, 64, 0, 0});
    bad += run("bf16  A K-major   B MN-major 128x32x32", {FMT_BF16, 0, 1, 128, 32, 32, 0, 0});
    // padded strides of the no-swizzle K-major layout (bank-conflict-free producer stores) and the 128-byte swizzle
    bad += run("tf32  K/K  SBO 160 LBO +16    128x64x32", {FMT_TF32, 0, 0, 128, 64, 32, 32, 0, 0, 16});
    bad += run("tf32  K/K  SBO 144            128x64x32", {FMT_TF32, 0, 0, 128, 64, 32, 16, 0, 0, 0});
    bad += run("tf32  3xTF32 K/K SBO 160 LBO+16 128x64x64", {FMT_TF32, 0, 0, 128, 64, 64, 32, 1, 0, 16});
    bad += run("tf32  K/K  SWIZZLE_128B       128x64x32", {FMT_TF32, 0, 0, 128, 64, 32, 0, 0, 1, 0});
    bad += run("tf32  K/K  SWIZZLE_128B       128x256x32", {FMT_TF32, 0, 0, 128, 256, 32, 0, 0, 1, 0});
    bad += run("tf32  3xTF32 K/K SWIZZLE_128B 128x64x32", {FMT_TF32, 0, 0, 128, 64, 32, 0, 1, 1, 0});
    bad += run("bf16  K/K  SWIZZLE_128B       128x64x64", {FMT_BF16, 0, 0, 128, 64, 64, 0, 0, 1, 0});
    bad += run("tf32  K/K  N=16               128x16x32", {FMT_TF32, 0, 0, 128, 16, 32, 0, 0, 0, 0});
    bad += run("tf32  K/K  N=32 3xTF32        128x32x32", {FMT_TF32, 0, 0, 128, 32, 32, 0, 1, 0, 0});
    return bad;
}
