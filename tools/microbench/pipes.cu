// Pipe-throughput probes for the roofline arguments in DESIGN.md: results per clock per SM of MUFU.EX2, FFMA, FFMA2 (packed
// fp32x2), FMUL2 and of the legacy mma.sync TF32 path on this GPU.  Each warp runs long unrolled chains with 8 independent
// accumulators; cycles are read with clock64() inside the kernel, so the numbers do not depend on the SM clock.
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipes tools/microbench/pipes.cu && tools/microbench/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t &>(d))
                 : "l"(reinterpret_cast<uint64_t &>(a)), "l"(reinterpret_cast<uint64_t &>(b)), "l"(reinterpret_cast<uint64_t &>(c)));
    return d;
}

template <int kMode>
__global__ void probe(float *out, long long *cycles, float seed) {
    float a[8];
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; p[i] = make_float2(a[i], a[i] + 0.5f); }
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kMode == 0) a[i] = ex2(a[i]);                                    // MUFU only
            if (kMode == 1) a[i] = fmaf(a[i], 0.999f, 1e-3f);                    // FFMA
            if (kMode == 2) p[i] = fma2(p[i], m, c);                             // FFMA2
            if (kMode == 3) { a[i] = ex2(a[i]); p[i] = fma2(p[i], m, c); }       // 1 MUFU : 1 FFMA2
            if (kMode == 4) { a[i] = ex2(a[i]); p[i] = fma2(p[i], m, c); p[i] = fma2(p[i], c, m); p[i] = fma2(p[i], m, c); p[i] = fma2(p[i], c, m); }   // 1 : 4
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void probe_mma(float *out, long long *cycles) {
    // legacy tensor path: mma.sync.aligned.m16n8k8 tf32, 8 independent accumulator tiles per warp
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    const uint32_t a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 7, a2 = a0 + 11, a3 = a0 + 13, b0 = 0x3f000000u + threadIdx.x, b1 = b0 + 5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
double run(F launch, int blocks, long long *dcyc) {
    launch();
    cudaDeviceSynchronize();
    launch();
    cudaDeviceSynchronize();
    long long *h = new long long[blocks];
    cudaMemcpy(h, dcyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    delete[] h;
    return mx;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, threads = 512, per_sm = 2, blocks = sms * per_sm;   // 32 warps per SM
    float *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    const double warps_per_sm = per_sm * threads / 32.0, n = (double)kIters * 8;
    printf("%s, %d SMs, %d warps per SM, clock-counted\n", prop.name, sms, (int)warps_per_sm);
    double c;
    c = run([&] { probe<0><<<blocks, threads>>>(out, cyc, 0.1f); }, blocks, cyc);
    printf("MUFU.EX2 alone      : %.2f results / clk / SM\n", warps_per_sm * 32 * n / c);
    c = run([&] { probe<1><<<blocks, threads>>>(out, cyc, 0.1f); }, blocks, cyc);
    printf("FFMA alone          : %.2f FMA lanes / clk / SM\n", warps_per_sm * 32 * n / c);
    c = run([&] { probe<2><<<blocks, threads>>>(out, cyc, 0.1f); }, blocks, cyc);
    printf("FFMA2 alone         : %.2f FMA lanes / clk / SM (%.2f instructions / clk / SM)\n", warps_per_sm * 64 * n / c, warps_per_sm * n / c);
    c = run([&] { probe<3><<<blocks, threads>>>(out, cyc, 0.1f); }, blocks, cyc);
    printf("1 MUFU : 1 FFMA2    : %.2f MUFU results + %.2f FMA lanes / clk / SM\n", warps_per_sm * 32 * n / c, warps_per_sm * 64 * n / c);
    c = run([&] { probe<4><<<blocks, threads>>>(out, cyc, 0.1f); }, blocks, cyc);
    printf("1 MUFU : 4 FFMA2    : %.2f MUFU results + %.2f FMA lanes / clk / SM\n", warps_per_sm * 32 * n / c, warps_per_sm * 256 * n / c);
    c = run([&] { probe_mma<<<blocks, threads>>>(out, cyc); }, blocks, cyc);
    printf("mma.sync m16n8k8 tf32: %.1f MACs / clk / SM = %.0f dense TFLOP/s at 1.9 GHz on %d SMs\n", warps_per_sm * n * 1024 / c,
           warps_per_sm * n * 1024 / c * 2 * 1.9e9 * sms / 1e12, sms);
    return 0;
}
