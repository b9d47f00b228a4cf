// tcgen05 / TMEM / mbarrier building blocks for sm_100a (inline PTX; no CUTLASS).
//
// Conventions, all checked on the hardware by tools/microbench/umma_probe.cu (profiles/r2_umma_probe.md):
//   * operands are K-major tiles in shared memory with the 128-byte swizzle: one row = 128 bytes of K (32 tf32 or 64 bf16
//     values), 8-row groups of 1024 bytes, the 16-byte chunk c of row r stored at chunk position c ^ (r % 8).  Tile bases are
//     1024-byte aligned.  Matrix descriptor: start >> 4, LBO (ignored for this layout) = 1, SBO = 1024 >> 4, version 1,
//     layout type 2 (SWIZZLE_128B).  One instruction consumes 32 bytes of K: K = 8 (tf32) or 16 (bf16 / fp16); the next
//     K step is start address + 32 bytes.
//   * kind::tf32 reads fp32 bit patterns and TRUNCATES them to tf32.  fp32-grade products therefore come from the
//     3xTF32 split D = Ah Bh + Al Bh + Ah Bl with Xh = X as stored and Xl = X - trunc(X) (measured 1e-6 relative).
//     (tf32 operands must be K-major: the MN-major no-swizzle layout that works for bf16 returns garbage for tf32.)
//   * the accumulator of an M = 128 instruction occupies TMEM lanes 0..127 (row m = lane m) and N consecutive columns;
//     warp w of the CTA may read lanes 32 (w % 4) .. +31 with tcgen05.ld.32x32b (thread t gets lane 32 (w % 4) + t).
#pragma once

#include <stdint.h>

namespace dimsum {
namespace umma {

#define UMMA_DEV __device__ __forceinline__

UMMA_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor of a K-major SWIZZLE_128B tile starting at shared address `saddr` (+ 32 bytes per K step)
UMMA_DEV uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a SWIZZLE_128B tile
UMMA_DEV uint32_t sw128_off(int row, int chunk) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4)); }

enum { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2 };
// instruction descriptor: fp32 accumulate, A and B K-major, M x N
__host__ __device__ constexpr uint32_t idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool kTf32>
UMMA_DEV void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc_, uint32_t accumulate) {
    if (kTf32) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc_), "r"(accumulate));
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc_), "r"(accumulate));
    }
}
// completion of all tcgen05 operations issued so far by this thread -> one arrival on the mbarrier
UMMA_DEV void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
UMMA_DEV void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
UMMA_DEV void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core's operand reads)
UMMA_DEV void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

UMMA_DEV void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
UMMA_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra LAB_DONE;\n\tbra LAB_WAIT;\n\tLAB_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// one warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in *slot (shared memory)
UMMA_DEV void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
UMMA_DEV void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 32 lanes x 8 / 16 / 32 consecutive 32-bit columns: thread t of the warp receives lane (taddr.lane + t)
UMMA_DEV void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
UMMA_DEV void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
UMMA_DEV void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
UMMA_DEV void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 columns, registers -> TMEM (e.g. softmax probabilities handed back to the tensor core as operand A)
UMMA_DEV void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
UMMA_DEV void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// fp32 value minus its tf32 truncation: the low part of the 3xTF32 split (exact in fp32)
UMMA_DEV float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

}  // namespace umma
}  // namespace dimsum
