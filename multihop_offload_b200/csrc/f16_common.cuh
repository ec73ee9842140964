// Device helpers shared by the fp16-part tensor-core kernels (cheb_forward_f16.cu, cheb_mlp_f16.cu): mbarrier / bulk-copy /
// tensor-memory wrappers, UMMA descriptors for fp16 operands, the two-part split, packed fp32 arithmetic.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "mho_common.cuh"

namespace {

constexpr int HF_TILE_BYTES = 128 * 128;  // part tile ([node][h 64 B | l' 64 B]), input staging, output staging

// ---- small PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void sts_u128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld_() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st_() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major / MN-major SWIZZLE_128B shared-memory operand descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, M = 128
__device__ __forceinline__ uint32_t idesc_f16(uint32_t n, uint32_t b_mn_major, uint32_t a_negate) {
    return (1u << 4) | (a_negate << 13) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// two fp32 -> one f16x2 word of the rounded values and one f16x2 word of (rounded - exact): x = h - l' to 2^-22
__device__ __forceinline__ void split2(float y0, float y1, uint32_t& h, uint32_t& l) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y1), "f"(y0));
    float r0, r1;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.f16 %0, lo, %3;\n\tsub.rn.f32.f16 %1, hi, %4;\n\t}" : "=f"(r0), "=f"(r1) : "r"(h), "f"(y0), "f"(y1));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }   // 2^e, -126 <= e <= 127
__device__ __forceinline__ int expo_above(float v) { return (int)((__float_as_uint(v) >> 23) & 0xffu) - 126; }  // v < 2^result (v >= 0, finite)

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2): two accumulator columns per instruction
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// index of the graph of node `node` inside its tile: the number of graph starts <= node among the (sorted) node offsets of
// the graphs that follow the tile's first one (`gb`: 128 staged entries; those beyond the tile are >= its end).  Tiles of
// up to nine graphs take two 16 B loads; more graphs (tiny ones) walk on.
__device__ __forceinline__ int group_of(uint32_t gb, int node, int node_end) {
    const uint4 a = lds_u128(gb), b4 = lds_u128(gb + 16u);
    int g = ((int)a.x <= node) + ((int)a.y <= node) + ((int)a.z <= node) + ((int)a.w <= node) + ((int)b4.x <= node) + ((int)b4.y <= node) +
            ((int)b4.z <= node) + ((int)b4.w <= node);
    if ((int)b4.w < node_end) {   // a ninth graph starts inside the tile
        for (int e = 8; e < 128; ++e) {
            const int bnd = (int)lds_u32(gb + (uint32_t)e * 4u);
            if (bnd > node) break;
            ++g;
        }
    }
    return g;
}

__device__ __forceinline__ void bar_quadrant(int q) {   // the two warps of one TMEM lane quadrant (immediate barrier ids 2..5)
    if (q == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
    else if (q == 1) asm volatile("bar.sync 3, 64;" ::: "memory");
    else if (q == 2) asm volatile("bar.sync 4, 64;" ::: "memory");
    else asm volatile("bar.sync 5, 64;" ::: "memory");
}


}  // namespace
