// Microbenchmark 3: the exact UMMA of the Clenshaw step (fp16, A = adjacency in tensor memory, B = SWIZZLE_128B MN-major part
// tile, N = 64, 8 K slices per group) and of the VJP's dW product (SS mode, A MN-major), groups of 8 + commit + wait.
// (derived from umma_probe2.cu)
//   (a) one CTA, two issuing threads in different warps, each 32 UMMAs into its own accumulator columns
//   (b) two CTAs per SM (256 TMEM columns each), one issuer each
// Timing only (operands are zeros).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int ISSUERS, int COUNT, int TCOLS, int MODE>
__global__ void __launch_bounds__(128, 2) probe2(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"((uint32_t)TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tslot;
    {
        uint32_t z = 0u;
        for (int col = TCOLS - 64; col < TCOLS; ++col)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((uint32_t)((tid >> 5) * 32) << 16) + col), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const int w = tid >> 5;
    long long t0 = 0, t1 = 0, t2 = 0;
    if ((tid & 31) == 0 && w < ISSUERS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // MODE 0: TS, B SW128 MN-major distinct slices; 1: same but every UMMA reads slice 0; 2: SS, A MN-major SW128 (dW product); 3: TS, B K-major SW128
        const uint32_t idesc = (1u << 4) | ((MODE == 3 ? 0u : 1u) << 16) | ((MODE == 2 ? 1u : 0u) << 15) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t bdesc = make_desc(smem_u32(smem) + (uint32_t)w * 32768u, 16, 1024, 2);
        const uint64_t adesc = make_desc(smem_u32(smem) + 16384u + (uint32_t)w * 32768u, 16, 1024, 2);
        const uint32_t d = tb + (uint32_t)w * 64u;
        uint32_t phase = 0;
        for (int rep = 0; rep < 5; ++rep) {
            t0 = clock64();
#pragma unroll
            for (int i = 0; i < COUNT; ++i) {
                const uint64_t off = (uint64_t)(((MODE == 1 ? 0 : (i & 7)) * 2048) >> 4);
                if (MODE == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc + off), "l"(bdesc + off), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
                else mma_ts(d, tb + (TCOLS - 64) + (uint32_t)(i & 7) * 8u, bdesc + (MODE == 3 ? (uint64_t)((i & 3) * 2) : off), idesc, i > 0);
            }
            umma_commit(smem_u32(&bar[w]));
            t1 = clock64();
            mbar_wait(smem_u32(&bar[w]), phase);
            phase ^= 1u;
            t2 = clock64();
        }
        if (blockIdx.x < 2) { out[(blockIdx.x * 2 + w) * 2] = t2 - t0; out[(blockIdx.x * 2 + w) * 2 + 1] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"((uint32_t)TCOLS) : "memory");
}

template <int N, int ISSUERS, int COUNT, int TCOLS, int MODE>
static void run(const char* name, int grid) {
    long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    auto kern = probe2<N, ISSUERS, COUNT, TCOLS, MODE>;
    const size_t smem = 66 * 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem);
    kern<<<grid, 128, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8] = {0};
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    cudaFree(d);
    printf("%-36s N%3d issuers %d cnt %2d grid %3d occ %d: cta0 w0 %5lld (issue %5lld) w1 %5lld (issue %5lld) | cta1 w0 %5lld w1 %5lld  %s\n", name, N, ISSUERS, COUNT, grid, occ,
           h[0], h[1], h[2], h[3], h[4], h[6], e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
    if (e != cudaSuccess) exit(1);
}

int main() {
    // grid 1: one CTA on an SM; grid 296: two CTAs per SM
    run<64, 1, 8, 256, 0>("TS B=SW128 MN-major, 8 slices", 1);
    run<64, 1, 8, 256, 0>("TS B=SW128 MN-major, 8 slices", 296);
    run<64, 1, 8, 256, 1>("TS same slice", 1);
    run<64, 1, 8, 256, 1>("TS same slice", 296);
    run<64, 1, 8, 256, 3>("TS B K-major", 1);
    run<64, 1, 8, 256, 3>("TS B K-major", 296);
    run<64, 1, 8, 256, 2>("SS A MN-major (dW)", 1);
    run<64, 1, 8, 256, 2>("SS A MN-major (dW)", 296);
    run<64, 2, 8, 256, 0>("TS two issuers in one CTA", 1);
    run<64, 1, 32, 256, 0>("TS 32 per group", 1);
    run<64, 1, 32, 256, 0>("TS 32 per group", 296);
    return 0;
}
