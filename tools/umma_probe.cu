// Microbenchmark: cost of back-to-back tcgen05.mma instructions of the shapes the dense ChebConv kernel issues.
// One CTA per SM slot, one issuing thread; times n UMMAs + commit -> mbarrier completion with clock64.
// Timing only: operand contents are garbage (descriptors are valid, data is not checked).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}


template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 3) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 2) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 3) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}


template <int KIND, int TS, int M, int N, int BMN, int LAYOUT, int COUNT, int SAMEA, int GROUPS>
__global__ void __launch_bounds__(128, 1) probe_kernel(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tslot;
    {
        uint32_t z = 0u;
        for (int col = 256; col < 512; ++col)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((uint32_t)((tid >> 5) * 32) << 16) + col), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t idesc = 0;
        if (KIND == 0) idesc = (1u << 4) | (1u << 7) | (1u << 10);
        if (KIND == 1) idesc = (2u << 4) | (1u << 7) | (1u << 10);
        if (KIND == 2) idesc = (1u << 4) | (2u << 7) | (2u << 10);
        if (KIND == 3) idesc = (1u << 4) | (0u << 7) | (0u << 10);
        idesc |= ((uint32_t)BMN << 16) | (((uint32_t)N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
        const uint32_t sa = smem_u32(smem);
        const uint64_t adesc = make_desc(sa, 16, 512, (uint32_t)LAYOUT);
        const uint64_t bdesc = make_desc(sa + 16384, BMN ? 8192 : 16, 512, (uint32_t)LAYOUT);
        uint32_t phase = 0;
        long long t_total = 0, t_first = 0;
        for (int rep = 0; rep < 5; ++rep) {
            const long long t0 = clock64();
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
#pragma unroll
                for (int i = 0; i < COUNT; ++i) {
                    const uint32_t aoff = SAMEA ? 0u : (uint32_t)(i & 7);
                    if (TS) mma_ts<KIND>(tb, tb + 384 + aoff * 8, bdesc + (uint64_t)(((i & 7) * 1024) >> 4), idesc, i > 0);
                    else mma_ss<KIND>(tb, adesc + 2 * (aoff & 1), bdesc + (uint64_t)(((i & 7) * 1024) >> 4), idesc, i > 0);
                }
                umma_commit(smem_u32(&bar));
            }
            const long long t1 = clock64();
            for (int g = 0; g < GROUPS; ++g) { mbar_wait(smem_u32(&bar), phase); phase ^= 1u; }
            const long long t2 = clock64();
            if (rep == 4) { t_total = t2 - t0; t_first = t1 - t0; }
        }
        if (blockIdx.x == 0) { out[0] = t_total; out[1] = t_first; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int KIND, int TS, int M, int N, int BMN, int LAYOUT, int COUNT, int SAMEA = 0, int GROUPS = 1>
static void run(const char* name, int grid = 1) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    const size_t smem = 100 * 1024;
    auto kern = probe_kernel<KIND, TS, M, N, BMN, LAYOUT, COUNT, SAMEA, GROUPS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 128, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    const int total = COUNT * GROUPS;
    printf("%-40s M%3d N%3d %s cnt %3d x%d : total %6lld cyc (issue %5lld)  per-UMMA %.1f  %s\n", name, M, N, TS ? "TS" : "SS", COUNT, GROUPS,
           h[0], h[1], (double)h[0] / total, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
    if (e != cudaSuccess) exit(1);
}

#define SW64 4
#define SW128 2
#define SW32 6
template <int KIND, int TS, int M, int BMN, int LAYOUT>
static void sweep_n(const char* name) {
    run<KIND, TS, M, 32, BMN, LAYOUT, 32>(name);
    run<KIND, TS, M, 64, BMN, LAYOUT, 32>(name);
    run<KIND, TS, M, 96, BMN, LAYOUT, 32>(name);
    run<KIND, TS, M, 128, BMN, LAYOUT, 32>(name);
    run<KIND, TS, M, 160, BMN, LAYOUT, 32>(name);
}

int main() {
    run<0, 1, 128, 96, 1, SW64, 1>("f16 TS MN SW64 (current adjacency)");
    run<0, 1, 128, 96, 1, SW64, 8>("f16 TS MN SW64 (current adjacency)");
    run<0, 1, 128, 96, 1, SW64, 64>("f16 TS MN SW64 (current adjacency)");
    run<0, 1, 128, 64, 1, SW64, 1>("f16 TS MN SW64");
    run<0, 1, 128, 64, 1, SW64, 8>("f16 TS MN SW64");
    run<0, 1, 128, 64, 1, SW64, 64>("f16 TS MN SW64");
    sweep_n<0, 1, 128, 1, SW64>("f16 TS MN-major SW64");
    sweep_n<0, 1, 128, 1, SW128>("f16 TS MN-major SW128");
    sweep_n<0, 1, 128, 0, SW128>("f16 TS K-major SW128");
    run<0, 1, 128, 32, 1, SW64, 32, 1>("f16 TS MN SW64 same A slice");
    run<0, 1, 128, 64, 1, SW64, 32, 1>("f16 TS MN SW64 same A slice");
    run<0, 1, 128, 96, 1, SW64, 32, 1>("f16 TS MN SW64 same A slice");
    sweep_n<0, 0, 128, 0, SW64>("f16 SS K-major SW64 (X W shape)");
    sweep_n<0, 0, 128, 0, SW128>("f16 SS K-major SW128");
    run<0, 0, 128, 160, 0, SW64, 6>("f16 SS K-major SW64 X W group of 6");
    run<0, 0, 128, 160, 0, SW64, 12>("f16 SS K-major SW64 X W group of 12");
    sweep_n<0, 1, 64, 1, SW64>("f16 TS M=64 MN SW64");
    sweep_n<1, 1, 128, 1, SW32>("i8 TS MN-major SW32");
    sweep_n<1, 1, 128, 1, SW64>("i8 TS MN-major SW64");
    sweep_n<1, 1, 128, 0, SW128>("i8 TS K-major SW128");
    sweep_n<1, 0, 128, 0, SW128>("i8 SS K-major SW128");
    run<1, 1, 128, 96, 1, SW32, 4>("i8 TS MN SW32 group of 4");
    sweep_n<3, 1, 128, 0, SW128>("f8f6f4 TS K-major SW128");
    sweep_n<3, 1, 128, 1, SW64>("f8f6f4 TS MN-major SW64");
    sweep_n<2, 1, 128, 0, SW128>("tf32 TS K-major SW128");
    run<0, 1, 128, 96, 1, SW64, 8, 0, 8>("f16 TS N=96 8 groups of 8");
    run<0, 1, 128, 64, 1, SW64, 8, 0, 8>("f16 TS N=64 8 groups of 8");
    run<1, 1, 128, 96, 1, SW32, 4, 0, 8>("i8 TS N=96 8 groups of 4");
    run<0, 1, 128, 96, 1, SW64, 64>("f16 TS N=96 all SMs", 148);
    run<0, 1, 128, 64, 1, SW64, 64>("f16 TS N=64 all SMs", 148);
    run<1, 1, 128, 96, 1, SW32, 64>("i8 TS N=96 all SMs", 148);
    return 0;
}
