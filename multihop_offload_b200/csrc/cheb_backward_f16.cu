// VJP of ONE ChebConv layer on the tensor cores (sm_100a, tcgen05 + TMEM): the training step of the benchmark
// configuration and of the reference's K > 1 models - 32 -> 32 features, 2 <= K <= 10, BINARY operator (vals == NULL),
// graphs of <= 128 nodes, first layer (no input gradient).  Replaces, for those shapes, the tape replay
//   gradients = g.gradient(delay_mtx_ts, self.model.trainable_weights, output_gradients=grad_dist_np)
// (gnn_offloading_agent.py:448); everything else stays with cheb_backward.cu.  One gradient vector PER GRAPH is written
// (the reference memorises one gradient list per instance, :142/:450, and replays them one by one, :156-169).
//
//   G = dOut (.) act'(Out)        db = sum_i G[i, :]        dW_k = T_k^T G,   T_0 = X, T_1 = A X, T_k = 2 A T_k-1 - T_k-2
//
// One graph per CTA pass, two CTAs per SM; thread = one node row x 16 of the 32 columns (8 warps).
//   * fp32 values travel as TWO fp16 parts (x = h - l', 22 significand bits) after ONE power-of-two scale per graph and per
//     recurrence step (the node rows are the reduction dimension of both products, so a scale may not vary along them):
//     max |X|, max |G| and the maximum degree are reduced across the CTA once per graph; the scales of T_1 .. T_K-1 come
//     from the a-priori bound beta_k = 2 dmax beta_k-1 + beta_k-2 (no further reductions).
//   * the recurrence is the forward kernel's adjacency product: the graph's 128 x 128 adjacency block sits in tensor
//     memory as fp16 0 / 1 (A operand), the part tile [node][h 64 B | l' 64 B] of T_k is the MN-major B operand with
//     N = 64: D = A [h | l'] -> T_k+1 = c (D_h - D_l') - T_k-1 in packed fp32 registers.
//   * dW_k^T = G^T T_k is ONE more UMMA per 16-node slice on the same part tile: A = the part tile of G read MN-major
//     (rows of D: [G_h columns o | G_l' columns o]), B = the part tile of T_k.  The four 32 x 32 blocks of D are the four
//     part products; (hh - hl') - (l'h - l'l') is formed while the accumulator is drained, one step later, under the next
//     step's UMMAs (two 64-column accumulators alternate), and leaves as coalesced 128 B rows of the gradient vector.
//   * db: exact fp32 column sums of G by warp butterflies, added in a fixed order.
// Tensor memory: 64 columns recurrence accumulator | 2 x 64 dW accumulators | 64 adjacency = 256 -> two CTAs per SM.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "mho_common.cuh"
#include "mho_internal.h"
#include "f16_common.cuh"

namespace {

constexpr int BF_THREADS = 256;
constexpr uint32_t BF_TCOLS = 256u, BF_ADJ_COL = 192u, BF_DW_COL = 64u;

struct BfParams {
    const int32_t* graph_off;
    const uint32_t* adj_bits;   // [total_nodes][4], bit j of word w of node i: i ~ graph_node0(i) + 32 w + j
    int n_graphs;
    const float* X;
    const float* Y;
    const float* dY;
    float* grads;               // [n_graphs][n_params]: W[K][32][32] then b[32]
    long long n_params;
    int K;
    int act;
    float slope;
};

// shared memory (1024-aligned): two T part tiles | G part tile | X, dY, Y staging (swizzled 128 B rows) | bit rows | drain
// staging [2][32][32] | control block
constexpr uint32_t BF_PT = 0, BF_PG = 2 * HF_TILE_BYTES, BF_XS = 3 * HF_TILE_BYTES, BF_DS = 4 * HF_TILE_BYTES, BF_YS = 5 * HF_TILE_BYTES,
                   BF_BITS = 6 * HF_TILE_BYTES, BF_STG = BF_BITS + 2048, BF_CTL = BF_STG + 8192, BF_SMEM = BF_CTL + 1024;

__device__ __forceinline__ void bar_drain(int which, bool wait) {   // 128 threads: the four warps that drain a dW accumulator
    if (which == 0) { if (wait) asm volatile("bar.sync 6, 128;" ::: "memory"); else asm volatile("bar.arrive 6, 128;" ::: "memory"); }
    else { if (wait) asm volatile("bar.sync 7, 128;" ::: "memory"); else asm volatile("bar.arrive 7, 128;" ::: "memory"); }
}

template <bool TRACK>
__global__ void __launch_bounds__(BF_THREADS, 2) cheb_backward_f16_kernel(const __grid_constant__ BfParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t smem_a = smem_u32(smem);
    const uint32_t pt_a = smem_a + BF_PT, pg_a = smem_a + BF_PG, xs_a = smem_a + BF_XS, ds_a = smem_a + BF_DS, ys_a = smem_a + BF_YS,
                   bits_a = smem_a + BF_BITS, ctl_a = smem_a + BF_CTL;
    unsigned char* ctl_s = smem + BF_CTL;
    // control block: loads +0, recurrence +8, dW[2] +16, tmem slot +32, reductions +64 ([2][4]), LUT +128, db partials +256 ([4][32]),
    // running maxima of |T_k| +768 ([2][16])
    const uint32_t bar_ld = ctl_a, bar_mma = ctl_a + 8, bar_dw = ctl_a + 16, tslot = ctl_a + 32, lut_a = ctl_a + 128;
    unsigned int* red_s = reinterpret_cast<unsigned int*>(ctl_s + 64);
    float* dbs_s = reinterpret_cast<float*>(ctl_s + 256);
    unsigned int* trk_s = reinterpret_cast<unsigned int*>(ctl_s + 768);
    float* stg_s = reinterpret_cast<float*>(smem + BF_STG);

    const int G = (int)gridDim.x, K = p.K;
    const int n_my = (int)blockIdx.x < p.n_graphs ? (p.n_graphs - (int)blockIdx.x + G - 1) / G : 0;

    // graph extents are fetched one graph ahead
    int nx_node0 = 0, nx_rows = 0;
    auto prefetch_extent = [&](int j) {
        if (j < n_my) {
            const int g = (int)blockIdx.x + j * G;
            nx_node0 = __ldg(p.graph_off + g);
            nx_rows = __ldg(p.graph_off + g + 1) - nx_node0;
        }
    };
    // rows of X, dY, Y (16 B chunks XOR-swizzled with the row: the row-per-lane reads below are conflict-free) and bit rows
    auto issue_loads = [&](int node0, int rows) {   // all threads
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + 256 * i, row = c >> 3, ch = c & 7;
            if (row < rows) {
                const uint32_t d = (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
                const size_t s = (size_t)(node0 + row) * 32 + (size_t)ch * 4;
                cp_async16(xs_a + d, p.X + s);
                cp_async16(ds_a + d, p.dY + s);
                cp_async16(ys_a + d, p.Y + s);
            }
        }
        if (tid < rows) cp_async16(bits_a + (uint32_t)tid * 16u, p.adj_bits + (size_t)(node0 + tid) * 4);
        cp_async_mbar_arrive(bar_ld);
    };

    if (tid == 0) {
        mbar_init(bar_ld, BF_THREADS);
        mbar_init(bar_mma, 1);
        mbar_init(bar_dw, 1);
        mbar_init(bar_dw + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 16) {
        const uint32_t x = ((tid & 1) ? 0x3C00u : 0u) | ((tid & 2) ? 0x3C000000u : 0u), y = ((tid & 4) ? 0x3C00u : 0u) | ((tid & 8) ? 0x3C000000u : 0u);
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(lut_a + (uint32_t)tid * 8u), "r"(x), "r"(y) : "memory");
    }
    if (tid < 8) red_s[tid] = 0u;
    if (tid < 32) trk_s[tid] = 0u;
    if (warp == 0) tmem_alloc(tslot, BF_TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(ctl_s + 32);

    prefetch_extent(0);
    if (n_my > 0) issue_loads(nx_node0, nx_rows);

    const int q = warp & 3, hh = warp >> 2;                 // TMEM lane quadrant, column half
    const uint32_t r = (uint32_t)(q * 32 + lane);           // graph row = TMEM lane
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t key = r & 7u;
    uint32_t ph_mma = 0, ph_dw0 = 0, ph_dw1 = 0;

    // scaled fp32 row -> two fp16 parts into row r of a part tile
    auto split_row = [&](uint32_t tile_a, const float (&v)[16], float scale) {
        const uint64_t S2 = pk2(scale, scale);
        uint32_t h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float y0, y1;
            upk2(mul2(pk2(v[2 * e], v[2 * e + 1]), S2), y0, y1);
            split2(y0, y1, h[e], l[e]);
        }
        const uint32_t prow_a = tile_a + r * 128u;
        sts_u128(prow_a + (((uint32_t)(2 * hh) ^ key) << 4), h[0], h[1], h[2], h[3]);
        sts_u128(prow_a + (((uint32_t)(2 * hh + 1) ^ key) << 4), h[4], h[5], h[6], h[7]);
        sts_u128(prow_a + (((uint32_t)(4 + 2 * hh) ^ key) << 4), l[0], l[1], l[2], l[3]);
        sts_u128(prow_a + (((uint32_t)(5 + 2 * hh) ^ key) << 4), l[4], l[5], l[6], l[7]);
    };

    for (int j = 0; j < n_my; ++j) {
        const int node0 = nx_node0, rows = nx_rows;
        prefetch_extent(j + 1);
        float* gout = p.grads + (size_t)((int)blockIdx.x + j * G) * p.n_params;
        const bool live = (int)r < rows;
        const int nks = (rows + 15) >> 4;   // 16-node slices beyond the graph's rows are all zero

        // ---- input rows: T_0 = X, G = dOut (.) act'(Out); the graph's maxima; adjacency -> tensor memory
        mbar_wait(bar_ld, (uint32_t)(j & 1));
        float tp[16], tpp[16], gr[16];
        float xm = 0.f, gm = 0.f;
        unsigned int deg = 0u;
        uint2 m2v = make_uint2(0u, 0u);
        if (live) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t off = r * 128u + (((uint32_t)(4 * hh + c) ^ key) << 4);
                const float4 x4 = lds_f128(xs_a + off), d4 = lds_f128(ds_a + off), y4 = lds_f128(ys_a + off);
                tp[4 * c] = x4.x; tp[4 * c + 1] = x4.y; tp[4 * c + 2] = x4.z; tp[4 * c + 3] = x4.w;
                gr[4 * c] = d4.x * act_grad_from_out(y4.x, p.act, p.slope);
                gr[4 * c + 1] = d4.y * act_grad_from_out(y4.y, p.act, p.slope);
                gr[4 * c + 2] = d4.z * act_grad_from_out(y4.z, p.act, p.slope);
                gr[4 * c + 3] = d4.w * act_grad_from_out(y4.w, p.act, p.slope);
            }
            const uint4 m4 = lds_u128(bits_a + r * 16u);
            deg = __popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w);
            m2v = hh ? make_uint2(m4.z, m4.w) : make_uint2(m4.x, m4.y);
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) { tp[e] = 0.f; gr[e] = 0.f; }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) { tpp[e] = 0.f; xm = fmaxf(xm, fabsf(tp[e])); gm = fmaxf(gm, fabsf(gr[e])); }
        {
            const unsigned int wx = __reduce_max_sync(0xffffffffu, __float_as_uint(xm));
            const unsigned int wg = __reduce_max_sync(0xffffffffu, __float_as_uint(gm));
            const unsigned int wd = __reduce_max_sync(0xffffffffu, deg);
            unsigned int* red = red_s + (j & 1) * 4;
            if (lane == 0) { atomicMax(red, wx); atomicMax(red + 1, wg); atomicMax(red + 2, wd); }
        }
        // adjacency block -> tensor memory as fp16 0 / 1 pairs (the previous graph's recurrence UMMAs have completed)
#pragma unroll
        for (int w2 = 0; w2 < 2; ++w2) {
            const uint32_t m = w2 ? m2v.y : m2v.x;
            uint32_t aw[16];
#pragma unroll
            for (int b4 = 0; b4 < 8; ++b4) {
                uint2 v;
                const uint32_t idx = b4 == 0 ? ((m << 3) & 0x78u) : ((m >> (4 * b4 - 3)) & 0x78u);
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(lut_a + idx));
                aw[2 * b4] = v.x; aw[2 * b4 + 1] = v.y;
            }
            tmem_st16(tmem_lane + BF_ADJ_COL + (uint32_t)(32 * hh + 16 * w2), aw);
        }
        // db: column sums of this warp's 32 rows (butterfly: 16 -> 8 -> 4 -> 2 -> 1 values per lane), fixed order
        {
            float s8[8], s4[4], s2[2], s1;
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float o = __shfl_xor_sync(0xffffffffu, b16 ? gr[i] : gr[i + 8], 16); s8[i] = (b16 ? gr[i + 8] : gr[i]) + o; }
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float o = __shfl_xor_sync(0xffffffffu, b8 ? s8[i] : s8[i + 4], 8); s4[i] = (b8 ? s8[i + 4] : s8[i]) + o; }
#pragma unroll
            for (int i = 0; i < 2; ++i) { const float o = __shfl_xor_sync(0xffffffffu, b4 ? s4[i] : s4[i + 2], 4); s2[i] = (b4 ? s4[i + 2] : s4[i]) + o; }
            { const float o = __shfl_xor_sync(0xffffffffu, b2 ? s2[0] : s2[1], 2); s1 = (b2 ? s2[1] : s2[0]) + o; }
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            const int col = (b16 ? 8 : 0) + (b8 ? 4 : 0) + (b4 ? 2 : 0) + (b2 ? 1 : 0);
            if (!(lane & 1)) dbs_s[q * 32 + 16 * hh + col] = s1;
        }
        bar_compute();   // #1: the maxima are complete, the staging rows are consumed
        if (j + 1 < n_my) issue_loads(nx_node0, nx_rows);   // the next graph's rows land during this graph's recurrence
        const unsigned int* red = red_s + (j & 1) * 4;
        const float xmax = __uint_as_float(red[0]), gmax = __uint_as_float(red[1]);
        const float dmax = (float)red[2];
        if (tid == 0) { unsigned int* o = red_s + ((j + 1) & 1) * 4; o[0] = 0u; o[1] = 0u; o[2] = 0u; }
        unsigned int* trk = trk_s + (j & 1) * 16;
        if (TRACK && tid < 16) trk_s[((j + 1) & 1) * 16 + tid] = 0u;   // last read a graph ago
        const int eg = max(-100, min(110, expo_above(gmax)));   // |G| < 2^eg
        int e_cur = max(-100, min(110, expo_above(xmax)));     // |T_0| < 2^e_cur
        float bet1 = xmax, bet2 = 0.f;                          // bounds of |T_k-1|, |T_k-2|
        split_row(pt_a, tp, pow2f(15 - e_cur));
        split_row(pg_a, gr, pow2f(15 - eg));
        fence_proxy_async();
        tmem_wait_st_();
        tc_fence_before();
        bar_compute();   // #2
        const uint32_t id_adj = idesc_f16(64u, 1u, 0u);                 // A from tensor memory, B MN-major
        const uint32_t id_dw = idesc_f16(64u, 1u, 0u) | (1u << 15);     // A (the G part tile) MN-major too
        auto issue = [&](int k) {   // thread 0: recurrence product A T_k (unless T_k is the last one), then dW_k
            tc_fence_after();
            const uint32_t ptk = pt_a + (uint32_t)(k & 1) * HF_TILE_BYTES;
            if (k + 1 < K) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    if (ks == 0 || ks < nks) umma_f16_ts(tmem_base, tmem_base + BF_ADJ_COL + (uint32_t)(ks * 8), desc_sw128(ptk + (uint32_t)ks * 2048u), id_adj, ks > 0 ? 1u : 0u);
                umma_commit(bar_mma);
            }
            const uint32_t d = tmem_base + BF_DW_COL + 64u * (uint32_t)(k & 1);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                if (ks == 0 || ks < nks) umma_f16_ss(d, desc_sw128(pg_a + (uint32_t)ks * 2048u), desc_sw128(ptk + (uint32_t)ks * 2048u), id_dw, ks > 0 ? 1u : 0u);
            umma_commit(bar_dw + 8u * (uint32_t)(k & 1));
        };
        if (tid == 0) issue(0);
        if (tid < 32) gout[(size_t)K * 1024 + tid] = ((dbs_s[tid] + dbs_s[32 + tid]) + dbs_s[64 + tid]) + dbs_s[96 + tid];

        // dW_k accumulator -> gradient rows.  D rows: [G_h o | G_l' o] (lanes 0-31 / 32-63), D columns: [T_h f | T_l' f]
        auto drain = [&](int k, int e_k) {
            if (k & 1) { mbar_wait(bar_dw + 8, ph_dw1); ph_dw1 ^= 1u; } else { mbar_wait(bar_dw, ph_dw0); ph_dw0 ^= 1u; }
            tc_fence_after();
            if (q < 2) {
                uint32_t a[16], b[16];
                const uint32_t col = BF_DW_COL + 64u * (uint32_t)(k & 1) + (uint32_t)(16 * hh);
                tmem_ld16(tmem_lane + col, a);
                tmem_ld16(tmem_lane + col + 32u, b);
                tmem_wait_ld_();
                float w[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) w[e] = __uint_as_float(a[e]) - __uint_as_float(b[e]);
                float* st = stg_s + (k & 1) * 1024 + (16 * hh) * 32 + lane;
                if (q == 1) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) st[e * 32] = w[e];
                    bar_drain(k & 1, false);
                } else {
                    bar_drain(k & 1, true);
                    const float us = pow2f(e_k - 15) * pow2f(eg - 15);
                    float* dst = gout + ((size_t)k * 32 + 16 * hh) * 32 + lane;
#pragma unroll
                    for (int e = 0; e < 16; ++e) dst[e * 32] = (w[e] - st[e * 32]) * us;
                }
            }
            tc_fence_before();
        };

        int e_prev = e_cur;
        for (int k = 1; k < K; ++k) {
            // ---- T_k = c (A T_k-1) - T_k-2
            mbar_wait(bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc_fence_after();
            uint32_t vh[16], vl[16];
            tmem_ld16(tmem_lane + (uint32_t)(16 * hh), vh);
            tmem_ld16(tmem_lane + 32u + (uint32_t)(16 * hh), vl);
            tmem_wait_ld_();
            const float cfac = pow2f(e_prev - 15 + (k > 1 ? 1 : 0));   // (k > 1 ? 2 : 1) / tau_k-1
            const uint64_t C2 = pk2(cfac, cfac);
            float m = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint64_t dv = sub2(pk2(__uint_as_float(vh[2 * e]), __uint_as_float(vh[2 * e + 1])), pk2(__uint_as_float(vl[2 * e]), __uint_as_float(vl[2 * e + 1])));
                const uint64_t tk = fma2(dv, C2, pk2(-tpp[2 * e], -tpp[2 * e + 1]));
                tpp[2 * e] = tp[2 * e]; tpp[2 * e + 1] = tp[2 * e + 1];
                upk2(tk, tp[2 * e], tp[2 * e + 1]);
                if (TRACK) m = fmaxf(m, fmaxf(fabsf(tp[2 * e]), fabsf(tp[2 * e + 1])));
            }
            if (TRACK) {
                // the maxima of |T_k-1| and |T_k-2| are complete (their atomics preceded a barrier and an UMMA round): a bound of
                // |T_k| that overshoots by one step's factor at most - the a-priori recurrence loses ~3 bits per step
                if (k >= 2) bet1 = __uint_as_float(trk[k - 1]);
                if (k >= 3) bet2 = __uint_as_float(trk[k - 2]);
                const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
                if (lane == 0) atomicMax(trk + k, wm);
            }
            const float bet = (k > 1 ? 2.f : 1.f) * dmax * bet1 + bet2;
            bet2 = bet1;
            bet1 = bet;
            e_cur = max(-100, min(110, expo_above(bet)));
            split_row(pt_a + (uint32_t)(k & 1) * HF_TILE_BYTES, tp, pow2f(15 - e_cur));
            fence_proxy_async();
            tc_fence_before();
            bar_compute();
            if (tid == 0) issue(k);
            drain(k - 1, e_prev);   // under this step's UMMAs
            e_prev = e_cur;
        }
        drain(K - 1, e_prev);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, BF_TCOLS);
}

}  // namespace

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
bool cheb_backward_f16_eligible(const mho_batch_t* b, const mho_layer_t* layers, int n_layers, const void* X, const void* Y, const void* dY,
                                const void* dX, int max_smem_optin) {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MHO_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (dbg & 512) return false;   // MHO_DEBUG & 512: keep the CUDA-core VJP
    if (n_layers != 1 || dX != nullptr || b->vals != nullptr || b->adj_bits == nullptr || b->max_tile_rows > 128) return false;
    const mho_layer_t& L = layers[0];
    if (L.f_in != 32 || L.f_out != 32 || L.K < 2 || L.K > 10) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15u) || (reinterpret_cast<uintptr_t>(Y) & 15u) || (reinterpret_cast<uintptr_t>(dY) & 15u) ||
        (reinterpret_cast<uintptr_t>(b->adj_bits) & 15u))
        return false;
    return (size_t)BF_SMEM + 1024 <= (size_t)max_smem_optin;
}

cudaError_t cheb_backward_f16_launch(const mho_batch_t* b, const mho_layer_t* layers, const float* X, const float* Y, const float* dY,
                                     float* grads, long long n_params, int num_sms, cudaStream_t st) {
    BfParams p;
    memset(&p, 0, sizeof(p));
    p.graph_off = b->graph_off;
    p.adj_bits = b->adj_bits;
    p.n_graphs = b->n_graphs;
    p.X = X; p.Y = Y; p.dY = dY;
    p.grads = grads;
    p.n_params = n_params;
    p.K = layers[0].K;
    p.act = layers[0].act;
    p.slope = layers[0].slope;
    const size_t smem = (size_t)BF_SMEM + 1024;
    static int smem_set[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!smem_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(cheb_backward_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(cheb_backward_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set[dev & 63] = 1;
    }
    int grid = std::min(2 * num_sms, std::max(1, p.n_graphs));
    if (p.K > 6) cheb_backward_f16_kernel<true><<<grid, BF_THREADS, smem, st>>>(p);     // running-maximum scales
    else cheb_backward_f16_kernel<false><<<grid, BF_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}
