// Batched ChebConv forward (single layer or fused L-layer stack) for sm_100a.
//
// Replaces, for a whole batch of graph instances, the reference's eager per-graph call
//   ACOAgent.predict -> self.model([x_in, a_in])        (src/gnn_offloading_agent.py:144-150)
// of the Keras model built at :81-123 out of spektral.layers.ChebConv:
//   Y = act( sum_k T_k W[k] + b ),  T_0 = X, T_1 = A X, T_k = 2 A T_{k-1} - T_{k-2}.
//
// Design (see DESIGN.md "forward kernel"):
//   * persistent CTAs walk "tiles" = runs of consecutive graphs (block-diagonal operator => a run
//     of graphs is just a bigger graph) of at most rows_cap nodes; the next tile's X rows and CSR
//     slice are fetched with cp.async into a third tile buffer while the current tile computes;
//   * the CSR slice is turned in place into a stream of pre-swizzled shared-memory row offsets
//     whose low bits flag "last entry of its row" - each warp then walks a contiguous, nnz-balanced
//     chunk of that stream 4 entries per LDS.128, one lane per feature: every neighbour gather is a
//     single conflict-free 128 B wavefront and four of them are in flight per warp;
//   * the Chebyshev recurrence runs on-chip on two swizzled [rows][32] fp32 tiles (T_k overwrites
//     T_{k-2} in place);
//   * the dense contraction [T_0|..|T_{K-1}] . W runs on the tensor cores (mma.sync m16n8k8 TF32,
//     3xTF32 split in registers for fp32-grade accuracy) straight out of the same tiles via
//     ldmatrix, interleaved with the sparse step for T_{k+1};
//   * HBM traffic per tile is exactly X in, Y out, CSR once; T_k never leaves the SM.
#include "mho_common.cuh"

struct TileInfo {
    int node0, rows, nz0, nnz;
};

__device__ __forceinline__ TileInfo load_tile_info(const BatchDev& b, int tile) {
    TileInfo t;
    const int g0 = b.tile_off ? __ldg(b.tile_off + tile) : tile;
    const int g1 = b.tile_off ? __ldg(b.tile_off + tile + 1) : tile + 1;
    t.node0 = __ldg(b.graph_off + g0);
    const int node1 = __ldg(b.graph_off + g1);
    t.rows = node1 - t.node0;
    t.nz0 = __ldg(b.rowptr + t.node0);
    t.nnz = __ldg(b.rowptr + node1) - t.nz0;
    return t;
}

// -------------------------------------------------------------------------------------------
// Stage one layer's weights: Keras layout W[k][f][o] -> transposed swizzled fp32 image
// Wt[(k*fo_pad + o)][f] (128 B rows) so that ldmatrix yields mma B fragments; the TF32 hi/lo
// split happens in registers at use.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_weights(const LayerDev& L, unsigned char* wimg, float* bias_s, int tid) {
    const int fo_pad = pad8(L.f_out), fi_pad = pad8(L.f_in);
    const int per_k = fi_pad * fo_pad;
    const int total = L.K * per_k;
    for (int idx = tid; idx < total; idx += MHO_THREADS) {
        const int k = idx / per_k, rem = idx - k * per_k;
        const int f = rem / fo_pad, o = rem - f * fo_pad;
        float w = 0.f;
        if (f < L.f_in && o < L.f_out) w = __ldg(L.W + ((size_t)k * L.f_in + f) * L.f_out + o);
        *reinterpret_cast<float*>(wimg + (uint32_t)(k * fo_pad) * 128u + swz_off((uint32_t)o, (uint32_t)f)) = w;
    }
    if (tid < 32) bias_s[tid] = (L.b != nullptr && tid < L.f_out) ? __ldg(L.b + tid) : 0.f;
}

// -------------------------------------------------------------------------------------------
// acc[nt][4] += T[row0..row0+15][0..fi_pad) . W_k   (3xTF32: lo*hi + hi*lo + hi*hi, fp32 accumulate)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tile(float (&acc)[4][4], uint32_t T_addr, int row0, uint32_t w_k, int nchunks,
                                         int nnt, int lane) {
    const uint32_t arow = (uint32_t)row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const uint32_t a_base = T_addr + (arow << 7);
    const uint32_t a_key = arow & 7u;
    const uint32_t a_sel = (uint32_t)(lane >> 4);        // 0: chunk 2c, 1: chunk 2c+1
    const uint32_t b_r = (uint32_t)(lane & 7);
    const uint32_t b_sel = (uint32_t)((lane >> 3) & 1);  // k half
    const uint32_t b_nt = (uint32_t)(lane >> 4);         // which n-tile of the pair
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        uint32_t a[4], ah[4], al[4];
        ldmatrix_x4(a_base + ((((uint32_t)(2 * c) + a_sel) ^ a_key) << 4), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(__uint_as_float(a[i]), ah[i], al[i]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            if (np * 2 < nnt) {
                uint32_t o_nt = (uint32_t)(np * 2) + b_nt;
                if ((int)o_nt >= nnt) o_nt = (uint32_t)(nnt - 1);  // odd n-tile count: duplicate, result unused
                const uint32_t o = o_nt * 8u + b_r;
                uint32_t b[4], bh[4], bl[4];
                ldmatrix_x4(w_k + (o << 7) + ((((uint32_t)(2 * c) + b_sel) ^ (o & 7u)) << 4), b[0], b[1], b[2], b[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(__uint_as_float(b[i]), bh[i], bl[i]);
                mma_tf32(acc[np * 2], al, bh[0], bh[1]);
                mma_tf32(acc[np * 2], ah, bl[0], bl[1]);
                mma_tf32(acc[np * 2], ah, bh[0], bh[1]);
                if (np * 2 + 1 < nnt) {
                    mma_tf32(acc[np * 2 + 1], al, bh[2], bh[3]);
                    mma_tf32(acc[np * 2 + 1], ah, bl[2], bl[3]);
                    mma_tf32(acc[np * 2 + 1], ah, bh[2], bh[3]);
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// Sparse recurrence step, stream form.  The warp owns rows [rb, re) and walks their entries
//   first step : Tdst[r] = sum_j A[r,j] Tsrc[j]
//   later steps: Tdst[r] = 2 sum_j A[r,j] Tsrc[j] - Tdst[r]      (T_{k+1} overwrites T_{k-1})
// stream word = swizzled smem row offset | bit0 (last entry of its row) | bit1 (next row(s) empty).
// rp_s holds the RAW global row pointers of the tile (subtract nz0).
// -------------------------------------------------------------------------------------------
template <bool HAS_VALS>
__device__ __forceinline__ void spmm_walk(uint32_t Tsrc, uint32_t Tdst, bool first, int rb, int re, const int* rp_s,
                                          int nz0, uint32_t pre_a, uint32_t val_a, uint32_t key) {
    int row = rb;
    float acc = 0.f;
    auto emit_empty_run = [&]() {
        while (row < re && rp_s[row + 1] == rp_s[row]) {
            const uint32_t d = Tdst + (swz_row((uint32_t)row) ^ key);
            sts_f32(d, first ? 0.f : -lds_f32(d));
            ++row;
        }
    };
    auto step = [&](uint32_t pw, float t, float v) {
        acc = HAS_VALS ? fmaf(v, t, acc) : acc + t;
        if (pw & 1u) {
            const uint32_t d = Tdst + (swz_row((uint32_t)row) ^ key);
            sts_f32(d, first ? acc : 2.f * acc - lds_f32(d));
            acc = 0.f;
            ++row;
            if (pw & 2u) emit_empty_run();
        }
    };
    emit_empty_run();
    if (row >= re) return;
    int e = rp_s[row] - nz0;
    const int ee = rp_s[re] - nz0;
    for (; (e & 3) && e < ee; ++e) {  // head: up to the next 16 B boundary of the stream
        const uint32_t pw = lds_u32(pre_a + e * 4);
        const float t = lds_f32(Tsrc + ((pw & ~3u) ^ key));
        step(pw, t, HAS_VALS ? lds_f32(val_a + e * 4) : 1.f);
    }
    for (; e + 4 <= ee; e += 4) {  // body: 4 stream words per LDS.128, 4 gathers in flight
        const uint4 pw = lds_u128(pre_a + e * 4);
        const float t0 = lds_f32(Tsrc + ((pw.x & ~3u) ^ key));
        const float t1 = lds_f32(Tsrc + ((pw.y & ~3u) ^ key));
        const float t2 = lds_f32(Tsrc + ((pw.z & ~3u) ^ key));
        const float t3 = lds_f32(Tsrc + ((pw.w & ~3u) ^ key));
        uint4 vv = make_uint4(0, 0, 0, 0);
        if (HAS_VALS) vv = lds_u128(val_a + e * 4);
        step(pw.x, t0, __uint_as_float(vv.x));
        step(pw.y, t1, __uint_as_float(vv.y));
        step(pw.z, t2, __uint_as_float(vv.z));
        step(pw.w, t3, __uint_as_float(vv.w));
    }
    for (; e < ee; ++e) {  // tail
        const uint32_t pw = lds_u32(pre_a + e * 4);
        const float t = lds_f32(Tsrc + ((pw & ~3u) ^ key));
        step(pw, t, HAS_VALS ? lds_f32(val_a + e * 4) : 1.f);
    }
}

// row-per-warp fallback used when the CSR slice stays in global memory (huge / dense tiles)
template <bool HAS_VALS>
__device__ __forceinline__ void spmm_rows_global(uint32_t Tsrc, uint32_t Tdst, bool first, int rows, const int* rp_s,
                                                 const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                 int node0, int warp, uint32_t key) {
    for (int r = warp; r < rows; r += MHO_NWARPS) {
        const float s = gather_row<HAS_VALS, false>(Tsrc, rp_s[r], rp_s[r + 1], 0, 0, colidx, vals, node0, key);
        const uint32_t d = Tdst + (swz_row((uint32_t)r) ^ key);
        sts_f32(d, first ? s : 2.f * s - lds_f32(d));
    }
}

// -------------------------------------------------------------------------------------------
// Issue the loads of one tile: X rows -> swizzled tile buffer, raw rowptr / colidx / vals -> staging.
// Everything that can be is a cp.async (completion: cp_async_wait + __syncthreads by the caller).
// -------------------------------------------------------------------------------------------
template <bool HAS_VALS, bool STAGED>
__device__ __forceinline__ void issue_tile_loads(const FwdParams& p, const TileInfo& t, uint32_t Tbuf, uint32_t rp_a,
                                                 uint32_t pre_a, uint32_t val_a, int tid) {
    const int fi = p.layers[0].f_in;
    if ((fi & 3) == 0) {
        const int cpr = fi >> 2;  // 16 B chunks per row
        const float* src = p.X + (size_t)t.node0 * fi;
        const int total = t.rows * cpr;
        if (cpr == 8) {
            for (int c = tid; c < total; c += MHO_THREADS) {
                const uint32_t r = (uint32_t)c >> 3, ch = (uint32_t)c & 7u;
                cp_async16(Tbuf + (r << 7) + ((ch ^ (r & 7u)) << 4), src + (size_t)c * 4);
            }
        } else {
            for (int c = tid; c < total; c += MHO_THREADS) {
                const uint32_t r = (uint32_t)(c / cpr), ch = (uint32_t)(c - (int)r * cpr);
                cp_async16(Tbuf + (r << 7) + ((ch ^ (r & 7u)) << 4), src + (size_t)c * 4);
            }
            if (cpr & 1) {  // f_in = 4 (mod 8): zero the pad chunk so columns [f_in, pad8(f_in)) are defined
                for (int r = tid; r < t.rows; r += MHO_THREADS) {
                    const uint32_t a = Tbuf + ((uint32_t)r << 7) + ((((uint32_t)cpr) ^ ((uint32_t)r & 7u)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u));
                }
            }
        }
    } else {
        const int fi_pad = pad8(fi);
        for (int idx = tid; idx < t.rows * fi_pad; idx += MHO_THREADS) {
            const int r = idx / fi_pad, c = idx - r * fi_pad;
            sts_f32(Tbuf + swz_off((uint32_t)r, (uint32_t)c), c < fi ? __ldg(p.X + (size_t)(t.node0 + r) * fi + c) : 0.f);
        }
    }
    for (int i = tid; i <= t.rows; i += MHO_THREADS) cp_async4(rp_a + i * 4, p.b.rowptr + t.node0 + i);
    if (STAGED) {
        for (int e = tid; e < t.nnz; e += MHO_THREADS) {
            cp_async4(pre_a + e * 4, p.b.colidx + t.nz0 + e);
            if (HAS_VALS) cp_async4(val_a + e * 4, p.b.vals + t.nz0 + e);
        }
    }
}

template <int MT, bool HAS_VALS, bool STAGED>
__global__ void __launch_bounds__(MHO_THREADS, (MT == 1 ? 3 : (MT == 2 ? 2 : 1)))
cheb_forward_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t key = swz_key((uint32_t)lane);

    // ---- shared memory carve-up
    const uint32_t tile_bytes = (uint32_t)p.rows_cap * 128u;
    const int n_tbuf = p.prefetch ? 3 : 2;
    unsigned char* Wimg = smem + (size_t)n_tbuf * tile_bytes;
    float* bias_all = reinterpret_cast<float*>(Wimg + (size_t)p.w_rows_cap * 128);
    const int n_bias = p.w_resident ? p.n_layers : 1;
    int* csr0 = reinterpret_cast<int*>(bias_all + 32 * n_bias);
    const int rp_words = (p.rows_cap + 2 + 3) & ~3;
    const int csr_words = rp_words + p.nnz_cap * (HAS_VALS ? 2 : 1);
    const uint32_t smem_a = smem_u32(smem);
    const uint32_t w_a = smem_u32(Wimg);
    const uint32_t csr_a0 = smem_u32(csr0);

    // rotating tile buffers: bx = T_0 / X of the current tile, bs = scratch, bp = prefetch target
    uint32_t bx = smem_a, bs = smem_a + tile_bytes, bp = smem_a + 2u * tile_bytes;
    int cs = 0;  // CSR staging set of the current tile

    if (p.w_resident)
        for (int l = 0; l < p.n_layers; ++l)
            stage_weights(p.layers[l], Wimg + (size_t)p.w_row_off[l] * 128, bias_all + 32 * l, tid);

    int tile = blockIdx.x;
    if (tile >= p.b.n_tiles) return;
    TileInfo cur = load_tile_info(p.b, tile);
    TileInfo nxt = cur;
    bool has_nxt = (tile + (int)gridDim.x) < p.b.n_tiles;
    if (p.prefetch) {
        issue_tile_loads<HAS_VALS, STAGED>(p, cur, bx, csr_a0, csr_a0 + rp_words * 4, csr_a0 + (rp_words + p.nnz_cap) * 4, tid);
        cp_async_commit();
        if (has_nxt) nxt = load_tile_info(p.b, tile + gridDim.x);
    }

    for (; tile < p.b.n_tiles; tile += gridDim.x) {
        const uint32_t rp_a = csr_a0 + (uint32_t)(cs * csr_words) * 4u;
        const uint32_t pre_a = rp_a + rp_words * 4, val_a = pre_a + p.nnz_cap * 4;
        const int* rp_s = csr0 + cs * csr_words;
        TileInfo nn = nxt;
        bool has_nn = false;
        if (p.prefetch) {
            // next tile's loads go out now; the tile after that has its bounds fetched (consumed next iteration)
            if (has_nxt) {
                const uint32_t rp_n = csr_a0 + (uint32_t)((cs ^ 1) * csr_words) * 4u;
                issue_tile_loads<HAS_VALS, STAGED>(p, nxt, bp, rp_n, rp_n + rp_words * 4, rp_n + (rp_words + p.nnz_cap) * 4, tid);
                has_nn = (tile + 2 * (int)gridDim.x) < p.b.n_tiles;
                if (has_nn) nn = load_tile_info(p.b, tile + 2 * gridDim.x);
            }
            cp_async_commit();
            cp_async_wait<1>();  // the current tile's group has landed; the next tile's may still fly
        } else {
            issue_tile_loads<HAS_VALS, STAGED>(p, cur, bx, rp_a, pre_a, val_a, tid);
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();

        const int rows = cur.rows, node0 = cur.node0, nz0 = cur.nz0;
        const int n_mtiles = (rows + 15) >> 4;
        int rb = 0, re = 0;
        if (STAGED) {
            // column ids -> swizzled smem row offsets, with row-end / next-row-empty flags
            for (int r = tid; r < rows; r += MHO_THREADS) {
                const int e0 = rp_s[r] - nz0, e1 = rp_s[r + 1] - nz0;
                const bool next_empty = (r + 1 < rows) && (rp_s[r + 2] == rp_s[r + 1]);
                for (int e = e0; e < e1; ++e) {
                    uint32_t w = swz_row(lds_u32(pre_a + e * 4) - (uint32_t)node0);
                    if (e == e1 - 1) w |= 1u | (next_empty ? 2u : 0u);
                    sts_u32(pre_a + e * 4, w);
                }
            }
            // nnz-balanced contiguous row chunks: lane l finds boundary l of cost(r) = nnz_before(r) + 2 r
            {
                const int total = cur.nnz + 2 * rows;
                const int target = lane <= MHO_NWARPS ? (int)(((long long)total * lane) / MHO_NWARPS) : 0;
                int lo = 0, hi = rows;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((rp_s[mid] - nz0) + 2 * mid < target) lo = mid + 1; else hi = mid;
                }
                rb = __shfl_sync(0xffffffffu, lo, warp);
                re = __shfl_sync(0xffffffffu, lo, warp + 1);
            }
            __syncthreads();
        }

        for (int li = 0; li < p.n_layers; ++li) {
            const LayerDev& L = p.layers[li];
            float* bias_s = bias_all + (p.w_resident ? 32 * li : 0);
            uint32_t w_l = w_a + (uint32_t)(p.w_resident ? p.w_row_off[li] : 0) * 128u;
            if (!p.w_resident) {
                stage_weights(L, Wimg, bias_s, tid);
                __syncthreads();
            } else if (li > 0) {
                __syncthreads();  // H_l written by the previous layer's epilogue
            }
            const int fi_pad = pad8(L.f_in), fo_pad = pad8(L.f_out);
            const int nchunks = fi_pad >> 3, nnt = fo_pad >> 3;

            float acc[MT][4][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;

            uint32_t tk = bx, tprev = bs;  // T_k, and the buffer T_{k+1} is written to (holds T_{k-1})
            for (int k = 0; k < L.K; ++k) {
                const uint32_t w_k = w_l + (uint32_t)(k * fo_pad) * 128u;
                if (k + 1 < L.K) {
                    if (STAGED) spmm_walk<HAS_VALS>(tk, tprev, k == 0, rb, re, rp_s, nz0, pre_a, val_a, key);
                    else spmm_rows_global<HAS_VALS>(tk, tprev, k == 0, rows, rp_s, p.b.colidx, p.b.vals, node0, warp, key);
                }
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int mt = warp + m * MHO_NWARPS;
                    if (mt < n_mtiles) mma_tile(acc[m], tk, mt * 16, w_k, nchunks, nnt, lane);
                }
                __syncthreads();
                const uint32_t tmp = tk; tk = tprev; tprev = tmp;
            }

            // ---- epilogue: bias + activation; last layer -> Y, hidden layer -> tile bx (+ saved)
            const bool last = (li == p.n_layers - 1);
            const int fo = L.f_out;
            float* gout = last ? p.Y : (p.saved ? p.saved + p.layers[li + 1].saved_off : nullptr);
            const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int mt = warp + m * MHO_NWARPS;
                if (mt < n_mtiles) {
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        if (n < nnt) {
                            const int col = n * 8 + 2 * t4;
                            const float b0 = bias_s[col], b1 = bias_s[col + 1];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int r = mt * 16 + g + h * 8;
                                const float y0 = apply_act(acc[m][n][2 * h] + b0, L.act, L.slope);
                                const float y1 = apply_act(acc[m][n][2 * h + 1] + b1, L.act, L.slope);
                                if (!last) {
                                    // padded columns are exact zeros (zero W rows, zero bias, act(0)=0)
                                    const uint32_t a = bx + swz_off((uint32_t)r, (uint32_t)col);
                                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(y0), "f"(y1));
                                }
                                if (gout != nullptr && r < rows) {
                                    float* dst = gout + (size_t)(node0 + r) * fo + col;
                                    if ((fo & 1) == 0 && col + 1 < fo) {
                                        *reinterpret_cast<float2*>(dst) = make_float2(y0, y1);
                                    } else {
                                        if (col < fo) dst[0] = y0;
                                        if (col + 1 < fo) dst[1] = y1;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        // rotate: the prefetched buffer becomes the next tile's X, the old X/scratch become scratch/prefetch
        if (p.prefetch) {
            const uint32_t o_bx = bx, o_bs = bs;
            bx = bp; bs = o_bx; bp = o_bs;
            cs ^= 1;
            cur = nxt; nxt = nn; has_nxt = has_nn;
        } else {
            if (tile + (int)gridDim.x < p.b.n_tiles) cur = load_tile_info(p.b, tile + gridDim.x);
            __syncthreads();  // staging buffers are rewritten right away by the next tile's loads
        }
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------
// host launcher
// -------------------------------------------------------------------------------------------
static size_t fwd_smem_bytes(int rows_cap, int nnz_cap, int w_rows, int n_bias, bool has_vals, bool prefetch) {
    size_t s = (size_t)rows_cap * 128 * (prefetch ? 3 : 2) + (size_t)w_rows * 128 + (size_t)n_bias * 128;
    const size_t csr_words = (size_t)((rows_cap + 2 + 3) & ~3) + (size_t)nnz_cap * (has_vals ? 2 : 1);
    s += csr_words * 4 * (prefetch ? 2 : 1);
    return s + 16;
}

template <int MT, bool HAS_VALS, bool STAGED>
static cudaError_t launch_one(const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = cheb_forward_kernel<MT, HAS_VALS, STAGED>;
    // the attribute is sticky per (function, device): only raise it when a launch needs more
    static int smem_set[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if ((int)smem > smem_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set[dev & 63] = (int)smem;
    }
    kern<<<grid, MHO_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

template <int MT>
static cudaError_t launch_mt(const FwdParams& p, bool has_vals, bool staged, int grid, size_t smem, cudaStream_t st) {
    if (has_vals) return staged ? launch_one<MT, true, true>(p, grid, smem, st) : launch_one<MT, true, false>(p, grid, smem, st);
    return staged ? launch_one<MT, false, true>(p, grid, smem, st) : launch_one<MT, false, false>(p, grid, smem, st);
}

// Returns cudaSuccess or an error; *too_large set when the tile cannot fit in shared memory.
cudaError_t cheb_forward_launch(FwdParams& p, int max_tile_rows, int max_tile_nnz, int num_sms, int max_smem_optin,
                                cudaStream_t st, bool* too_large) {
    *too_large = false;
    p.rows_cap = pad16(max_tile_rows < 16 ? 16 : max_tile_rows);
    const bool has_vals = p.b.vals != nullptr;
    const int MTn = (p.rows_cap + 16 * MHO_NWARPS - 1) / (16 * MHO_NWARPS);
    if (MTn > 4) { *too_large = true; return cudaSuccess; }
    const int mt_sel = MTn <= 1 ? 1 : (MTn <= 2 ? 2 : 4);
    const int reg_limit = mt_sel == 1 ? 3 : (mt_sel == 2 ? 2 : 1);

    // weight images: keep every layer resident when that costs <= 48 KB, else restage per layer
    int w_sum = 0, w_max = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        const int r = p.layers[l].K * pad8(p.layers[l].f_out);
        p.w_row_off[l] = w_sum;
        w_sum += r;
        w_max = r > w_max ? r : w_max;
    }
    p.w_resident = (p.n_layers == 1 || w_sum * 128 <= 48 * 1024) ? 1 : 0;
    p.w_rows_cap = p.w_resident ? w_sum : w_max;
    if (!p.w_resident) for (int l = 0; l < p.n_layers; ++l) p.w_row_off[l] = 0;
    const int n_bias = p.w_resident ? p.n_layers : 1;

    // preference order: staged+prefetch with >=2 CTAs/SM, staged+prefetch, staged, global CSR
    const int nnz_cap = (max_tile_nnz + 3) & ~3;
    bool staged = true, prefetch = true;
    size_t smem = fwd_smem_bytes(p.rows_cap, nnz_cap, p.w_rows_cap, n_bias, has_vals, true);
    const size_t smem_np = fwd_smem_bytes(p.rows_cap, nnz_cap, p.w_rows_cap, n_bias, has_vals, false);
    auto per_sm_of = [&](size_t s) { int v = (int)((size_t)(228 * 1024) / (s + 1024)); return v > reg_limit ? reg_limit : v; };
    if (smem > (size_t)max_smem_optin || (per_sm_of(smem) < 2 && per_sm_of(smem_np) >= 2 && reg_limit >= 2)) {
        prefetch = false;
        smem = smem_np;
    }
    if (smem > (size_t)max_smem_optin) {
        staged = false;
        smem = fwd_smem_bytes(p.rows_cap, 0, p.w_rows_cap, n_bias, has_vals, false);
        if (smem > (size_t)max_smem_optin) { *too_large = true; return cudaSuccess; }
    }
    p.nnz_cap = staged ? nnz_cap : 0;
    p.prefetch = prefetch ? 1 : 0;
    int per_sm = per_sm_of(smem);
    if (per_sm < 1) per_sm = 1;
    int grid = num_sms * per_sm;
    if (grid > p.b.n_tiles) grid = p.b.n_tiles;
    if (grid < 1) grid = 1;
    switch (mt_sel) {
        case 1: return launch_mt<1>(p, has_vals, staged, grid, smem, st);
        case 2: return launch_mt<2>(p, has_vals, staged, grid, smem, st);
        default: return launch_mt<4>(p, has_vals, staged, grid, smem, st);
    }
}
