// Batched ChebConv forward (single layer or fused L-layer stack) for sm_100a.
//
// Replaces, for a whole batch of graph instances, the reference's eager per-graph call
//   ACOAgent.predict -> self.model([x_in, a_in])        (src/gnn_offloading_agent.py:144-150)
// of the Keras model built at :81-123 out of spektral.layers.ChebConv:
//   Y = act( sum_k T_k W[k] + b ),  T_0 = X, T_1 = A X, T_k = 2 A T_{k-1} - T_{k-2}.
//
// Design (see DESIGN.md "forward kernel"):
//   * one persistent CTA per SM slot walks "tiles" = runs of consecutive graphs (block-diagonal
//     operator => a run of graphs is just a bigger graph) of at most rows_cap nodes;
//   * the tile's CSR slice is staged once in shared memory with column ids pre-translated into
//     swizzled smem row offsets, the node features land in a 128B-swizzled [rows][32] tile;
//   * the Chebyshev recurrence runs entirely on-chip on two such tiles (T_k overwrites T_{k-2}
//     in place), one warp per operator row, one lane per feature => every neighbour gather is a
//     single conflict-free 128 B shared-memory wavefront, partial sums stay in a register;
//   * the dense contraction [T_0|..|T_{K-1}] . W runs on the tensor cores (mma.sync m16n8k8
//     TF32, 3xTF32 split for fp32-grade accuracy) straight out of the same swizzled tiles via
//     ldmatrix, overlapped with the sparse step for T_{k+1};
//   * HBM traffic per tile is exactly X in, Y out, CSR once; T_k never leaves the SM.
#include "mho_common.cuh"


// -------------------------------------------------------------------------------------------
// Stage one layer's weights: Keras layout W[k][f][o] -> transposed, hi/lo split, swizzled
// images Wt_hi/Wt_lo[(k*fo_pad + o)][f] (128 B rows) so that ldmatrix yields mma B fragments.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_weights(const LayerDev& L, unsigned char* whi, unsigned char* wlo,
                                              float* bias_s, int tid) {
    const int fo_pad = pad8(L.f_out), fi_pad = pad8(L.f_in);
    const int per_k = fi_pad * fo_pad;
    const int total = L.K * per_k;
    for (int idx = tid; idx < total; idx += MHO_THREADS) {
        const int k = idx / per_k, rem = idx - k * per_k;
        const int f = rem / fo_pad, o = rem - f * fo_pad;
        float w = 0.f;
        if (f < L.f_in && o < L.f_out) w = __ldg(L.W + ((size_t)k * L.f_in + f) * L.f_out + o);
        uint32_t hi, lo;
        split_tf32(w, hi, lo);
        const uint32_t off = (uint32_t)(k * fo_pad) * 128u + swz_off((uint32_t)o, (uint32_t)f);
        *reinterpret_cast<uint32_t*>(whi + off) = hi;
        *reinterpret_cast<uint32_t*>(wlo + off) = lo;
    }
    if (tid < 32) bias_s[tid] = (L.b != nullptr && tid < L.f_out) ? __ldg(L.b + tid) : 0.f;
}

// -------------------------------------------------------------------------------------------
// acc[nt][4] += T[row0..row0+15][0..fi_pad) . W_k   (3xTF32)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tile(float (&acc)[4][4], uint32_t T_addr, int row0, uint32_t whi_k,
                                         uint32_t wlo_k, int nchunks, int nnt, int lane) {
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
                const uint32_t boff = (o << 7) + ((((uint32_t)(2 * c) + b_sel) ^ (o & 7u)) << 4);
                uint32_t bh[4], bl[4];
                ldmatrix_x4(whi_k + boff, bh[0], bh[1], bh[2], bh[3]);
                ldmatrix_x4(wlo_k + boff, bl[0], bl[1], bl[2], bl[3]);
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
// One step of the recurrence for the rows owned by this warp:
//   first step : Tdst[r] = sum_j A[r,j] Tsrc[j]
//   later steps: Tdst[r] = 2 sum_j A[r,j] Tsrc[j] - Tdst[r]      (T_{k+1} overwrites T_{k-1})
// lane = feature column; every gather is one 128 B wavefront (gather_row in mho_common.cuh).
// -------------------------------------------------------------------------------------------
template <bool HAS_VALS, bool STAGED>
__device__ __forceinline__ void spmm_step(uint32_t Tsrc, uint32_t Tdst, bool first, int rows, const int* rp_s,
                                          uint32_t pre_a, uint32_t val_a, const int32_t* __restrict__ colidx,
                                          const float* __restrict__ vals, int node0, int warp, int lane,
                                          bool lane_on) {
    const uint32_t key = swz_key((uint32_t)lane);
    for (int r = warp; r < rows; r += MHO_NWARPS) {
        if (lane_on) {
            const float s = gather_row<HAS_VALS, STAGED>(Tsrc, rp_s[r], rp_s[r + 1], pre_a, val_a, colidx, vals, node0, key);
            const uint32_t d = Tdst + (swz_row((uint32_t)r) ^ key);
            sts_f32(d, first ? s : 2.f * s - lds_f32(d));
        }
    }
}

template <int MT, bool HAS_VALS, bool STAGED>
__global__ void __launch_bounds__(MHO_THREADS, (MT <= 2 ? 2 : 1))
cheb_forward_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- shared memory carve-up (all tile buffers 1024 B aligned: rows_cap % 16 == 0 => rows_cap*128 % 1024 == 0 iff rows_cap % 8 == 0)
    unsigned char* T0 = smem;
    unsigned char* T1 = T0 + (size_t)p.rows_cap * 128;
    unsigned char* Whi = T1 + (size_t)p.rows_cap * 128;
    unsigned char* Wlo = Whi + (size_t)p.w_rows_cap * 128;
    float* bias_s = reinterpret_cast<float*>(Wlo + (size_t)p.w_rows_cap * 128);
    int* rp_s = reinterpret_cast<int*>(bias_s + 32);
    const int rp_words = (p.rows_cap + 1 + 3) & ~3;
    uint32_t* pre_s = reinterpret_cast<uint32_t*>(rp_s + rp_words);
    float* val_s = reinterpret_cast<float*>(pre_s + p.nnz_cap);
    const uint32_t T_a[2] = {smem_u32(T0), smem_u32(T1)};
    const uint32_t whi_a = smem_u32(Whi), wlo_a = smem_u32(Wlo);
    const uint32_t pre_a = smem_u32(pre_s), val_a = smem_u32(val_s);

    const bool single = (p.n_layers == 1);
    if (single) stage_weights(p.layers[0], Whi, Wlo, bias_s, tid);  // made visible by the first tile's barrier

    for (int tile = blockIdx.x; tile < p.b.n_tiles; tile += gridDim.x) {
        const int g0 = p.b.tile_off ? __ldg(p.b.tile_off + tile) : tile;
        const int g1 = p.b.tile_off ? __ldg(p.b.tile_off + tile + 1) : tile + 1;
        const int node0 = __ldg(p.b.graph_off + g0), node1 = __ldg(p.b.graph_off + g1);
        const int rows = node1 - node0;
        const int nz0 = __ldg(p.b.rowptr + node0);
        const int nnz = __ldg(p.b.rowptr + node1) - nz0;
        const int n_mtiles = (rows + 15) >> 4;

        // ---- stage the tile: CSR slice (column ids -> swizzled smem row offsets) and X
        for (int i = tid; i <= rows; i += MHO_THREADS)
            rp_s[i] = STAGED ? (__ldg(p.b.rowptr + node0 + i) - nz0) : __ldg(p.b.rowptr + node0 + i);
        if (STAGED) {
            for (int e = tid; e < nnz; e += MHO_THREADS) {
                pre_s[e] = swz_row((uint32_t)(__ldg(p.b.colidx + nz0 + e) - node0));
                if (HAS_VALS) val_s[e] = __ldg(p.b.vals + nz0 + e);
            }
        }
        {
            const LayerDev& L0 = p.layers[0];
            const int fi = L0.f_in;
            if (fi == 32) {
                const float* src = p.X + (size_t)node0 * 32;
                for (int c = tid; c < rows * 8; c += MHO_THREADS) {
                    const uint32_t r = (uint32_t)c >> 3, ch = (uint32_t)c & 7u;
                    cp_async16(T_a[0] + (r << 7) + ((ch ^ (r & 7u)) << 4), src + (size_t)c * 4);
                }
                cp_async_commit();
                cp_async_wait<0>();
            } else {
                const int fi_pad = pad8(fi);
                for (int idx = tid; idx < rows * fi_pad; idx += MHO_THREADS) {
                    const int r = idx / fi_pad, c = idx - r * fi_pad;
                    const float v = c < fi ? __ldg(p.X + (size_t)(node0 + r) * fi + c) : 0.f;
                    *reinterpret_cast<float*>(T0 + swz_off((uint32_t)r, (uint32_t)c)) = v;
                }
            }
        }
        __syncthreads();

        for (int li = 0; li < p.n_layers; ++li) {
            const LayerDev& L = p.layers[li];
            if (!single) {
                stage_weights(L, Whi, Wlo, bias_s, tid);
                __syncthreads();
            }
            const int fi_pad = pad8(L.f_in), fo_pad = pad8(L.f_out);
            const int nchunks = fi_pad >> 3, nnt = fo_pad >> 3;
            const bool lane_on = lane < fi_pad;

            float acc[MT][4][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;

            int cur = 0;
            for (int k = 0; k < L.K; ++k) {
                const uint32_t whi_k = whi_a + (uint32_t)(k * fo_pad) * 128u;
                const uint32_t wlo_k = wlo_a + (uint32_t)(k * fo_pad) * 128u;
                // sparse step for T_{k+1} (reads T_k, overwrites T_{k-1}) ...
                if (k + 1 < L.K)
                    spmm_step<HAS_VALS, STAGED>(T_a[cur], T_a[cur ^ 1], k == 0, rows, rp_s, pre_a, val_a,
                                                p.b.colidx, p.b.vals, node0, warp, lane, lane_on);
                // ... and the dense contribution of T_k on the tensor cores
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int mt = warp + m * MHO_NWARPS;
                    if (mt < n_mtiles) mma_tile(acc[m], T_a[cur], mt * 16, whi_k, wlo_k, nchunks, nnt, lane);
                }
                __syncthreads();
                cur ^= 1;
            }

            // ---- epilogue: bias + activation; last layer -> Y, hidden layer -> smem tile (+ saved)
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
                                    *reinterpret_cast<float2*>(T0 + swz_off((uint32_t)r, (uint32_t)col)) =
                                        make_float2(y0, y1);
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
            // the barrier that publishes H_{l+1} in T0 is the one after the next stage_weights
        }
        if (single) __syncthreads();  // T0/rp_s are rewritten by the next tile's staging
    }
}

// -------------------------------------------------------------------------------------------
// host launcher
// -------------------------------------------------------------------------------------------
size_t cheb_forward_smem_bytes(int rows_cap, int nnz_cap, int w_rows_cap, bool has_vals) {
    size_t s = (size_t)rows_cap * 128 * 2 + (size_t)w_rows_cap * 128 * 2 + 128;
    s += (size_t)((rows_cap + 1 + 3) & ~3) * 4;
    s += (size_t)nnz_cap * 4 * (has_vals ? 2 : 1);
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
    int w_rows = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        const int r = p.layers[l].K * pad8(p.layers[l].f_out);
        w_rows = r > w_rows ? r : w_rows;
    }
    p.w_rows_cap = w_rows;
    const bool has_vals = p.b.vals != nullptr;
    const int MTn = (p.rows_cap + 16 * MHO_NWARPS - 1) / (16 * MHO_NWARPS);
    if (MTn > 4) { *too_large = true; return cudaSuccess; }
    int nnz_cap = (max_tile_nnz + 3) & ~3;
    bool staged = true;
    size_t smem = cheb_forward_smem_bytes(p.rows_cap, nnz_cap, w_rows, has_vals);
    if (smem > (size_t)max_smem_optin) {  // CSR slice does not fit next to the tiles: read it through L1/L2
        staged = false;
        nnz_cap = 0;
        smem = cheb_forward_smem_bytes(p.rows_cap, 0, w_rows, has_vals);
        if (smem > (size_t)max_smem_optin) { *too_large = true; return cudaSuccess; }
    }
    p.nnz_cap = nnz_cap;
    // persistent grid: as many CTAs as fit per SM (smem-limited), never more than tiles
    int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int mt_sel = MTn <= 1 ? 1 : (MTn <= 2 ? 2 : 4);
    const int reg_limit = mt_sel <= 2 ? 2 : 1;
    if (per_sm > reg_limit) per_sm = reg_limit;
    int grid = num_sms * per_sm;
    if (grid > p.b.n_tiles) grid = p.b.n_tiles;
    if (grid < 1) grid = 1;
    switch (mt_sel) {
        case 1: return launch_mt<1>(p, has_vals, staged, grid, smem, st);
        case 2: return launch_mt<2>(p, has_vals, staged, grid, smem, st);
        default: return launch_mt<4>(p, has_vals, staged, grid, smem, st);
    }
}
