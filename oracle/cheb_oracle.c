/*
 * cheb_oracle.c - plain-C fp64 restatement of the batched ChebConv layer/stack forward.
 *
 * TEST INFRASTRUCTURE ONLY (checker + the CPU thing bench.py times); never linked into the
 * product.  Restates spektral.layers.ChebConv.call [upstream] as invoked by
 * /root/reference/src/gnn_offloading_agent.py:95-110 via :149, one graph at a time (the
 * reference's eager per-graph call), graphs distributed over OpenMP threads.
 *   T_0 = X, T_1 = A X, T_k = 2 A T_{k-1} - T_{k-2};  Y = act(sum_k T_k W[k] + b)
 * Parity pinning: see oracle/chebnet_oracle.py header ("parity unpinned" at the Spektral/TF
 * boundary; validated against chebnet_oracle.py in tests/test_oracle_c.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { int K, f_in, f_out, act; double slope; const double* W; const double* b; } layer64_t;

static void spmm(const int32_t* restrict rp, const int32_t* restrict ci, const double* restrict va, int node0, int n, int F,
                 const double* restrict src, double* restrict dst) {
    for (int i = 0; i < n; ++i) {
        double* d = dst + (size_t)i * F;
        for (int f = 0; f < F; ++f) d[f] = 0.0;
        for (int e = rp[node0 + i]; e < rp[node0 + i + 1]; ++e) {
            const double v = va ? va[e] : 1.0;
            const double* s = src + (size_t)(ci[e] - node0) * F;
            for (int f = 0; f < F; ++f) d[f] += v * s[f];
        }
    }
}

static void gemm_acc(int n, int Fi, int Fo, const double* restrict T, const double* restrict W, double* restrict out) {
    for (int i = 0; i < n; ++i) {
        double acc[64];
        double* restrict o = out + (size_t)i * Fo;
        for (int c = 0; c < Fo; ++c) acc[c] = o[c];
        for (int f = 0; f < Fi; ++f) {
            const double t = T[(size_t)i * Fi + f];
            const double* restrict w = W + (size_t)f * Fo;
#pragma GCC ivdep
            for (int c = 0; c < Fo; ++c) acc[c] += t * w[c];
        }
        for (int c = 0; c < Fo; ++c) o[c] = acc[c];
    }
}

/* Returns 0 on success.  X [total_nodes, layers[0].f_in], Y [total_nodes, layers[L-1].f_out]. */
int cheb_stack_forward_f64(int n_graphs, const int32_t* graph_off, const int32_t* rowptr, const int32_t* colidx,
                           const double* vals, const layer64_t* layers, int n_layers, const double* X, double* Y,
                           int n_threads) {
    int maxn = 0, maxf = 0;
    for (int g = 0; g < n_graphs; ++g) {
        int n = graph_off[g + 1] - graph_off[g];
        if (n > maxn) maxn = n;
    }
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].f_in > maxf) maxf = layers[l].f_in;
        if (layers[l].f_out > maxf) maxf = layers[l].f_out;
    }
    int fail = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        const size_t sz = (size_t)(maxn > 0 ? maxn : 1) * maxf;
        double* buf = (double*)malloc(sizeof(double) * sz * 5);
        if (!buf) {
#pragma omp atomic write
            fail = 1;
        } else {
            double *h = buf, *t0 = buf + sz, *t1 = buf + 2 * sz, *t2 = buf + 3 * sz, *out = buf + 4 * sz;
#pragma omp for schedule(dynamic, 8)
            for (int g = 0; g < n_graphs; ++g) {
                const int node0 = graph_off[g], n = graph_off[g + 1] - node0;
                memcpy(h, X + (size_t)node0 * layers[0].f_in, sizeof(double) * (size_t)n * layers[0].f_in);
                for (int l = 0; l < n_layers; ++l) {
                    const layer64_t* L = &layers[l];
                    const int Fi = L->f_in, Fo = L->f_out;
                    memset(out, 0, sizeof(double) * (size_t)n * Fo);
                    double *a = t0, *b = t1, *c = t2;
                    memcpy(a, h, sizeof(double) * (size_t)n * Fi);
                    gemm_acc(n, Fi, Fo, a, L->W, out);
                    if (L->K > 1) {
                        spmm(rowptr, colidx, vals, node0, n, Fi, a, b);
                        gemm_acc(n, Fi, Fo, b, L->W + (size_t)Fi * Fo, out);
                    }
                    for (int k = 2; k < L->K; ++k) {
                        spmm(rowptr, colidx, vals, node0, n, Fi, b, c);
                        for (size_t i = 0; i < (size_t)n * Fi; ++i) c[i] = 2.0 * c[i] - a[i];
                        gemm_acc(n, Fi, Fo, c, L->W + (size_t)k * Fi * Fo, out);
                        double* tmp = a; a = b; b = c; c = tmp;
                    }
                    double* dst = (l == n_layers - 1) ? Y + (size_t)node0 * Fo : h;
                    for (int i = 0; i < n; ++i)
                        for (int o = 0; o < Fo; ++o) {
                            double z = out[(size_t)i * Fo + o] + (L->b ? L->b[o] : 0.0);
                            if (L->act == 1) z = z > 0 ? z : 0.0;
                            else if (L->act == 2) z = z > 0 ? z : L->slope * z;
                            dst[(size_t)i * Fo + o] = z;
                        }
                }
            }
            free(buf);
        }
    }
    return fail;
}

int cheb_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
