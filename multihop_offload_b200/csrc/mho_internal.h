// Internal (non-ABI) declarations shared by the libmho translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "../../include/mho.h"

struct mho_scratch_slot {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct mho_wkey {
    const void* W; const void* b; int K, f_in, f_out;
    bool operator==(const mho_wkey& o) const { return W == o.W && b == o.b && K == o.K && f_in == o.f_in && f_out == o.f_out; }
};

#define MHO_MAX_CHUNKS 8
#define MHO_EV_PER_SLOT (2 * MHO_MAX_CHUNKS + 2)   // per staging slot: upload / kernel events per chunk, start, done

struct mho_ctx {
    // pipelined host call: upload / download streams and per-chunk events
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev[2 * MHO_EV_PER_SLOT] = {};
    int64_t host_calls = 0;          // host-buffer calls so far: call i uses staging slot i & 1
    bool slot_used[2] = {false, false};
    // prepared-weight cache (see mho_invalidate_weights)
    std::vector<mho_wkey> wkey;
    bool wprep_valid = false;
    unsigned char* wprep = nullptr;
    size_t wprep_bytes = 0;
    int wprep_row_off[MHO_MAX_LAYERS] = {0};
    // bf16-part weight images of the dense-adjacency tcgen05 path (cheb_forward_dense.cu), cached the same way
    std::vector<mho_wkey> wdkey;
    bool wdense_valid = false;
    unsigned char* wdense = nullptr;
    size_t wdense_bytes = 0;
    int wd_off[MHO_MAX_LAYERS] = {0};
    int wd_bytes = 0;
    // fp16 two-part weight image of the second-generation dense kernel (cheb_forward_f16.cu), cached the same way
    mho_wkey wfkey = {nullptr, nullptr, 0, 0, 0};
    bool wf16_valid = false;
    unsigned char* wf16 = nullptr;
    size_t wf16_bytes = 0;
    // fp16 weight images of the fused K = 1 stack kernel (cheb_mlp_f16.cu)
    std::vector<mho_wkey> wmkey;
    bool wmlp_valid = false;
    unsigned char* wmlp = nullptr;
    size_t wmlp_bytes = 0;
    // fp16 W^T images of the K = 1 stack VJP (cheb_mlp_backward_f16.cu)
    std::vector<mho_wkey> wbkey;
    bool wmb_valid = false;
    unsigned char* wmb = nullptr;
    size_t wmb_bytes = 0;
    int* sched = nullptr;  // two zero-initialised ints: dynamic tile scheduler state (self re-arming)
    int device = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    int64_t launches = 0;
    std::vector<mho_scratch_slot> scratch;
};

void mho_set_error(const char* fmt, ...);
void* mho_scratch(mho_ctx* c, int slot, size_t bytes);

struct LayerDev;
struct FwdParams;
void mho_fill_layers(const mho_layer_t* layers, int n_layers, int total_nodes, LayerDev* out);
int wprep_layer_rows(int K, int f_out);
cudaError_t prepare_weights_launch(const LayerDev* layers, int n_layers, const int* row_off, unsigned char* out,
                                   cudaStream_t st);
bool cheb_dense_eligible(const mho_layer_t* layers, int n_layers, bool has_vals, bool has_bits, int max_tile_rows, int max_tile_nnz,
                         int max_smem_optin);
int cheb_dense_weight_bytes(const mho_layer_t* layers, int n_layers, int* w_off);
cudaError_t prepare_dense_weights_launch(const LayerDev* layers, int n_layers, const int* w_off, unsigned char* out, cudaStream_t st);
cudaError_t cheb_dense_launch(const FwdParams& fp, const unsigned char* wimg, const int* w_off, int w_bytes, int max_tile_nnz,
                              int num_sms, cudaStream_t st);
bool cheb_f16_eligible(const mho_layer_t* layers, int n_layers, bool has_vals, bool has_bits, bool has_saved, bool has_graph_starts,
                       int max_tile_rows, int max_tile_nnz, const void* X, const void* Y, const void* bits, int max_smem_optin);
int cheb_f16_weight_bytes(int K);
cudaError_t prepare_f16_weights_launch(const LayerDev& L, unsigned char* out, cudaStream_t st);
cudaError_t cheb_f16_launch(const FwdParams& fp, const unsigned char* wimg, int max_tile_nnz, int num_sms, int max_smem_optin, cudaStream_t st);
bool cheb_mlp_eligible(const mho_layer_t* layers, int n_layers, int max_tile_rows, const void* X, int max_smem_optin);
int cheb_mlp_weight_bytes(int n_layers);
cudaError_t prepare_mlp_weights_launch(const LayerDev* layers, int n_layers, unsigned char* out, cudaStream_t st);
cudaError_t cheb_mlp_launch(const FwdParams& fp, const unsigned char* wimg, int num_sms, cudaStream_t st);
bool cheb_backward_f16_eligible(const mho_batch_t* b, const mho_layer_t* layers, int n_layers, const void* X, const void* Y, const void* dY,
                                const void* dX, int max_smem_optin);
cudaError_t cheb_backward_f16_launch(const mho_batch_t* b, const mho_layer_t* layers, const float* X, const float* Y, const float* dY,
                                     float* grads, long long n_params, int num_sms, cudaStream_t st);
bool cheb_mlp_backward_eligible(const mho_batch_t* b, const mho_layer_t* layers, int n_layers, const void* X, const void* saved, const void* dX,
                                int max_smem_optin);
int cheb_mlp_backward_weight_bytes(int n_layers);
cudaError_t prepare_mlp_backward_weights_launch(const LayerDev* layers, int n_layers, unsigned char* out, cudaStream_t st);
cudaError_t cheb_mlp_backward_launch(const mho_batch_t* b, const LayerDev* layers, int n_layers, const float* X, const float* Y, const float* saved,
                                     const float* dY, float* grads, long long n_params, const unsigned char* wT, int num_sms, cudaStream_t st);
cudaError_t apsp_launch(int n_graphs, const int32_t* node_off, const int32_t* rowptr, const int32_t* colidx, const double* weight,
                        const int64_t* out_off, double* dist, int max_smem_optin, cudaStream_t st);
cudaError_t cheb_forward_launch(FwdParams& p, int max_tile_rows, int max_tile_nnz, int num_sms, int max_smem_optin,
                                cudaStream_t st, bool* too_large);
