// Host-side launchers of the auxiliary kernels (vf_aux.cu).
#pragma once
#include "vf_common.cuh"

namespace vf {

constexpr int kMaxWorld = 8;  // GPUs of one NVLink box
struct PeerPtrs {
    unsigned long long* base[kMaxWorld];  // every rank's symmetric exchange buffer
};

int launch_finalize(double* partials, int nblocks, int n_dim, bool with_hist,
                    double* out_sums, double* out_hist, int accumulate, cudaStream_t stream);
int launch_finalize_epilogue(double* partials, int nblocks, int n_dim, bool with_hist,
                             int64_t n_events, int train, double* out_sums, double* out_hist,
                             double* divisions, double* result, double* result_host,
                             cudaStream_t stream);
size_t exchange_bytes(int n_dim, int world, int64_t n_cubes);
int launch_exchange_epilogue(double* partials, int nblocks, int n_dim, bool with_hist,
                             int64_t n_events, int train, double* out_sums, double* out_hist,
                             double* divisions, double* result, double* result_host, int rank,
                             int world, const PeerPtrs& peers, unsigned long long seq,
                             cudaStream_t stream);
int launch_plus_iteration_tail(double* workspace, int nblocks, int n_dim, int train,
                               double* out_hist, double* divisions, int64_t n_cubes, double* ress,
                               double* ress2, int adaptive, int min_neval, int64_t init_calls,
                               int32_t* n_ev, int64_t* ev_offset, double* arr_var, double* result,
                               double* result_host, int rank, int world, const PeerPtrs& peers,
                               unsigned long long seq, cudaStream_t stream);
int launch_refine(int n_dim, const double* hist, double* divisions, cudaStream_t stream);
int launch_epilogue(int n_dim, int64_t n_events, int train, const double* sums, const double* hist,
                    double* divisions, double* result, cudaStream_t stream);
int launch_uniforms(int n_dim, uint64_t ev_begin, int64_t n, uint64_t seed, uint32_t iteration,
                    int rng_bits, double* rnds, cudaStream_t stream);
int launch_sample(int mode, int n_dim, uint64_t ev_begin, int64_t n, double xjac, uint64_t seed,
                  uint32_t iteration, int rng_bits, const double* divisions, const Limits& lim,
                  double* x, double* w, int32_t* ind, cudaStream_t stream);
int launch_accumulate(int n_dim, int64_t n, const double* w, const double* f, const int32_t* ind,
                      int do_hist, double* partials, int* nblocks_out, cudaStream_t stream);
int launch_plus_epilogue(int64_t n_cubes, const double* ress, const double* ress2, int adaptive,
                         int min_neval, int64_t init_calls, int32_t* n_ev, int64_t* ev_offset,
                         double* arr_var, double* result, int64_t* n_events_out,
                         cudaStream_t stream);
int run_fp64_probe(int iters, double* tflops);

}  // namespace vf
