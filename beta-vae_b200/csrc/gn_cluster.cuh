// Cluster-resident GroupNorm + ReLU + SE kernels (gn_se_cluster.cu).  `*handled` is false when the shape is
// outside what they cover (the multi-pass kernels of gn_se.cu then run).
#pragma once
#include <cuda_runtime.h>

namespace bvae {

template <typename T>
int gn_cluster_fwd(const T* y, T* out, const float* gamma, const float* beta, const float* w1, const float* b1,
                   const float* w2, const float* b2, int B, int P, int C, int R, float* stats, float* cpos, float* spos,
                   float* hidden, float* gate, cudaStream_t st, bool* handled);

// Writes dy and the per-sample G1, G2 (and dz2, dz1 with SE) that the gn_bwd_params_* kernels (gn_se.cu) sum over the batch.
template <typename T>
int gn_cluster_bwd(const T* dout, const T* y, T* dy, const float* gamma, const float* beta, const float* w1,
                   const float* w2, int B, int P, int C, int R, const float* stats, const float* cpos, const float* spos,
                   const float* hidden, const float* gate, float* G1, float* G2, float* dz2, float* dz1, cudaStream_t st,
                   bool* handled);

// true when the cluster-resident forward kernel covers the shape (element size 2 or 4 bytes)
bool gn_cluster_eligible(int B, int P, int C, int R, int elem_bytes);

}  // namespace bvae
