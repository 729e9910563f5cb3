// Shared declarations of the tcgen05 convolution kernels (conv_tc.cu, conv_tc_wgrad.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace bvae {

struct TcTap {
  int8_t map;   // which source tensor map (stride-2 fprop: parity view ph*2+pw)
  int8_t di;    // row offset of the tap in the source grid
  int8_t dj;    // column offset
  int8_t widx;  // weight slice r*3+s
};

struct TcMaps {
  CUtensorMap src[4];
  CUtensorMap w;
};

struct TcParams {
  int ncls;                      // 1, or 4 output parity classes of a stride-2 dgrad (blockIdx.z)
  int ntaps[4];
  TcTap taps[4][9];
  int kc, kchunks;               // channels per chunk (64/32/16) and chunks per tap
  int bn;                        // output channels per CTA (UMMA N)
  int tw, th, tn;                // tile box: tw*th*tn = 128 pixels
  int tiles_w, tiles_h;
  int B, N;                      // batch, total output channels
  int DH, DW, ds, po, pw;        // destination tensor and scatter (dst row = oi*ds + po)
  int stages, tmem_cols;
  __nv_bfloat16* dst;
  const float* bias;
};

// kind 0 = fprop, 1 = dgrad.  True when the tcgen05 kernel covers the shape (bf16 only).
bool conv_tc_supported(int H, int W, int Cin, int Cout, int stride, int kind);
int conv_tc_fprop(const void* x, const void* wp_nk, const float* bias, void* y, int B, int H, int W, int Cin,
                  int Cout, int stride, cudaStream_t st, double* stat_part = nullptr);
int conv_tc_dgrad(const void* dy, const void* wpt_nk, void* dx, int B, int H, int W, int Cin, int Cout, int stride,
                  cudaStream_t st);

// ---- stride-1 halo kernel (conv_tc_halo.cu): persistent, one shared-memory halo per tile, weights resident ----
bool conv_halo_supported(int H, int W, int K, int N);
// stat_part (optional): [B][conv_halo_stat_parts(H, W, N)][2] doubles, per-sample partial (sum, sum of squares) of dst
int conv_halo_launch(const void* src, const void* wp_nk, const float* bias, void* dst, int B, int H, int W, int K, int N,
                     int dgrad, cudaStream_t st, double* stat_part = nullptr);
int conv_halo_stat_parts(int H, int W, int N);

bool conv_halo_final_supported(int H, int W, int C, int Co);
int conv_halo_final_rows(int Co);   // rows of the [rows][C] weight pack (9 * Co padded to the UMMA N)
int conv_halo_final_fwd(const void* h, const void* wg, const float* bias, const float* target, float* recon,
                        double* lpart, int* n_lpart, int loss_type, int B, int H, int W, int C, int Co, cudaStream_t st);

// ---- final conv backward on tcgen05 (finalconv_tc.cu): C = 64 -> 1 channel ----
bool finalconv_bwd_tc_supported(int H, int W, int C, int Co);
size_t finalconv_bwd_tc_dcol_bytes(int B, int H, int W);
int finalconv_bwd_tc(const void* h, const float* w, const float* dlogit, void* dh, float* wpart, float* bpart, void* dcol_ws,
                     int* ncta, int B, int H, int W, cudaStream_t st);

// ---- wgrad (conv_tc_wgrad.cu) ----
struct WgradTcPlan {
  int bn, slots, groups;           // co tile, accumulator slots per CTA, row-block groups
  int tiles_total, tiles_per_split, splits;
  int bias_chunks, bias_pix_per_chunk;
  size_t part_bytes, bpart_bytes;  // workspace: partial dW [split][9][Cin][Cout] (or tile order), partial db [chunk][Cout]
};
bool conv_tc_wgrad_supported(int H, int W, int Cin, int Cout, int stride);
WgradTcPlan conv_tc_wgrad_plan(int B, int H, int W, int Cin, int Cout, int stride);
int conv_tc_wgrad(const void* x, const void* dy, float* part, float* bpart, const WgradTcPlan& pl, int B, int H,
                  int W, int Cin, int Cout, int stride, cudaStream_t st);

// true: conv_tc_wgrad leaves its partials in accumulator-tile order (bias sums included); reduce them with
// conv_tc_wgrad_tile_reduce instead of the [split][tap][Cin][Cout] reduce of conv_simt.cu
bool conv_tc_wgrad_tile_partials(int H, int W, int Cin, int Cout, int stride);
int conv_tc_wgrad_tile_reduce(const float* part, float* dw, float* dbias, const WgradTcPlan& pl, int Cin, int Cout, int stride,
                              cudaStream_t st);

// ---- stride-2 dgrad, halo formulation (conv_tc_s2dgrad.cu): H, W = extent of dx; N = Cin (dx), K = Cout (dy) ----
bool conv_s2_dgrad_halo_supported(int H, int W, int N, int K);
int conv_s2_dgrad_halo(const void* dy, const void* wpt_nk, void* dx, int B, int H, int W, int N, int K, cudaStream_t st);

// ---- small-channel wgrad on the warp-level MMA path (conv_wgrad_mma.cu): 16/32/64 channels, not 64 x 64 ----
bool conv_wgrad_mma_small_supported(int H, int W, int Cin, int Cout, int stride);
WgradTcPlan conv_wgrad_mma_small_plan(int B, int H, int W, int Cin, int Cout, int stride);
int conv_wgrad_mma_small(const void* x, const void* dy, float* part, float* bpart, const WgradTcPlan& pl, int B, int H, int W,
                         int Cin, int Cout, int stride, cudaStream_t st);

}  // namespace bvae
