// tcgen05 (5th-gen tensor core) gather-GEMM path of libcgvc.so: bf16 hi/lo split operands, fp32 TMEM accumulators.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stddef.h>
#include <vector>

#define TC_UNSUPPORTED (-12345)

// One convolution (or gated pair of convolutions sharing an input) whose weights are kept as bf16 hi/lo planes in
// the operand layouts the tensor-core kernels consume.
struct TcLayer {
  size_t ka, kg, ba, bg;          // offsets (elements) of kernel_a / kernel_g / bias_a / bias_g in the PARAM arena
  int kh, kw, cin, cout, gated;   // TF shapes [kh,kw,cin,cout]; Ntot = cout * (gated ? 2 : 1)
  int shuffle;                    // 2: the layer's output goes through the pixel shuffler (module.py:135-146); forward weight rows are then
                                  // stored so that one 256-column tile holds both conv channels of 64 post-shuffle channels (tc_gemm.cu perm_row)
  int fold;                       // > 0: a 1 x 1 layer whose `cout` output columns are (t, n) pairs, column = t * (cout / fold) + n, of a TF kernel
                                  // [1, fold, cin, cout / fold] at ka (a stride-1 multi-tap layer with few output channels, computed as one dense GEMM
                                  // whose per-tap results the caller sums with the tap shifts: engine.cu `edge_lower`); not gated, no bias
  // padded extents: x_k = rounded up to 64 (contraction), x_n = rounded up to 128 (output tile); pads are zero
  __nv_bfloat16 *wf_hi, *wf_lo;   // forward B operand  [taps][Ntot_n][cin_k]   (K = cin contiguous)
  __nv_bfloat16 *wd_hi, *wd_lo;   // dgrad   B operand  [taps][cin_n][Ntot_k]   (K = Ntot contiguous)
  float* bias;                    // [Ntot_n]
  CUtensorMap tm_f_hi, tm_f_lo;   // TMA descriptors of wf (box [1][BN][64]) and wd, built once the planes are allocated
  CUtensorMap tm_d_hi, tm_d_lo;
  CUtensorMap tm_f64_hi, tm_f64_lo, tm_d64_hi, tm_d64_lo;   // ... and with 64-row boxes (128-wide pair tiles)
  CUtensorMap tm_f2_hi, tm_f2_lo, tm_d2_hi, tm_d2_lo;   // the same planes with half-tile boxes: each CTA of a pair loads half a weight tile
  // CGVC_PREC_F16F8 (allocated when TcWeights::quant): forward operand [taps][Ntot_n][cin_q], cin_q = cin rounded up
  // to 128, as fp16 + two e4m3 planes with the weight scales of kernels.cuh
  void* wq16; uint8_t *wq8hi, *wq8lo;
  CUtensorMap tm_q16, tm_q8hi, tm_q8lo;
  CUtensorMap tm_q16h, tm_q8hih, tm_q8loh;          // the same planes with half-tile boxes (CTA-pair kernels)
  // F16F8 training (TcWeights::quant_bwd): data-gradient operand [taps][cin_n][nt_q], nt_q = Ntot rounded up to 128
  void* wdq16; uint8_t *wdq8hi, *wdq8lo;
  CUtensorMap tm_dq16, tm_dq8hi, tm_dq8lo, tm_dq16h, tm_dq8hih, tm_dq8loh;
};

struct TcWeights {
  std::vector<TcLayer> layers;
  void* pool = nullptr;
  size_t pool_bytes = 0;
  bool ready = false;
  bool quant = false;             // also keep the F16F8 forward planes (set before tc_alloc)
  bool wgrad16 = false;           // F16F8 only: weight gradients from the fp16 planes alone (one MMA unit per product instead of two)
  bool quant_bwd = false;         // ... and the F16F8 data-gradient planes (training in that precision)
  void* prep_jobs = nullptr;      // device job table of the batched F16F8 plane kernel (tc_gemm.cu PrepJob), one job per layer branch
  std::vector<int> job_first;     // first block of every job + total (size jobs + 1)
  std::vector<size_t> job_ka;     // PARAM offset of the job's layer (range filter of tc_refresh_weights_range)
};

// what the fused forward epilogue needs besides the convolution itself (see tc_conv_fwd_fused)
struct TcFuse {
  int R;                                            // positions per sample of the layer output
  const float *gamma_a, *beta_a, *gamma_g, *beta_g; // instance-norm affine parameters (g: gate branch, null when not gated)
  float* stats;                                     // [n,4,C] saved (mean, rstd) pairs for the backward pass
  const float* resid;                               // residual input [rows, C] (non-gated residual layer) or null
  float* y; __nv_bfloat16 *y_hi, *y_lo;             // outputs [rows, C] (fp32 optional)
};

// what the fused backward epilogue of a data-gradient launch needs (see tc_conv_dgrad_fused): the gradient this launch computes
// is d loss / d (output of an upstream layer); that layer's instance-norm (+ GLU) backward runs in the epilogue
struct TcBwdFuse {
  int R;                                            // positions per sample
  int gated;                                        // 1: y = IN(a) * sigmoid(IN(g)) (EPI 3); 0: y = resid + IN(a) (EPI 4, dx also receives dY)
  const float* bp; int bp_ld;                       // the upstream layer's saved pre-norm conv outputs [rows, bp_ld]
  const float* stats;                               // its saved (mean_a, rstd_a, mean_g, rstd_g) [n,4,C]
  const float *gamma_a, *beta_a, *gamma_g, *beta_g;
  __nv_bfloat16 *dp_hi, *dp_lo; int dp_ld;          // its dP planes (output) [rows, dp_ld]
  float *dbeta_a, *dgamma_a, *dbeta_g, *dgamma_g;   // parameter gradients (accumulated; null: data gradient only)
};

int tc_register(TcWeights& w, size_t ka, size_t kg, size_t ba, size_t bg, int kh, int kw, int cin, int cout, int gated, int shuffle = 1, int fold = 0);
int tc_alloc(TcWeights& w);                                     // cudaError_t as int
void tc_free(TcWeights& w);
int tc_refresh_weights(TcWeights& w, const float* params, cudaStream_t st);
int tc_refresh_weights_range(TcWeights& w, const float* params, size_t begin, size_t end, cudaStream_t st);

// activation / gradient planes handed to these functions have their channel count rounded up to a multiple of 64
// (zero-filled): x [n,H,W,ru64(cin)], dP [rows, ru64(Ntot)]
// P[rows, Ntot] = conv(x) + bias
int tc_conv_fwd(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                int n, int H, int W, int sh, int sw, float* P, cudaStream_t st);
// Same, with instance norm (+ GLU | + residual) fused into the epilogue when the shape allows (1-D layer, whole samples per
// 128-row tile: R in {32,64,128}); *fused tells the caller whether it happened (if not, P is written and the caller runs
// the separate instance-norm kernels).
int tc_conv_fwd_fused(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                      int n, int H, int W, int sh, int sw, float* P, const TcFuse& fuse, bool* fused, cudaStream_t st);
// dx[n,H,W,cin] (+)= dgrad(dP)            (dP planes [rows_out, Ntot]; H, W are the INPUT dims)
int tc_conv_dgrad(TcWeights& w, int slot, int precision, const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo,
                  int n, int H, int W, int sh, int sw, float* dx, int accumulate, cudaStream_t st);
// Same, with the upstream layer's instance-norm (+ GLU) backward fused into the epilogue when the shape allows (stride-1 1-D
// layer, whole samples per 128-row tile): the launch then writes that layer's dP planes (and, for gated = 0, dx = dY) instead
// of / besides dx; *fused tells whether it happened (if not, dx holds the plain data gradient as with tc_conv_dgrad).
int tc_conv_dgrad_fused(TcWeights& w, int slot, int precision, const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo,
                        int n, int H, int W, int sh, int sw, float* dx, int accumulate, const TcBwdFuse& fuse, bool* fused, cudaStream_t st);
// dW_a/dW_g (TF layout) += x^T dP ; db += colsum(dP)
int tc_conv_wgrad(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                  const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, int n, int H, int W, int sh, int sw,
                  float* dwa, float* dwg, float* dba, float* dbg, cudaStream_t st);
// self-contained versions for unit tests (fp32 in/out, temporary planes allocated internally)
int tc_conv_fwd_adhoc(int precision, const float* x, const float* w, const float* bias, float* y,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, cudaStream_t st);
int tc_conv_bwd_adhoc(int precision, const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, cudaStream_t st, int w16 = 0);   // w16: F16F8 weight gradient from the fp16 planes alone

// per-launch CUDA-event timing of the tensor-core kernels (class 0 = forward/dgrad kernel with the plain epilogue,
// 1 = wgrad kernel, 2 = forward kernel with the fused instance-norm epilogue)
void tc_profile_enable(int on);
bool tc_profile_is_on();
int tc_profile_collect(double ms[3], double flops[3], long long launches[3]);
int tc_profile_launches(double* ms, double* flops, long long* meta4, int capacity, int* n_out);
void tc_set_prep_batched(int v);   // 1 (default): F16F8 weight planes of all layers in one launch; 0: per-layer kernels
void tc_set_pair(int v);      // 1 (default): CTA-pair kernels where the shape allows; 0: one-CTA kernels only
void tc_set_debug(int v);     // diagnostic knobs of the NT kernel (timing experiments only; see TcNTParams::debug)
