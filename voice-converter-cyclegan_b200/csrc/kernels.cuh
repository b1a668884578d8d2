// Internal kernel-launcher declarations for libcgvc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

// ---- "F16F8" operand planes (CGVC_PREC_F16F8; forward, data gradient and weight gradient since round 2): an fp32 tensor x is kept as
//        q16  = fp16(x)                                   2 bytes / element   (hi * hi product: one kind::f16 MMA)
//        q8hi = e4m3(sat(float(q16) * S_hi))              1 byte              } the two cross products hi * lo, lo * hi as
//        q8lo = e4m3(sat((x - float(q16)) * S_lo))        1 byte              } kind::f8f6f4 MMAs at twice the rate
//      with static power-of-two scales chosen so that BOTH cross products carry 2^15, which the first kind::f16 MMA of a
//      tile removes again (scale-input-d = 15):  activations S_hi = 1, S_lo = 2^12;  weights S_hi = 2^3, S_lo = 2^15.
//      Emulated end to end in tests/precision_study.py (scheme fp16_f8_static): 4.7e-5 on the generator output.
#define CGVC_Q_ACT_SHI 1.0f
#define CGVC_Q_ACT_SLO 4096.0f
#define CGVC_Q_W_SHI 8.0f
#define CGVC_Q_W_SLO 32768.0f
#define CGVC_Q_ACC_SHIFT 15
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t cgvc_e4m3x4(float a, float b, float c, float d) {      // 4 floats -> 4 saturating e4m3 bytes
  const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
// 4 consecutive values -> 8 bytes of q16, 4 bytes of q8hi, 4 bytes of q8lo
__device__ __forceinline__ void cgvc_quant4(const float (&v)[4], float s_hi, float s_lo, uint2& q16, uint32_t& q8hi, uint32_t& q8lo) {
  const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  q16.x = *reinterpret_cast<const uint32_t*>(&h01); q16.y = *reinterpret_cast<const uint32_t*>(&h23);
  q8hi = cgvc_e4m3x4(f01.x * s_hi, f01.y * s_hi, f23.x * s_hi, f23.y * s_hi);
  q8lo = cgvc_e4m3x4((v[0] - f01.x) * s_lo, (v[1] - f01.y) * s_lo, (v[2] - f23.x) * s_lo, (v[3] - f23.y) * s_lo);
}
#endif

extern unsigned long long g_cgvc_launches;   // incremented by every kernel launch of the library

#define CGVC_MAX_TAPS 18   // largest filter on the path: discriminator d3, 6x3 (module.py:208)

// Geometry of a "gather-GEMM":  D[m, n] = sum_t sum_c  S[src(m,t), c] * Wt[t][c][n]
//   m enumerates a logical output grid (b, y, x); src(m,t) = (b, y*sy + oy[t], x*sx + ox[t]) in the source
//   tensor [B,Hs,Ws,*] (zero outside), and row m is written to (b, y*dsy + doy, x*dsx + dox) of the
//   destination tensor [B,Hd,Wd,*].  This one primitive expresses
//     - forward conv (TF SAME, any stride):  oy = i - pad_top, sy = stride
//     - data gradient, stride 1:             oy = pad_top - i (flipped taps), source = dY
//     - data gradient, stride 2:             one launch per output-parity class p with the taps of matching
//                                            parity, oy = (p + pad - i)/2, dsy = 2, doy = p
//   (module.py:22-64 of the reference).
struct GatherGeom {
  int B, Hy, Wx;
  int Hs, Ws;
  int sy, sx;
  int ntaps;
  int Hd, Wd, dsy, dsx, doy, dox;
  short oy[CGVC_MAX_TAPS], ox[CGVC_MAX_TAPS], widx[CGVC_MAX_TAPS];
};

struct GemmOperands {
  const float* src; int s_ld; int s_coff; int C;      // gathered operand: row stride, column offset, #channels contracted
  const float* w; long long w_ts; int w_cs; int w_ns; // weight element (tap slab, c, n) = w[widx*w_ts + c*w_cs + n*w_ns]
  int N;
  float* dst; int d_ld; int d_coff;
  const float* bias;                                  // [N] or null
  int accumulate;                                     // dst += result
};

// ---- fp32 SIMT path (reference arithmetic on the GPU; also the permanent path for the tiny-K layers)
cudaError_t launch_gg_simt(const GatherGeom& g, const GemmOperands& op, cudaStream_t st);
// weight gradient in forward geometry: dW[widx[t]][c][n] += sum_m S[src(m,t), c] * G[m, n]   (atomic accumulate)
cudaError_t launch_wgrad_simt(const GatherGeom& g, const float* src, int s_ld, int s_coff, int C,
                              const float* grad, int g_ld, int g_coff, int N,
                              float* dw, long long w_ts, int w_cs, int w_ns, cudaStream_t st);
// db[n] += sum_m G[m, g_coff + n]
cudaError_t launch_colsum(const float* grad, long long rows, int g_ld, int g_coff, int N, float* db, cudaStream_t st);

// ---- instance-norm / GLU / residual "post" kernels (module.py:3-20, 66-146)
struct PostParams {
  const float* p; int ldp; int Cc;      // conv output rows [B*(R/sh), ldp]; 'a' cols [0,Cc), gate cols [Cc,2Cc); Cc = C*sh
  int B, R, C, sh;                      // after the pixel-shuffle view: R positions x C channels per sample
  const float *beta_a, *gamma_a, *beta_g, *gamma_g;
  int has_in, has_gate;
  const float* resid;                   // [B,R,C] added to the result (residual1d_block), or null
  float* y;                             // [B,R,C] (may be null when only the planes are wanted)
  float* stats;                         // [B,4,C]: mean_a, rstd_a, mean_g, rstd_g (written if has_in)
  __nv_bfloat16 *y_hi, *y_lo;           // optional bf16 split planes of y for the tensor-core path
  float* scratch;                       // [B,4,C] fp32 workspace for the instance-norm sums (null: internal buffer, single-stream use only)
  int qmode;                            // 1: the planes are F16F8 planes instead: y_hi = q16 [B*R*C halves], y_lo = q8hi [B*R*C bytes] followed by q8lo
};
cudaError_t launch_post_fwd(const PostParams& pp, cudaStream_t st);

struct PostBwdParams {
  const float* dy1; const float* dy2;   // upstream gradient(s) [B,R,C]; dy2 may be null (summed if present)
  const float* p; int ldp; int Cc;
  int B, R, C, sh;
  const float *beta_a, *gamma_a, *beta_g, *gamma_g;
  int has_in, has_gate;
  const float* stats;
  float* dp;                            // same layout as p (fp32), may be null if only planes wanted
  __nv_bfloat16 *dp_hi, *dp_lo;         // optional bf16 split planes, same layout as p
  float *dbeta_a, *dgamma_a, *dbeta_g, *dgamma_g;   // accumulated atomically (may be null when has_in == 0)
  float *dbias_a, *dbias_g;             // conv-bias gradients [Cc] = column sums of dp (accumulated atomically; may be null)
  int qmode;                            // 1: dp_hi / dp_lo are F16F8 planes (q16; q8hi followed by q8lo) with the activation-role scales
  float* scratch;                       // [B,4,C] fp32 workspace (null: internal buffer, single-stream use only)
};
cudaError_t launch_post_bwd(const PostBwdParams& pp, cudaStream_t st);
void post_set_stream(int on);      // 1 (default): gated layers without shuffle and 32 / 48 / 64 positions per sample take the streaming (cp.async double-buffered) form of it
cudaError_t post_init_kernels();   // shared-memory opt-in of the streaming kernels (call once, outside any stream capture)
void post_set_onepass(int on);     // 1 (default): samples of <= 64 positions take the one-pass backward kernel; 0: always sums + apply

// ---- discriminator head: dense(1024->1) + sigmoid (module.py:211) and LSGAN loss (model.py:68-69,81-86)
cudaError_t launch_head_fwd(const float* y, long long rows, int C, const float* w, const float* b, float* prob, cudaStream_t st);
// loss_slot += coef * mean((p - target)^2) over `rows`; dz = coef * 2 (p-target)/rows * p (1-p);
// dy[row,:] = dz * w (if dy);  dw += sum dz*y[row,:], db += sum dz (if dw)
cudaError_t launch_head_loss_bwd(const float* prob, const float* y, long long rows, int C, const float* w,
                                 float target, float coef, float* loss_slot,
                                 float* dy, float* dw, float* db, cudaStream_t st, float grad_mult = 1.f);

// ---- L1 loss + gradient (utils.py:6-8): loss_slot += mean|yhat - y|; d[i] = gscale * sign(yhat - y)/n  (accumulate optional)
cudaError_t launch_l1_loss_grad(const float* yhat, const float* y, long long n, float* loss_slot,
                                const float* gscale_dev, float* d, int accumulate, cudaStream_t st, float grad_mult = 1.f);
// grad_mult (both loss kernels): the gradients -- not the loss values -- are multiplied by it: the loss scale of the F16F8 gradient planes
cudaError_t launch_scale(float* x, long long n, float a, cudaStream_t st);
// [B,F,T] <-> [B,T,F]
cudaError_t launch_transpose_ft(const float* in, float* out, int B, int F, int T, cudaStream_t st);
// y = a + b
cudaError_t launch_add(const float* a, const float* b, float* y, long long n, cudaStream_t st);

// ---- TF-style Adam over a flat range (tf.train.AdamOptimizer)
// hyper_dev: [0] = lr_t (already bias-corrected step size), [1] = grad_scale
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n,
                        const float* hyper_dev, float beta1, float beta2, float eps, cudaStream_t st);
// final loss algebra (model.py:72,83,88,90) on the 8-slot buffer
cudaError_t launch_finalize_losses(float* losses8, const float* lambdas_dev, cudaStream_t st);

// fp32 -> bf16 hi/lo split planes (x ~= hi + lo, |x - hi - lo| <= 2^-17 |x|)
cudaError_t launch_split_bf16(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t st);

// ---- single-input-channel specials (discriminator h1, K = 9, HBM-bound)
// dW[t][0][n] += sum_m x[src(m,t)] * G[m,n]; columns [0,n_split) -> dw_a, the rest -> dw_g; db = column sums (optional)
cudaError_t launch_wgrad_c1(const GatherGeom& g, const float* src, const float* grad, int g_ld, int N,
                            float* dw_a, float* dw_g, int n_split, float* db_a, float* db_g, cudaStream_t st);
// dx[B,H,W] = conv-transpose of G [rows, C] with w = [wa | wg] ([taps][c_split], [taps][C - c_split]); Z is scratch [rows, taps]
cudaError_t launch_dgrad_c1(const float* G, int C, const float* wa, const float* wg, int c_split, float* Z, float* dx,
                            int B, int H, int W, int kh, int kw, int sh, int sw, cudaStream_t st);
// fp32 [M, C] (row stride ld) -> zero-padded bf16 hi/lo planes [M, Cpad]
cudaError_t launch_pad_split(const float* x, long long M, int C, int ld, int Cpad, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);
// same into F16F8 planes: q16 [M, Cpad] halves, q8 = [M*Cpad bytes of q8hi | M*Cpad bytes of q8lo] (activation scales)
cudaError_t launch_pad_split_q(const float* x, long long M, int C, int ld, int Cpad, void* q16, void* q8, cudaStream_t st);
// tap lowering of the generator's 15-tap edge layers (simt_kernels.cu): im2col of a narrow channels-last tensor over the taps of a
// stride-1 1-D TF-SAME convolution into operand planes [M, Cpad] (qmode 1: F16F8 planes q16 / q8hi|q8lo, else bf16 hi / lo), dir = +1:
// out[m, t*C + c] = x[m + t - pl, c], dir = -1: x[m - t + pl, c] (zero outside the sample, pl = (kw - 1) / 2), and the matching sum
// y[m, c] = bias[c] + sum_t z[m + dir*(t - pl), t*C + c]
cudaError_t launch_im2col_taps(const float* x, long long M, int T, int C, int kw, int dir, int Cpad, int qmode, void* hi, void* lo, cudaStream_t st);
cudaError_t launch_col2im_taps(const float* z, int ldz, long long M, int T, int C, int kw, int dir, const float* bias, float* y, cudaStream_t st);
// P[m, 0:2*cout] = [bias_a | bias_g] + sum_t x[src(m,t)] * [wa | wg][t]   (single input channel, TF kernels [taps][1][cout])
cudaError_t launch_conv_c1_fwd(const GatherGeom& g, const float* x, const float* wa, const float* wg, const float* ba, const float* bg,
                               int cout, float* P, cudaStream_t st);

// ... and with the layer's GLU in the same pass (gate without instance norm): also y = a * sigmoid(g) as fp32 [M, cout] (optional) and as
// operand planes (bf16 hi / lo, or with qmode the F16F8 planes q16; q8hi followed by q8lo); <= 9 taps
cudaError_t launch_conv_c1_glu_fwd(const GatherGeom& g, const float* x, const float* wa, const float* wg, const float* ba, const float* bg,
                                   int cout, float* P, float* y, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, int qmode, cudaStream_t st);

// s[off .. off+n) = v6_host[0..n)  (n <= 6), passed by value in the kernel arguments (no host-memory copy node)
cudaError_t launch_set_scalars(float* s, int off, int n, const float* v6_host, cudaStream_t st);

// ---- device-resident training data: epoch plan (pairing + crops from a counter-based generator) and minibatch gather
// (train.py:90-107, preprocess.py:207-238; contract in simt_kernels.cu).  off_X: n_X + 1 frame prefix sums; plan: [4][min(n_A, n_B)] ints;
// err: one int, set to (utterance index + 1, bit 30 = side B) if an utterance is shorter than the crop
cudaError_t launch_sample_plan(const long long* off_A, int n_A, const long long* off_B, int n_B, unsigned long long seed, long long epoch,
                               int crop, int* plan, int* err, cudaStream_t st);
cudaError_t launch_gather_minibatch(const float* cA, const long long* off_A, const float* cB, const long long* off_B, const int* plan,
                                    int num_pairs, int first_pair, int batch, int F, int crop, float* out_A, float* out_B, cudaStream_t st);

// ---- discriminator input layer backward, fused (no instance norm, one input channel): dP = GLU backward of (dY, P = [a | g]) is
// formed in registers and consumed in place -- weight + bias gradients, or the data gradient (Z scratch [rows, taps]) -- instead of
// being written to HBM and read back.  C = channels per branch (128).
cudaError_t launch_glu_bwd_wgrad_c1(const GatherGeom& g, const float* src, const float* dy, const float* P, int C,
                                    float* dw_a, float* dw_g, float* db_a, float* db_g, cudaStream_t st);
cudaError_t launch_glu_bwd_dgrad_c1(const float* dy, const float* P, int C, const float* wa, const float* wg, float* Z, float* dx,
                                    int B, int H, int W, int kh, int kw, int sh, int sw, cudaStream_t st);
