/*
 * cgvc.h -- C ABI of libcgvc.so, the B200-native CycleGAN-VC training/inference engine.
 *
 * The reference (leimao/Voice-Converter-CycleGAN) has no FFI of its own: its seam is the Python
 * class `CycleGAN` (model.py:7-169) driving a TensorFlow-1 session.  Each entry point below names
 * the reference interface it replaces (file:line into /root/reference).  The Python mirror of the
 * reference class lives in voice-converter-cyclegan_b200/model.py and binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every call returns int: 0 = ok, <0 = error (message via cgvc_last_error); nothing throws across the ABI
 *   - pointers are DEVICE pointers unless the parameter name ends in _host
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls enqueue work and do
 *     not synchronise unless documented
 *   - one handle per device, not thread-safe per handle
 *   - storage (all arenas) is owned by the caller (torch tensors on the Python side); the engine never frees it
 *   - activations/IO are fp32; MCEP frames are [batch, 24, frames] row-major exactly like the reference's
 *     placeholders (model.py:35-42)
 */
#ifndef CGVC_H
#define CGVC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGVC_ABI_VERSION 1
#define CGVC_NUM_LOSSES 8   /* model.py:153-169, in that order: cycle, identity, G_A2B, G_B2A, generator, D_A, D_B, discriminator */

typedef struct cgvc_engine* cgvc_handle;

typedef struct cgvc_config {
  int num_features;   /* 24: model.py:9 num_features (only 24 is supported; the discriminator head needs H=24) */
  int max_batch;      /* largest per-GPU minibatch a train step / forward will be called with */
  int max_frames;     /* largest frame count T (multiple of 4, module.py:166-179); training uses 128 (train.py:24) */
  int precision;      /* CGVC_PREC_*: arithmetic of the tensor-core contractions */
  int device;         /* CUDA device ordinal */
  int train;          /* 1: size arenas for training (model.py mode='train'); 0: forward only */
} cgvc_config;

enum {
  CGVC_PREC_FP32_SIMT = 0,  /* every contraction in fp32 FFMA (reference arithmetic; slow, used as on-GPU cross-check) */
  CGVC_PREC_BF16X3 = 1,     /* tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate (~2^-16 rel error; parity mode) */
  CGVC_PREC_BF16 = 2,       /* tcgen05 single bf16 MMA (fast, NOT parity-grade) */
  CGVC_PREC_F16F8 = 3       /* fp16 hi*hi MMA + the two cross terms as e4m3 kind::f8f6f4 MMAs at twice the rate, their common power of
                             * two folded out by scale-input-d: 2 MMA units per product instead of 3, parity-grade (4.7e-5 on the
                             * generator output).  Forward and data gradient; the weight gradient -- a leaf of the graph, its rounding
                             * error is not propagated into any other tensor -- reads the fp16 planes alone (1 unit; option "wgrad_f16" = 0
                             * restores the 2-unit form).  Training applies a power-of-two loss scale to the gradient planes and
                             * removes it in Adam */
};

enum cgvc_arena {
  CGVC_ARENA_PARAM = 0,     /* fp32 trainable variables, TF layouts, order [G_A2B | G_B2A | D_A | D_B] (model.py:93-95) */
  CGVC_ARENA_GRAD = 1,      /* fp32 gradients, same layout (what `minimize` computes, model.py:107-108) */
  CGVC_ARENA_ADAM_M = 2,    /* Adam first moment  (<var>/Adam   slots of tf.train.AdamOptimizer) */
  CGVC_ARENA_ADAM_V = 3,    /* Adam second moment (<var>/Adam_1 slots) */
  CGVC_ARENA_WORK = 4,      /* activations saved for backward, gradient scratch, bf16 operand planes */
  CGVC_ARENA_COUNT = 5
};

/* -- lifecycle: replaces CycleGAN.__init__ / build_model / optimizer_initializer (model.py:9-30,32-108) -- */
int cgvc_abi_version(void);
int cgvc_create(const cgvc_config* cfg, cgvc_handle* out);
int cgvc_destroy(cgvc_handle h);
const char* cgvc_last_error(cgvc_handle h);    /* h may be NULL: last error of a failed cgvc_create */

/* Bytes the caller must provide for an arena (WORK depends on cfg.max_batch/max_frames/train). */
int cgvc_arena_bytes(cgvc_handle h, int arena, size_t* bytes);
int cgvc_bind_arena(cgvc_handle h, int arena, void* dev_ptr, size_t bytes);

/* Variable table: TF variable names -> arena offsets (elements) and shapes, for checkpoints and weight
 * injection; replaces tf.trainable_variables() / tf.train.Saver (model.py:21,93,140-150). */
int cgvc_param_count(cgvc_handle h, int* n_tensors, size_t* n_elements);
int cgvc_param_info(cgvc_handle h, int index, const char** name, size_t* offset, int* ndim, int shape_out[4]);

/* Must be called after the caller (re)writes the PARAM arena (init, checkpoint load): refreshes the engine's
 * derived bf16 operand copies.  Replaces sess.run(global_variables_initializer) / Saver.restore side effects. */
int cgvc_params_updated(cgvc_handle h, void* stream);
/* Reset Adam step count t (beta-power accumulators of both optimizers, model.py:107-108) */
int cgvc_set_adam_step(cgvc_handle h, long long t);
int cgvc_get_adam_step(cgvc_handle h, long long* t);

/* -- the hot path: replaces CycleGAN.train (model.py:110-125) ------------------------------------------
 * One G step + one D step from the same pre-update weights, then both Adam updates.
 *   A_dev, B_dev       [batch, 24, frames] fp32 real samples of domain A / B
 *   gen_A_dev/gen_B_dev optional outputs [batch, 24, frames]: generation_A / generation_B (model.py:112)
 *   losses_dev         8 fp32 scalars (CGVC_NUM_LOSSES order), pre-update values, local-batch means
 * With a communicator attached (cgvc_comm_init) gradients are sum-all-reduced over ranks and averaged
 * before Adam; losses stay local. */
int cgvc_train_step(cgvc_handle h, const float* A_dev, const float* B_dev, int batch, int frames,
                    float lambda_cycle, float lambda_identity, float lr_generator, float lr_discriminator,
                    float* gen_A_dev, float* gen_B_dev, float* losses_dev, void* stream);

/* Same graph, but stops after the backward pass (GRAD arena holds d generator_loss/d G-vars and
 * d discriminator_loss/d D-vars); no all-reduce, no Adam.  For parity tests against the oracle. */
int cgvc_compute_gradients(cgvc_handle h, const float* A_dev, const float* B_dev, int batch, int frames,
                           float lambda_cycle, float lambda_identity,
                           float* gen_A_dev, float* gen_B_dev, float* losses_dev, void* stream);

/* TF-style Adam on the bound arenas (model.py:107-108; tf.train.AdamOptimizer beta1=0.5): advances t by one.
 * grad_scale multiplies every gradient first (1/nranks after a sum-all-reduce). */
int cgvc_adam_step(cgvc_handle h, float lr_generator, float lr_discriminator, float grad_scale, void* stream);

/* -- replaces CycleGAN.test (model.py:128-137): one generator forward.  direction 0 = 'A2B', 1 = 'B2A';
 * anything else returns CGVC_ERR_DIRECTION ("Conversion direction must be specified.", model.py:135). */
int cgvc_generator_forward(cgvc_handle h, int direction, const float* in_dev, float* out_dev,
                           int batch, int frames, void* stream);
/* discriminator forward, which 0 = discriminator_A, 1 = discriminator_B: out [batch, 6, frames/16] (module.py:188-213) */
int cgvc_discriminator_forward(cgvc_handle h, int which, const float* in_dev, float* out_dev,
                               int batch, int frames, void* stream);

/* Debug/parity taps: copy a named layer-boundary activation of the most recent cgvc_generator_forward /
 * cgvc_discriminator_forward (channels-last, fp32) into out_dev.  The generator forward only keeps them when the option
 * "debug_taps" is 1 (default 0: the inference path writes neither fp32 layer outputs nor anything for a backward pass).  Names: h1_glu d1 d2 r1..r6 u1 u2 (generator),
 * h1_glu d1 d2 d3 (discriminator).  n_out receives the element count. */
int cgvc_debug_activation(cgvc_handle h, const char* name, float* out_dev, size_t capacity, size_t* n_out, void* stream);

/* -- device-resident training data (replaces the per-step host feed of train.py:90-107 / preprocess.py:207-238) ---------------
 * The caller uploads each speaker's normalised MCEP corpus once: utterance u as a row-major [num_features][len_u] block at element
 * num_features * offsets[u] of corpus_X_dev, offsets_X_dev = n_X + 1 frame prefix sums (int64).
 * cgvc_sample_plan draws one epoch: both utterance lists shuffled independently, truncated to num_pairs = min(n_A, n_B), one uniform
 * crop start per utterance, from a counter-based generator keyed by (seed, epoch) -- the exact contract is stated in
 * csrc/simt_kernels.cu and mirrored on the host by cgvc.preprocess.counter_sample_plan.  plan_dev = int[4][num_pairs]
 * (utt_A, start_A, utt_B, start_B); *err_dev becomes non-zero (utterance index + 1, bit 30 set for speaker B) if an utterance is
 * shorter than crop_frames (the reference asserts this, preprocess.py:217).
 * cgvc_gather_minibatch writes pairs [first_pair, first_pair + batch) as A_out / B_out [batch][num_features][crop_frames]. */
int cgvc_sample_plan(cgvc_handle h, const long long* offsets_A_dev, int n_A, const long long* offsets_B_dev, int n_B,
                     unsigned long long seed, long long epoch, int crop_frames, int* plan_dev, int* err_dev, void* stream);
int cgvc_gather_minibatch(cgvc_handle h, const float* corpus_A_dev, const long long* offsets_A_dev, const float* corpus_B_dev,
                          const long long* offsets_B_dev, const int* plan_dev, int num_pairs, int first_pair, int batch, int crop_frames,
                          float* A_out_dev, float* B_out_dev, void* stream);

/* -- multi-GPU (no counterpart in the reference, which is single-device: model.py:22) --------------------
 * One process per GPU.  Rank 0 obtains a 128-byte NCCL unique id, the host side distributes it, every rank
 * calls cgvc_comm_init; cgvc_train_step then does ONE fp32 sum-all-reduce of the GRAD arena per step. */
int cgvc_comm_unique_id(cgvc_handle h, void* id128_host);
int cgvc_comm_init(cgvc_handle h, const void* id128_host, int rank, int nranks);
int cgvc_comm_destroy(cgvc_handle h);
int cgvc_allreduce_grads(cgvc_handle h, void* stream);

/* -- measurement hooks (bench.py) ----------------------------------------------------------------------------
 * cgvc_kernel_launches: number of CUDA kernels this library has launched so far (process-wide).
 * cgvc_profile_enable(1) starts recording a CUDA-event pair around every tensor-core kernel launch;
 * cgvc_profile_collect synchronises and returns, per kernel class (arrays of 3: 0 = forward/data-gradient
 * gather-GEMM with the plain epilogue, 1 = weight-gradient gather-GEMM, 2 = forward gather-GEMM with the fused instance-norm
 * epilogue), the summed device time [ms], algorithmic FLOPs (2*M*N*K, counted once, whatever the bf16 split multiplies it by)
 * and launch count since the enable call. */
int cgvc_kernel_launches(unsigned long long* count);
/* options: "two_streams" (default 1): run the two symmetric halves of a train step on two internal streams; 0 enqueues
 * everything on the caller's stream (used while per-kernel timings are taken).
 * "fuse_in" (default 1): instance norm + GLU / + residual fused into the forward conv kernel's epilogue where the shape
 * allows (generator layers whose 128-row tiles hold whole samples); 0 always uses the separate streaming kernels.
 * "fuse_bwd" (default 0): GLU / instance-norm backward of the generator's residual stack fused into the epilogue of the
 * data-gradient kernel that produces its upstream gradient (one kernel per layer backward instead of three); correct and tested,
 * but measured ~1 % slower than the streaming kernels on B200, hence opt-in.
 * "side_wgrad" (default 0): the weight-gradient GEMMs of a train step run on a side stream per lane, off the critical path of the
 * data-gradient chain (needs two_streams; not combined with fuse_bwd).  Correct and tested; measured neutral on a power-capped B200.
 * "pipelined_comm" (default 1): with a communicator attached, cgvc_train_step all-reduces the gradients network by network on a
 * communication stream and runs Adam + the weight-plane refresh of each network as soon as its all-reduce is done (0: one all-reduce
 * of the whole arena, then Adam).
 * "fuse_c1" (default 1): the discriminator's input layer (one input channel, K = 9, gate without instance norm) as fused HBM-bound kernels:
 * forward = convolution + GLU in one pass (the pre-gate outputs are written for the backward pass but not read back by a second kernel),
 * backward = the GLU backward recomputed inside its weight-gradient / data-gradient kernels instead of a dP tensor written to and read
 * from HBM.
 * "edge_lower" (default 1): the generator's two 15-tap layers with 24 channels on one side (h1: 24 -> 2 x 128, o1: 256 -> 24;
 * module.py:85-86,148) as dense 1 x 1 GEMMs -- h1 over the im2col of the 24-channel input (K = 360), o1 with its taps folded into the
 * output columns (N = 360) followed by the tap-shifted sum; forward, data gradient and weight gradient.  0 = 15-tap gather-GEMMs with the
 * 24-channel side padded to a 64 / 128-channel line per tap (forward / data gradient on a 32-wide tile).
 * "wgrad_f16" (CGVC_PREC_F16F8 only, default 1): weight-gradient GEMMs from the fp16 planes alone, one MMA unit per product; every
 * gradient tensor stays within 3.6e-4 of the float64 oracle (tolerance 1e-3; 1.8e-4 with 0 = fp16 + two e4m3 cross terms, 2 units).
 * "prep_batched" (CGVC_PREC_F16F8 only, default 1, process-wide): the fp16 + e4m3 weight planes of all layers (forward and data-gradient
 * layouts, biases) are rebuilt after Adam by ONE kernel that walks a job table in 32 x 64 tiles (one per network range in the pipelined
 * data-parallel schedule); 0 = three small kernels per layer branch (~210 launches per step).  Bit-identical planes.
 * "post_onepass" (default 1, process-wide): GLU / instance-norm backward of samples with <= 64 positions in one kernel that keeps the
 * sample's rows in registers (reads dY and the pre-norm outputs once); 0 = always the sums + apply kernel pair.
 * "post_stream" (default 1, process-wide): gated layers without pixel shuffle whose samples have 32, 48 or 64 positions take the streaming
 * form of the one-pass GLU / instance-norm backward: persistent CTAs walk (sample, channel block) items through a cp.async double buffer
 * in shared memory instead of holding a sample's rows in registers (needs post_onepass = 1).
 * "cta_pairs" (default 1, process-wide): tensor-core kernels on CTA pairs (tcgen05 cta_group::2, TMA im2col for the gathered
 * operand) where the shape allows; 0 = the one-CTA kernels everywhere.
 * "debug_taps" (default 0): see cgvc_debug_activation.
 * "tc_debug" (default 0): timing-experiment knobs of the forward/data-gradient kernel (results become garbage):
 * 1 = epilogue skips global stores, 2 = also skips TMEM loads, 4 = producers skip the activation gather. */
int cgvc_set_option(cgvc_handle h, const char* name, int value);
int cgvc_profile_enable(int on);
int cgvc_profile_collect(double* ms3, double* flops3, long long* launches3);
/* every recorded tensor-core launch in launch order: ms[i], flops[i], meta4[4i..4i+3] = (class, M rows, N columns,
 * K = taps * channels); *n_out = launches recorded (may exceed capacity; only `capacity` entries are written). */
int cgvc_profile_launches(double* ms, double* flops, long long* meta4, int capacity, int* n_out);

/* -- per-kernel entry points (unit parity against the oracle's primitives) -------------------------------
 * cgvc_conv_forward: channels-last TF-'SAME' cross-correlation (module.py:22-64), y = conv(x, w) + bias.
 *   x [B,H,W,Cin], w [kh,kw,Cin,Cout] (TF layout), y [B,Ho,Wo,Cout]; 1-D convs use H = kh = 1.
 *   precision: CGVC_PREC_*; shapes the tensor-core path cannot take fall back to CGVC_ERR_UNSUPPORTED. */
int cgvc_conv_forward(cgvc_handle h, int precision, const float* x, const float* w, const float* bias, float* y,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, void* stream);
/* gradients of the above: dx (may be NULL), dw and dbias are ACCUMULATED into (like the GRAD arena). */
int cgvc_conv_backward(cgvc_handle h, int precision, const float* x, const float* w, const float* dy,
                       float* dx, float* dw, float* dbias,
                       int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, void* stream);
/* fused instance-norm (+ optional pixel-shuffle view) + GLU (module.py:3-20,85-146):
 *   p [B, R/shuffle, 2*C*shuffle]: conv outputs, 'a' branch in columns [0,C*shuffle), gates after;
 *   y [B, R, C] = IN(a; beta_a, gamma_a) * sigmoid(IN(g; beta_g, gamma_g)); stats [B,4,C] = mean_a,rstd_a,mean_g,rstd_g */
int cgvc_in_glu_forward(cgvc_handle h, const float* p, const float* beta_a, const float* gamma_a,
                        const float* beta_g, const float* gamma_g, float* y, float* stats,
                        int B, int R, int C, int shuffle, void* stream);
int cgvc_in_glu_backward(cgvc_handle h, const float* dy, const float* p, const float* stats,
                         const float* beta_a, const float* gamma_a, const float* beta_g, const float* gamma_g,
                         float* dp, float* dbeta_a, float* dgamma_a, float* dbeta_g, float* dgamma_g,
                         int B, int R, int C, int shuffle, void* stream);

/* error codes */
enum {
  CGVC_OK = 0,
  CGVC_ERR_ARG = -1,
  CGVC_ERR_CUDA = -2,
  CGVC_ERR_UNBOUND = -3,      /* an arena needed by the call is not bound / too small */
  CGVC_ERR_DIRECTION = -4,    /* model.py:135 */
  CGVC_ERR_UNSUPPORTED = -5,
  CGVC_ERR_NCCL = -6
};

#ifdef __cplusplus
}
#endif
#endif /* CGVC_H */
