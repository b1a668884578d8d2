// libcgvc.so engine: parameter table, workspace planner, forward/backward schedules and the C ABI.
//
// Replaces the TensorFlow-1 session behind CycleGAN.train/test (model.py:110-137 of /root/reference):
// the graph wiring below follows model.py:44-90, the network shapes module.py:148-213.
#include "../../include/cgvc.h"
#include "kernels.cuh"
#include "tc_gemm.cuh"
#include "geom.h"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

static thread_local std::string g_create_error;

#define ADAM_B1 0.5f        // model.py:107-108
#define ADAM_B2 0.999f
#define ADAM_EPS 1e-8f

// ---------------------------------------------------------------------------------------------------
// parameter table (TF variable names / creation order)
// ---------------------------------------------------------------------------------------------------
struct TensorInfo { std::string name; size_t off; int ndim; int shape[4]; size_t numel; };
struct ConvW { size_t k, b; int kh, kw, cin, cout; };
struct InW { size_t beta, gamma; int c; };
struct Gated { ConvW a, g; InW ina, ing; int has_in; int sh, sw; int shuffle; int tc_slot; };
struct ResBlock { Gated h1; ConvW h2; InW in2; int tc_slot2; };
struct GenNet { Gated h1; Gated d[2]; ResBlock r[6]; Gated u[2]; ConvW o1; int o1_slot; size_t begin, end;
                int h1c_slot, o1f_slot; };      // the tap-lowered forms of the two 15-tap edge layers (`edge_lower`, see edge_on)
struct DiscNet { Gated h1; Gated d[3]; size_t dense_k, dense_b; size_t begin, end; };

// per-layer activations kept for backward
struct GLAct { float* P; float* stats; float* Y; __nv_bfloat16 *Yhi, *Ylo; };
struct GenActs {
  int n, T;
  const float* x_cl; __nv_bfloat16 *xhi, *xlo;
  __nv_bfloat16 *xchi, *xclo;   // im2col of the input over h1's taps: operand planes [n*T, ru128(kw*F)] (edge_lower)
  float* z;                     // o1's per-tap products [n*T, kw*F] before the tap-shifted sum (edge_lower)
  GLAct h1, d[2];
  struct { GLAct a; float *Pb, *sb, *Yr; __nv_bfloat16 *Yrhi, *Yrlo; } r[6];
  GLAct u[2];
  float* out_cl;
  float* post;                  // scratch for the instance-norm sums [n,4,1024] (null: the kernels' own lazily grown buffer)
};
struct DiscActs { int n, T; const float* x; GLAct h1, d[3]; float* prob; float* post; };

struct GraphKey {
  int batch, frames, id_off, kind, lanes, fuse;
  bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
};
struct GraphEntry { cudaGraphExec_t exec; unsigned long long launches; };

struct Bump {
  char* base = nullptr; size_t cap = 0, off = 0; bool overflow = false;
  void reset(void* b, size_t c) { base = (char*)b; cap = c; off = 0; overflow = false; }
  template <class T> T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) { overflow = true; off += bytes; return (T*)base; }
    T* p = (T*)(base + off); off += bytes; return p;
  }
};

// NCCL through dlopen: no link-time dependency, single-GPU use never touches it.
struct Id128 { char b[128]; };   // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

// Weight-gradient launches leave the critical path: a layer's backward is  (GLU / instance-norm backward -> dP planes) -> {weight gradient,
// data gradient}, and only the data gradient feeds the next layer.  With `on`, the weight-gradient GEMMs are enqueued on a side stream
// (a side branch of the captured graph) behind an event on their dP planes, so the tensor cores have them to run while the main chain
// is in its elementwise kernels.  The dP planes ping-pong between two buffers; a buffer is rewritten only after the weight gradient that
// read it has finished (`done`).
struct SideQ {
  cudaStream_t side = nullptr;
  cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  int cur = 0;
  bool on = false;
};

struct cgvc_engine {
  cgvc_config cfg;
  std::string err;
  std::vector<TensorInfo> tensors;
  size_t n_params = 0;        // arena length in elements (tensors padded to 16-byte boundaries)
  size_t n_real_params = 0;   // 119,787,058 trainable scalars
  GenNet gen[2];     // 0 = generator_A2B, 1 = generator_B2A
  DiscNet disc[2];   // 0 = discriminator_A, 1 = discriminator_B
  void* arena[CGVC_ARENA_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t arena_bytes[CGVC_ARENA_COUNT] = {0, 0, 0, 0, 0};
  long long adam_t = 0;
  // small device buffers owned by the engine
  float* d_scalars = nullptr;   // [0..1] lambdas, [2..3] adam hyper G (lr_t, gscale), [4..5] adam hyper D, [8..15] losses
  // tensor-core weight planes (owned; derived from PARAM)
  TcWeights tcw;
  // communicator
  NcclApi nccl; void* comm = nullptr; int rank = 0, nranks = 1;
  // the two lanes of a training step run on their own streams (forked from / joined into the caller's stream)
  cudaStream_t lane_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  int use_graphs = 1;           // replay the step as a CUDA graph (disabled automatically if capture is not possible)
  std::map<GraphKey, GraphEntry> graphs;
  float* stage = nullptr;       // [2][max_batch,num_features,max_frames]: fixed-address copies of the step's inputs for the graphs
  cudaStream_t graph_stream = nullptr; cudaEvent_t ev_bridge = nullptr, ev_bridge2 = nullptr;
  int fuse_in = 1;              // fuse instance norm (+GLU / +residual) into the forward GEMM epilogue where the shape allows
  int debug_taps = 0;           // cgvc_generator_forward also writes the fp32 copy of every layer output (cgvc_debug_activation)
  int fuse_bwd = 0;             // fuse the instance-norm (+GLU) backward into the upstream data-gradient GEMM's epilogue likewise.
                                // Off by default: measured 69.0 ms/step with it vs 68.2 without (profiles/r01_bench_v9_fusebwd*.json) --
                                // the 4 epilogue warps need 3x the tile's MMA time for it, and unlike the streaming kernels that
                                // work cannot overlap the other lane's tensor-core kernels
  // data-parallel step: the gradient all-reduce runs per network on its own stream; Adam and the weight-plane refresh of a network
  // start as soon as its all-reduce has finished, while the next network's is still on the wire
  cudaStream_t comm_stream = nullptr; cudaEvent_t ev_grads = nullptr, ev_ar[4] = {nullptr, nullptr, nullptr, nullptr};
  int pipelined_comm = 1;
  int fuse_c1 = 1;              // discriminator input layer backward: GLU backward fused into its weight / data gradient kernels
  int edge_lower = 1;           // the generator's 15-tap, 24-channel edge layers as dense 1 x 1 GEMMs (taps moved into the channel / column dimension)
  int two_streams = 1;          // 0: both lanes are enqueued on the caller's stream (clean per-kernel timing for profiling)
  int side_wgrad = 0;           // weight-gradient GEMMs on a side stream per lane (see SideQ); needs two_streams, excludes fuse_bwd.  Off by default:
                                // measured neutral under the 1 kW power cap (60.6-60.8 vs 60.5 ms/step, profiles/r02_bench_ab_*.json) -- the step is
                                // energy-bound there, so re-ordering work does not shorten it
  SideQ sideq[2];
  // debug taps of the last forward
  std::map<std::string, std::pair<const float*, size_t>> taps;

  float* P() const { return (float*)arena[CGVC_ARENA_PARAM]; }
  float* G() const { return (float*)arena[CGVC_ARENA_GRAD]; }
};

static int fail(cgvc_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                   \
  do { cudaError_t _e = (call);                                                                    \
       if (_e != cudaSuccess) return fail(e, CGVC_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); } while (0)
#define RET(call) do { int _r = (call); if (_r != 0) return _r; } while (0)

// Every entry point runs on the engine's device and leaves the caller's current device as it found it (a single process may
// drive several GPUs through torch, whose current device must not change behind its back).
struct DeviceGuard {
  int prev = -1;
  cudaError_t set(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev == dev) { prev = -1; return cudaSuccess; }
    return cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- table construction ------------------------------------------------------------------------------
struct TableBuilder {
  std::vector<TensorInfo>& t; size_t off = 0; std::string scope;
  size_t real = 0;
  size_t add(const std::string& name, std::initializer_list<int> shp) {
    off = (off + 3) & ~(size_t)3;     // every tensor starts 16-byte aligned (float4 / red.v4 / cp.async paths rely on it)
    TensorInfo ti; ti.name = scope + "/" + name; ti.off = off; ti.ndim = (int)shp.size(); ti.numel = 1;
    int i = 0; for (int s : shp) { ti.shape[i++] = s; ti.numel *= (size_t)s; }
    for (; i < 4; ++i) ti.shape[i] = 1;
    t.push_back(ti); off += ti.numel; real += ti.numel; return ti.off;
  }
  ConvW conv1d(const std::string& name, int k, int cin, int cout) {
    ConvW c; c.kh = 1; c.kw = k; c.cin = cin; c.cout = cout;
    c.k = add(name + "/kernel", {k, cin, cout}); c.b = add(name + "/bias", {cout}); return c;
  }
  ConvW conv2d(const std::string& name, int kh, int kw, int cin, int cout) {
    ConvW c; c.kh = kh; c.kw = kw; c.cin = cin; c.cout = cout;
    c.k = add(name + "/kernel", {kh, kw, cin, cout}); c.b = add(name + "/bias", {cout}); return c;
  }
  InW inorm(int idx, int c) {
    std::string n = idx == 0 ? "InstanceNorm" : "InstanceNorm_" + std::to_string(idx);
    InW w; w.c = c; w.beta = add(n + "/beta", {c}); w.gamma = add(n + "/gamma", {c}); return w;
  }
};

static void build_generator(TableBuilder& tb, GenNet& g, int nf) {
  g.begin = tb.off;
  g.h1 = Gated{}; g.h1.a = tb.conv1d("h1_conv", 15, nf, 128); g.h1.g = tb.conv1d("h1_conv_gates", 15, nf, 128);
  g.h1.has_in = 0; g.h1.sh = 1; g.h1.sw = 1; g.h1.shuffle = 1;
  int idx = 0, cin = 128;
  const int dco[2] = {256, 512};
  for (int i = 0; i < 2; ++i) {
    std::string p = "downsample1d_block" + std::to_string(i + 1) + "_";
    Gated& L = g.d[i]; L = Gated{};
    L.a = tb.conv1d(p + "h1_conv", 5, cin, dco[i]); L.ina = tb.inorm(idx++, dco[i]);
    L.g = tb.conv1d(p + "h1_gates", 5, cin, dco[i]); L.ing = tb.inorm(idx++, dco[i]);
    L.has_in = 1; L.sh = 1; L.sw = 2; L.shuffle = 1; cin = dco[i];
  }
  for (int i = 0; i < 6; ++i) {
    std::string p = "residual1d_block" + std::to_string(i + 1) + "_";
    ResBlock& R = g.r[i]; R = ResBlock{};
    R.h1.a = tb.conv1d(p + "h1_conv", 3, 512, 1024); R.h1.ina = tb.inorm(idx++, 1024);
    R.h1.g = tb.conv1d(p + "h1_gates", 3, 512, 1024); R.h1.ing = tb.inorm(idx++, 1024);
    R.h1.has_in = 1; R.h1.sh = 1; R.h1.sw = 1; R.h1.shuffle = 1;
    R.h2 = tb.conv1d(p + "h2_conv", 3, 1024, 512); R.in2 = tb.inorm(idx++, 512);
  }
  cin = 512;
  const int uco[2] = {1024, 512};
  for (int i = 0; i < 2; ++i) {
    std::string p = "upsample1d_block" + std::to_string(i + 1) + "_";
    Gated& L = g.u[i]; L = Gated{};
    L.a = tb.conv1d(p + "h1_conv", 5, cin, uco[i]); L.ina = tb.inorm(idx++, uco[i] / 2);
    L.g = tb.conv1d(p + "h1_gates", 5, cin, uco[i]); L.ing = tb.inorm(idx++, uco[i] / 2);
    L.has_in = 1; L.sh = 1; L.sw = 1; L.shuffle = 2; cin = uco[i] / 2;
  }
  g.o1 = tb.conv1d("o1_conv", 15, 256, nf);
  g.end = tb.off;
}

static void build_discriminator(TableBuilder& tb, DiscNet& d) {
  d.begin = tb.off;
  d.h1 = Gated{}; d.h1.a = tb.conv2d("h1_conv", 3, 3, 1, 128); d.h1.g = tb.conv2d("h1_conv_gates", 3, 3, 1, 128);
  d.h1.has_in = 0; d.h1.sh = 1; d.h1.sw = 2; d.h1.shuffle = 1;
  int idx = 0, cin = 128;
  const int kh[3] = {3, 3, 6}, co[3] = {256, 512, 1024}, sh[3] = {2, 2, 1};
  for (int i = 0; i < 3; ++i) {
    std::string p = "downsample2d_block" + std::to_string(i + 1) + "_";
    Gated& L = d.d[i]; L = Gated{};
    L.a = tb.conv2d(p + "h1_conv", kh[i], 3, cin, co[i]); L.ina = tb.inorm(idx++, co[i]);
    L.g = tb.conv2d(p + "h1_gates", kh[i], 3, cin, co[i]); L.ing = tb.inorm(idx++, co[i]);
    L.has_in = 1; L.sh = sh[i]; L.sw = 2; L.shuffle = 1; cin = co[i];
  }
  d.dense_k = tb.add("dense/kernel", {1024, 1}); d.dense_b = tb.add("dense/bias", {1});
  d.end = tb.off;
}

// ---- conv building blocks (dispatch: tcgen05 where the layer is registered, else fp32 SIMT) ---------------------
struct ConvIO {               // one convolution application
  const float* x; const __nv_bfloat16 *xhi, *xlo;   // input [n,H,W,Cin] fp32 (may be null on the tensor-core path) + bf16 planes
  int n, H, W;
};

static int conv_out_dims(const ConvW& c, int sh, int sw, int H, int W, int& Ho, int& Wo) {
  int p; same_pad(H, c.kh, sh, p, Ho); same_pad(W, c.kw, sw, p, Wo); return 0;
}

// y[., coff:coff+cout] = conv(x, w) + b  into a row-major [rows, ld] buffer
static int conv_fwd_simt(cgvc_engine* e, const float* Pm, const ConvW& c, int sh, int sw, const ConvIO& io,
                         float* dst, int ld, int coff, cudaStream_t st) {
  if (!io.x) return fail(e, CGVC_ERR_UNSUPPORTED, "fp32 activations were not kept for a layer that fell back to the SIMT path");
  GatherGeom g = fwd_geom(io.n, io.H, io.W, c.kh, c.kw, sh, sw);
  GemmOperands op; memset(&op, 0, sizeof op);
  op.src = io.x; op.s_ld = c.cin; op.s_coff = 0; op.C = c.cin;
  op.w = Pm + c.k; op.w_ts = (long long)c.cin * c.cout; op.w_cs = c.cout; op.w_ns = 1; op.N = c.cout;
  op.dst = dst; op.d_ld = ld; op.d_coff = coff; op.bias = Pm + c.b; op.accumulate = 0;
  CK(launch_gg_simt(g, op, st));
  return 0;
}

// dx (+)= dgrad(dy[., coff:coff+cout], w)
static int conv_dgrad_simt(cgvc_engine* e, const float* Pm, const ConvW& c, int sh, int sw, int n, int H, int W,
                           const float* dy, int ld, int coff, float* dx, int accumulate, cudaStream_t st) {
  if (!dy) return fail(e, CGVC_ERR_UNSUPPORTED, "fp32 gradients were not kept for a layer that fell back to the SIMT path");
  std::vector<GatherGeom> gs = dgrad_geoms(n, H, W, c.kh, c.kw, sh, sw);
  for (const GatherGeom& g : gs) {
    GemmOperands op; memset(&op, 0, sizeof op);
    op.src = dy; op.s_ld = ld; op.s_coff = coff; op.C = c.cout;
    op.w = Pm + c.k; op.w_ts = (long long)c.cin * c.cout; op.w_cs = 1; op.w_ns = c.cout; op.N = c.cin;
    op.dst = dx; op.d_ld = c.cin; op.d_coff = 0; op.bias = nullptr; op.accumulate = accumulate;
    CK(launch_gg_simt(g, op, st));
  }
  return 0;
}

// dW += x^T * dy (forward geometry); the bias gradient comes from the IN/GLU backward kernel (or launch_colsum for o1)
static int conv_wgrad_simt(cgvc_engine* e, float* Gm, const ConvW& c, int sh, int sw, const ConvIO& io,
                           const float* dy, int ld, int coff, cudaStream_t st) {
  if (!io.x || !dy) return fail(e, CGVC_ERR_UNSUPPORTED, "fp32 tensors were not kept for a layer that fell back to the SIMT path");
  GatherGeom g = fwd_geom(io.n, io.H, io.W, c.kh, c.kw, sh, sw);
  CK(launch_wgrad_simt(g, io.x, c.cin, 0, c.cin, dy, ld, coff, c.cout, Gm + c.k, (long long)c.cin * c.cout, c.cout, 1, st));
  return 0;
}

// Loss scale of the F16F8 gradient planes: gradients shrink as 1 / batch, and their fp16 + e4m3 planes have a
// window of about 8 binades in which the result does not depend on the scale; 2^(9 + floor(log2(batch))) sits in its middle
// (2^10 at batch 2 ... 2^17 at batch 256).  Applied where the loss gradients are formed, removed by Adam's grad_scale.  1 otherwise.
static float loss_scale(const cgvc_engine* e, int batch) {
  if (e->cfg.precision != CGVC_PREC_F16F8) return 1.f;
  int l = 0; while ((2 << l) <= batch && l < 9) ++l;       // floor(log2(batch)), capped
  return ldexpf(1.f, 9 + l);
}

static bool tc_enabled(const cgvc_engine* e) { return e->cfg.precision != CGVC_PREC_FP32_SIMT && e->tcw.ready; }
static bool use_tc(const cgvc_engine* e, int slot) { return slot >= 0 && tc_enabled(e); }

// gated layer: conv_a || conv_g -> P [rows, 2*cout]
static int gated_conv_fwd(cgvc_engine* e, const Gated& L, const ConvIO& io, float* P, cudaStream_t st) {
  if (use_tc(e, L.tc_slot) && io.xhi) {
    int r = tc_conv_fwd(e->tcw, L.tc_slot, e->cfg.precision, io.xhi, io.xlo, io.n, io.H, io.W, L.sh, L.sw, P, st);
    if (r == 0) return 0;
    if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc_conv_fwd failed: %s", cudaGetErrorString((cudaError_t)r));
  }
  if (L.a.cin == 1 && io.x && L.a.cout % 4 == 0 && 256 % (L.a.cout / 2) == 0) {   // discriminator h1: HBM-bound special
    GatherGeom g = fwd_geom(io.n, io.H, io.W, L.a.kh, L.a.kw, L.sh, L.sw);
    const float* Pm = e->P();
    CK(launch_conv_c1_fwd(g, io.x, Pm + L.a.k, Pm + L.g.k, Pm + L.a.b, Pm + L.g.b, L.a.cout, P, st));
    return 0;
  }
  RET(conv_fwd_simt(e, e->P(), L.a, L.sh, L.sw, io, P, 2 * L.a.cout, 0, st));
  RET(conv_fwd_simt(e, e->P(), L.g, L.sh, L.sw, io, P, 2 * L.a.cout, L.a.cout, st));
  return 0;
}

static int gated_conv_dgrad(cgvc_engine* e, const Gated& L, int n, int H, int W, const float* dP,
                            const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, float* dx, int accumulate, cudaStream_t st) {
  if (use_tc(e, L.tc_slot) && dPhi) {
    int r = tc_conv_dgrad(e->tcw, L.tc_slot, e->cfg.precision, dPhi, dPlo, n, H, W, L.sh, L.sw, dx, accumulate, st);
    if (r == 0) return 0;
    if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc_conv_dgrad failed: %s", cudaGetErrorString((cudaError_t)r));
  }
  RET(conv_dgrad_simt(e, e->P(), L.a, L.sh, L.sw, n, H, W, dP, 2 * L.a.cout, 0, dx, accumulate, st));
  RET(conv_dgrad_simt(e, e->P(), L.g, L.sh, L.sw, n, H, W, dP, 2 * L.a.cout, L.a.cout, dx, 1, st));
  return 0;
}

static int gated_conv_wgrad(cgvc_engine* e, const Gated& L, const ConvIO& io, const float* dP,
                            const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, cudaStream_t st) {
  if (use_tc(e, L.tc_slot) && dPhi && io.xhi) {
    int r = tc_conv_wgrad(e->tcw, L.tc_slot, e->cfg.precision, io.xhi, io.xlo, dPhi, dPlo, io.n, io.H, io.W, L.sh, L.sw,
                          e->G() + L.a.k, e->G() + L.g.k, nullptr, nullptr, st);
    if (r == 0) return 0;
    if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc_conv_wgrad failed: %s", cudaGetErrorString((cudaError_t)r));
  }
  RET(conv_wgrad_simt(e, e->G(), L.a, L.sh, L.sw, io, dP, 2 * L.a.cout, 0, st));
  RET(conv_wgrad_simt(e, e->G(), L.g, L.sh, L.sw, io, dP, 2 * L.a.cout, L.a.cout, st));
  return 0;
}

static PostParams post_params(const cgvc_engine* e, const Gated& L, const GLAct& A, int n, int rows_per_sample_out, bool keep_y, float* scratch) {
  PostParams q; memset(&q, 0, sizeof q);
  q.scratch = scratch;
  const float* Pm = e->P();
  q.p = A.P; q.ldp = 2 * L.a.cout; q.Cc = L.a.cout; q.B = n; q.sh = L.shuffle;
  q.R = rows_per_sample_out * L.shuffle; q.C = L.a.cout / L.shuffle;
  q.has_in = L.has_in; q.has_gate = 1;
  if (L.has_in) { q.beta_a = Pm + L.ina.beta; q.gamma_a = Pm + L.ina.gamma; q.beta_g = Pm + L.ing.beta; q.gamma_g = Pm + L.ing.gamma; }
  q.y = (keep_y || !A.Yhi) ? A.Y : nullptr;      // without planes the fp32 activation is the only copy
  q.stats = L.has_in ? A.stats : nullptr; q.y_hi = A.Yhi; q.y_lo = A.Ylo;
  q.qmode = e->cfg.precision == CGVC_PREC_F16F8;
  return q;
}

// gated layer forward with instance norm + GLU fused into the GEMM epilogue when the tensor-core path can (1-D layer whose
// 128-row tiles hold whole samples); otherwise conv kernel + the two streaming instance-norm kernels
static int gated_layer_forward(cgvc_engine* e, const Gated& L, const ConvIO& io, const GLAct& A, int n, int rows_per_sample_out,
                               bool keep_y, float* post_scratch, cudaStream_t st, bool save_pre = true) {
  const float* Pm = e->P();
  if (use_tc(e, L.tc_slot) && io.xhi && L.has_in && (L.shuffle == 1 || L.shuffle == 2) && io.H == 1 && A.Yhi && e->fuse_in) {
    TcFuse f; memset(&f, 0, sizeof f);
    f.R = rows_per_sample_out;
    f.gamma_a = Pm + L.ina.gamma; f.beta_a = Pm + L.ina.beta; f.gamma_g = Pm + L.ing.gamma; f.beta_g = Pm + L.ing.beta;
    f.stats = save_pre ? A.stats : nullptr; f.y = keep_y ? A.Y : nullptr; f.y_hi = A.Yhi; f.y_lo = A.Ylo;
    bool fused = false;
    int r = tc_conv_fwd_fused(e->tcw, L.tc_slot, e->cfg.precision, io.xhi, io.xlo, io.n, io.H, io.W, L.sh, L.sw, save_pre ? A.P : nullptr, f, &fused, st);
    if (r != 0 && !save_pre)                                  // shape not fusable: the two-kernel path needs P as its intermediate
      r = tc_conv_fwd_fused(e->tcw, L.tc_slot, e->cfg.precision, io.xhi, io.xlo, io.n, io.H, io.W, L.sh, L.sw, A.P, f, &fused, st);
    if (r != 0 && r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc_conv_fwd_fused failed: %s", cudaGetErrorString((cudaError_t)r));
    if (r == 0 && fused) return 0;
    if (r == 0) {                                             // conv done, epilogue not fusable for this shape
      PostParams q = post_params(e, L, A, n, rows_per_sample_out, keep_y, post_scratch);
      CK(launch_post_fwd(q, st));
      return 0;
    }
  }
  RET(gated_conv_fwd(e, L, io, A.P, st));
  PostParams q = post_params(e, L, A, n, rows_per_sample_out, keep_y, post_scratch);
  CK(launch_post_fwd(q, st));
  return 0;
}

// ---- generator -------------------------------------------------------------------------------------------
// Tap lowering of the two 15-tap edge layers (module.py:85-86 h1, module.py:148 o1; kernels and rationale in simt_kernels.cu above
// im2col_taps_kernel): h1 becomes a 1 x 1 gated layer over the im2col of the 24-channel input (TF's [1,15,24,128] kernel is that
// [360,128] matrix as it lies in memory, so forward, data and weight gradient address the same PARAM / GRAD ranges), o1 a 1 x 1 layer
// with the taps folded into its output columns (TcLayer::fold) followed by the tap-shifted sum.
static inline int edge_cpad(int c) { return (c + 127) / 128 * 128; }      // operand-plane width of kw * F channels (a multiple of 128 serves both precisions)
static bool edge_on(const cgvc_engine* e, const GenNet& N) { return e->edge_lower && use_tc(e, N.h1c_slot) && use_tc(e, N.o1f_slot); }

static void plan_gated(Bump& ws, GLAct& a, long long rows_out, int cout2, int n, int Cstat, bool planes, long long y_elems) {
  a.P = ws.take<float>((size_t)rows_out * cout2);
  a.stats = ws.take<float>((size_t)n * 4 * Cstat);
  a.Y = ws.take<float>((size_t)y_elems);
  a.Yhi = a.Ylo = nullptr;
  if (planes) { a.Yhi = ws.take<__nv_bfloat16>((size_t)y_elems); a.Ylo = ws.take<__nv_bfloat16>((size_t)y_elems); }
}

static void plan_generator(cgvc_engine* e, Bump& ws, GenActs& A, int n, int T) {
  const bool pl = e->cfg.precision != CGVC_PREC_FP32_SIMT;
  A.n = n; A.T = T; A.xhi = A.xlo = nullptr; A.post = nullptr;
  long long r1 = (long long)n * T, r2 = r1 / 2, r4 = r1 / 4;
  if (pl) { A.xhi = ws.take<__nv_bfloat16>((size_t)r1 * 128); A.xlo = ws.take<__nv_bfloat16>((size_t)r1 * 128); }   // input planes, channels padded to 64 (128: F16F8)
  A.xchi = A.xclo = nullptr; A.z = nullptr;
  if (pl) {                                                    // tap-lowered edge layers (edge_on)
    const size_t cp = (size_t)edge_cpad(e->gen[0].h1.a.kw * e->cfg.num_features);
    A.xchi = ws.take<__nv_bfloat16>((size_t)r1 * cp); A.xclo = ws.take<__nv_bfloat16>((size_t)r1 * cp);
    A.z = ws.take<float>((size_t)r1 * e->gen[0].o1.kw * e->cfg.num_features);
  }
  plan_gated(ws, A.h1, r1, 256, n, 128, pl, r1 * 128);
  plan_gated(ws, A.d[0], r2, 512, n, 256, pl, r2 * 256);
  plan_gated(ws, A.d[1], r4, 1024, n, 512, pl, r4 * 512);
  for (int i = 0; i < 6; ++i) {
    plan_gated(ws, A.r[i].a, r4, 2048, n, 1024, pl, r4 * 1024);
    A.r[i].Pb = ws.take<float>((size_t)r4 * 512);
    A.r[i].sb = ws.take<float>((size_t)n * 4 * 512);
    A.r[i].Yr = ws.take<float>((size_t)r4 * 512);
    A.r[i].Yrhi = A.r[i].Yrlo = nullptr;
    if (pl) { A.r[i].Yrhi = ws.take<__nv_bfloat16>((size_t)r4 * 512); A.r[i].Yrlo = ws.take<__nv_bfloat16>((size_t)r4 * 512); }
  }
  plan_gated(ws, A.u[0], r4, 2048, n, 512, pl, r2 * 512);
  plan_gated(ws, A.u[1], r2, 1024, n, 256, pl, r1 * 256);
  A.out_cl = ws.take<float>((size_t)r1 * e->cfg.num_features);
}

// keep_y: also write the fp32 copy of every activation (debug taps / SIMT path); the tensor-core training path only
// needs fp32 where a residual add or the discriminator head reads it.
// save_pre = false (inference): the fused layers do not write their pre-norm outputs / statistics (nothing runs backward)
static int generator_forward(cgvc_engine* e, const GenNet& N, GenActs& A, const float* x_cl, cudaStream_t st, bool keep_y, bool save_pre = true) {
  const int n = A.n, T = A.T, nf = e->cfg.num_features;
  const float* Pm = e->P();
  A.x_cl = x_cl;
  const bool edge = edge_on(e, N) && A.xchi && A.z;
  const int qm = e->cfg.precision == CGVC_PREC_F16F8;
  ConvIO io; io.x = x_cl; io.xhi = A.xhi; io.xlo = A.xlo; io.n = n; io.H = 1; io.W = T;
  if (edge) {
    // h1 = dense [n*T, kw*F] x [kw*F, 2*128] GEMM on the im2col of the input
    CK(launch_im2col_taps(x_cl, (long long)n * T, T, nf, N.h1.a.kw, +1, edge_cpad(N.h1.a.kw * nf), qm, A.xchi, A.xclo, st));
    int r = tc_conv_fwd(e->tcw, N.h1c_slot, e->cfg.precision, A.xchi, A.xclo, n, 1, T, 1, 1, A.h1.P, st);
    if (r != 0) return fail(e, r == TC_UNSUPPORTED ? CGVC_ERR_UNSUPPORTED : CGVC_ERR_CUDA, "tc h1 fwd (tap-lowered): %d", r);
  } else {
    if (A.xhi && tc_enabled(e)) {
      if (qm) CK(launch_pad_split_q(x_cl, (long long)n * T, nf, nf, 128, A.xhi, A.xlo, st));
      else CK(launch_pad_split(x_cl, (long long)n * T, nf, nf, 64, A.xhi, A.xlo, st));
    }
    RET(gated_conv_fwd(e, N.h1, io, A.h1.P, st));
  }
  { PostParams q = post_params(e, N.h1, A.h1, n, T, keep_y, A.post); CK(launch_post_fwd(q, st)); }
  const GLAct* cur = &A.h1;
  int W = T;
  for (int i = 0; i < 2; ++i) {
    io.x = cur->Y; io.xhi = cur->Yhi; io.xlo = cur->Ylo; io.W = W;
    if (!keep_y && cur->Yhi) io.x = nullptr;
    W /= 2;
    RET(gated_layer_forward(e, N.d[i], io, A.d[i], n, W, keep_y || i == 1, A.post, st, save_pre));   // d2's fp32 output is the first residual input
    cur = &A.d[i];
  }
  const float* res = A.d[1].Y; const __nv_bfloat16 *rhi = A.d[1].Yhi, *rlo = A.d[1].Ylo;
  for (int i = 0; i < 6; ++i) {
    const ResBlock& R = N.r[i];
    io.x = res; io.xhi = rhi; io.xlo = rlo; io.W = W;
    RET(gated_layer_forward(e, R.h1, io, A.r[i].a, n, W, keep_y, A.post, st, save_pre));
    ConvIO io2; io2.x = (keep_y || !A.r[i].a.Yhi) ? A.r[i].a.Y : nullptr; io2.xhi = A.r[i].a.Yhi; io2.xlo = A.r[i].a.Ylo; io2.n = n; io2.H = 1; io2.W = W;
    bool done = false, fused = false;
    if (use_tc(e, R.tc_slot2) && io2.xhi) {
      TcFuse f; memset(&f, 0, sizeof f);
      f.R = e->fuse_in && A.r[i].Yrhi ? W : 0;                // R = 0: plain conv epilogue
      f.gamma_a = Pm + R.in2.gamma; f.beta_a = Pm + R.in2.beta; f.stats = save_pre ? A.r[i].sb : nullptr; f.resid = res;
      f.y = A.r[i].Yr; f.y_hi = A.r[i].Yrhi; f.y_lo = A.r[i].Yrlo;
      int r = tc_conv_fwd_fused(e->tcw, R.tc_slot2, e->cfg.precision, io2.xhi, io2.xlo, n, 1, W, 1, 1, (save_pre || !f.R) ? A.r[i].Pb : nullptr, f, &fused, st);
      if (r != 0 && !save_pre && f.R) { f.stats = A.r[i].sb; r = tc_conv_fwd_fused(e->tcw, R.tc_slot2, e->cfg.precision, io2.xhi, io2.xlo, n, 1, W, 1, 1, A.r[i].Pb, f, &fused, st); }
      if (r == 0) done = true; else if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc h2 fwd: %s", cudaGetErrorString((cudaError_t)r));
    }
    if (!done) RET(conv_fwd_simt(e, Pm, R.h2, 1, 1, io2, A.r[i].Pb, 512, 0, st));
    if (fused) { res = A.r[i].Yr; rhi = A.r[i].Yrhi; rlo = A.r[i].Yrlo; continue; }
    PostParams q; memset(&q, 0, sizeof q);
    q.p = A.r[i].Pb; q.ldp = 512; q.Cc = 512; q.B = n; q.R = W; q.C = 512; q.sh = 1;
    q.beta_a = Pm + R.in2.beta; q.gamma_a = Pm + R.in2.gamma; q.has_in = 1; q.has_gate = 0;
    q.resid = res; q.y = A.r[i].Yr; q.stats = A.r[i].sb; q.y_hi = A.r[i].Yrhi; q.y_lo = A.r[i].Yrlo; q.scratch = A.post;
    q.qmode = e->cfg.precision == CGVC_PREC_F16F8;
    CK(launch_post_fwd(q, st));
    res = A.r[i].Yr; rhi = A.r[i].Yrhi; rlo = A.r[i].Yrlo;
  }
  io.x = res; io.xhi = rhi; io.xlo = rlo;
  for (int i = 0; i < 2; ++i) {
    io.W = W;
    RET(gated_layer_forward(e, N.u[i], io, A.u[i], n, W, keep_y, A.post, st, save_pre));      // W = conv rows per sample; the shuffle doubles them
    W *= 2;
    io.x = (keep_y || !A.u[i].Yhi) ? A.u[i].Y : nullptr; io.xhi = A.u[i].Yhi; io.xlo = A.u[i].Ylo;
  }
  io.W = W;
  {
    bool done = false;
    if (edge && io.xhi) {
      // o1: Z[m, (t, c)] = U[m, :] . W[t][:, c] as one dense GEMM, then out[m, c] = b[c] + sum_t Z[m + t - 7, (t, c)]
      int r = tc_conv_fwd(e->tcw, N.o1f_slot, e->cfg.precision, io.xhi, io.xlo, n, 1, W, 1, 1, A.z, st);
      if (r != 0) return fail(e, r == TC_UNSUPPORTED ? CGVC_ERR_UNSUPPORTED : CGVC_ERR_CUDA, "tc o1 fwd (tap-lowered): %d", r);
      CK(launch_col2im_taps(A.z, N.o1.kw * nf, (long long)n * W, W, nf, N.o1.kw, +1, Pm + N.o1.b, A.out_cl, st));
      done = true;
    }
    if (!done && use_tc(e, N.o1_slot) && io.xhi) {
      int r = tc_conv_fwd(e->tcw, N.o1_slot, e->cfg.precision, io.xhi, io.xlo, n, 1, W, 1, 1, A.out_cl, st);
      if (r == 0) done = true; else if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc o1 fwd: %s", cudaGetErrorString((cudaError_t)r));
    }
    if (!done) RET(conv_fwd_simt(e, Pm, N.o1, 1, 1, io, A.out_cl, nf, 0, st));
  }
  if (keep_y) {
    e->taps.clear();
    size_t r1 = (size_t)n * T;
    e->taps["h1_glu"] = {A.h1.Y, r1 * 128}; e->taps["d1"] = {A.d[0].Y, r1 / 2 * 256}; e->taps["d2"] = {A.d[1].Y, r1 / 4 * 512};
    for (int i = 0; i < 6; ++i) e->taps["r" + std::to_string(i + 1)] = {A.r[i].Yr, r1 / 4 * 512};
    e->taps["u1"] = {A.u[0].Y, r1 / 2 * 512}; e->taps["u2"] = {A.u[1].Y, r1 * 256};
    e->taps["out_cl"] = {A.out_cl, r1 * (size_t)nf};
  }
  return 0;
}

struct BwdScratch { float *bufA, *bufB, *dP; __nv_bfloat16 *dPhi, *dPlo; float* post;
                    __nv_bfloat16 *dP2hi, *dP2lo;      // second plane pair: a fused dgrad epilogue writes the next layer's dP while reading this one's
                    __nv_bfloat16 *dPbhi, *dPblo;      // ping-pong partner of dPhi / dPlo (same size) for the side-stream weight gradients
                    SideQ* sq; };

struct PlanePair { __nv_bfloat16 *hi, *lo; };

// fp32 dP is only materialised when a SIMT kernel will read it
static PostBwdParams post_bwd_params(const cgvc_engine* e, const Gated& L, const float* dy, const GLAct& A,
                                     int n, int rows_per_sample_out, const BwdScratch& S, bool wgrad, bool need_fp32,
                                     const PlanePair* out = nullptr) {
  PostBwdParams q; memset(&q, 0, sizeof q);
  const float* Pm = e->P(); float* Gm = e->G();
  q.dy1 = dy; q.p = A.P; q.ldp = 2 * L.a.cout; q.Cc = L.a.cout; q.B = n; q.sh = L.shuffle;
  q.R = rows_per_sample_out * L.shuffle; q.C = L.a.cout / L.shuffle;
  q.has_in = L.has_in; q.has_gate = 1; q.stats = A.stats;
  if (L.has_in) {
    q.beta_a = Pm + L.ina.beta; q.gamma_a = Pm + L.ina.gamma; q.beta_g = Pm + L.ing.beta; q.gamma_g = Pm + L.ing.gamma;
    if (wgrad) { q.dbeta_a = Gm + L.ina.beta; q.dgamma_a = Gm + L.ina.gamma; q.dbeta_g = Gm + L.ing.beta; q.dgamma_g = Gm + L.ing.gamma; }
  }
  if (wgrad) { q.dbias_a = Gm + L.a.b; q.dbias_g = Gm + L.g.b; }
  q.scratch = S.post;
  const bool tc = use_tc(e, L.tc_slot) && S.dPhi;
  q.dp = (!tc || need_fp32) ? S.dP : nullptr;
  if (tc) { q.dp_hi = out ? out->hi : S.dPhi; q.dp_lo = out ? out->lo : S.dPlo; }
  q.qmode = e->cfg.precision == CGVC_PREC_F16F8;
  return q;
}

// the plane pair the next GLU / instance-norm backward may write (waits, on st, for the weight gradient that last read it)
static PlanePair dp_acquire(const BwdScratch& S, cudaStream_t st) {
  SideQ* q = S.sq;
  if (!q || !q->on || !S.dPbhi) return PlanePair{S.dPhi, S.dPlo};
  q->cur ^= 1;
  if (q->used[q->cur]) cudaStreamWaitEvent(st, q->done[q->cur], 0);
  return q->cur ? PlanePair{S.dPbhi, S.dPblo} : PlanePair{S.dPhi, S.dPlo};
}
// stream for the weight gradient of the planes acquired last (their producer has been enqueued on st) ...
static cudaStream_t wgrad_begin(const BwdScratch& S, cudaStream_t st) {
  SideQ* q = S.sq;
  if (!q || !q->on || !S.dPbhi) return st;
  cudaEventRecord(q->ready[q->cur], st);
  cudaStreamWaitEvent(q->side, q->ready[q->cur], 0);
  return q->side;
}
// ... and the end of that weight gradient
static void wgrad_end(const BwdScratch& S) {
  SideQ* q = S.sq;
  if (!q || !q->on || !S.dPbhi) return;
  cudaEventRecord(q->done[q->cur], q->side);
  q->used[q->cur] = true;
}
// every side-stream weight gradient of this lane has finished before st continues
static void side_join(const BwdScratch& S, cudaStream_t st) {
  SideQ* q = S.sq;
  if (!q) return;
  for (int b = 0; b < 2; ++b) if (q->used[b]) { cudaStreamWaitEvent(st, q->done[b], 0); q->used[b] = false; }
}

// fused-backward descriptors (see tc_conv_dgrad_fused): instance-norm backward of a residual block's h2 convolution ...
static TcBwdFuse in2_bwd_fuse(const cgvc_engine* e, const ResBlock& R, const float* Pb, const float* sb, int rows_per_sample, PlanePair out) {
  TcBwdFuse f; memset(&f, 0, sizeof f);
  const float* Pm = e->P(); float* Gm = e->G();
  f.R = rows_per_sample; f.gated = 0; f.bp = Pb; f.bp_ld = 512; f.stats = sb;
  f.gamma_a = Pm + R.in2.gamma; f.beta_a = Pm + R.in2.beta;
  f.dp_hi = out.hi; f.dp_lo = out.lo; f.dp_ld = 512;
  f.dgamma_a = Gm + R.in2.gamma; f.dbeta_a = Gm + R.in2.beta;
  return f;
}
// ... and GLU + instance-norm backward of a gated layer (no pixel shuffle)
static TcBwdFuse gated_bwd_fuse(const cgvc_engine* e, const Gated& L, const GLAct& A, int rows_per_sample, PlanePair out) {
  TcBwdFuse f; memset(&f, 0, sizeof f);
  const float* Pm = e->P(); float* Gm = e->G();
  f.R = (L.has_in && L.shuffle == 1) ? rows_per_sample : 0;      // R = 0: not fusable
  f.gated = 1; f.bp = A.P; f.bp_ld = 2 * L.a.cout; f.stats = A.stats;
  f.gamma_a = Pm + L.ina.gamma; f.beta_a = Pm + L.ina.beta; f.gamma_g = Pm + L.ing.gamma; f.beta_g = Pm + L.ing.beta;
  f.dp_hi = out.hi; f.dp_lo = out.lo; f.dp_ld = 2 * L.a.cout;
  f.dgamma_a = Gm + L.ina.gamma; f.dbeta_a = Gm + L.ina.beta; f.dgamma_g = Gm + L.ing.gamma; f.dbeta_g = Gm + L.ing.beta;
  return f;
}

// Backward through one generator application.  d_out_cl: [n*T, 24] gradient w.r.t. the channels-last output.
// Weight gradients are accumulated into the GRAD arena; d_in_cl (optional) receives d loss / d input (channels-last).
static int generator_backward(cgvc_engine* e, const GenNet& N, const GenActs& A, const float* d_out_cl, float* d_in_cl,
                              const BwdScratch& S, cudaStream_t st) {
  const int n = A.n, T = A.T, nf = e->cfg.num_features;
  const float* Pm = e->P(); float* Gm = e->G();
  ConvIO io; io.n = n; io.H = 1;
  // o1 (no norm, no gate): bias gradient = column sums of d_out
  io.x = A.u[1].Y; io.xhi = A.u[1].Yhi; io.xlo = A.u[1].Ylo; io.W = T;
  CK(launch_colsum(d_out_cl, (long long)n * T, nf, 0, nf, Gm + N.o1.b, st));
  {
    bool done = false;
    const bool edge = edge_on(e, N) && A.xchi && A.z;
    if (edge && io.xhi && S.dPhi) {
      // tap-lowered o1: dZ[m, (t, c)] = d_out[m - t + 7, c] (im2col of the 24-channel gradient), then dense weight and data gradients
      const PlanePair pp = dp_acquire(S, st);
      CK(launch_im2col_taps(d_out_cl, (long long)n * T, T, nf, N.o1.kw, -1, edge_cpad(N.o1.kw * nf), e->cfg.precision == CGVC_PREC_F16F8, pp.hi, pp.lo, st));
      cudaStream_t ws = wgrad_begin(S, st);
      int r = tc_conv_wgrad(e->tcw, N.o1f_slot, e->cfg.precision, io.xhi, io.xlo, pp.hi, pp.lo, n, 1, T, 1, 1, Gm + N.o1.k, nullptr, nullptr, nullptr, ws);
      wgrad_end(S);
      if (r == 0) r = tc_conv_dgrad(e->tcw, N.o1f_slot, e->cfg.precision, pp.hi, pp.lo, n, 1, T, 1, 1, S.bufA, 0, st);
      if (r != 0) return fail(e, r == TC_UNSUPPORTED ? CGVC_ERR_UNSUPPORTED : CGVC_ERR_CUDA, "tc o1 bwd (tap-lowered): %d", r);
      done = true;
    }
    if (!done && use_tc(e, N.o1_slot) && io.xhi && S.dPhi) {
      const PlanePair pp = dp_acquire(S, st);
      if (e->cfg.precision == CGVC_PREC_F16F8) CK(launch_pad_split_q(d_out_cl, (long long)n * T, nf, nf, 128, pp.hi, pp.lo, st));
      else CK(launch_pad_split(d_out_cl, (long long)n * T, nf, nf, 64, pp.hi, pp.lo, st));
      cudaStream_t ws = wgrad_begin(S, st);
      int r = tc_conv_wgrad(e->tcw, N.o1_slot, e->cfg.precision, io.xhi, io.xlo, pp.hi, pp.lo, n, 1, T, 1, 1, Gm + N.o1.k, nullptr, nullptr, nullptr, ws);
      wgrad_end(S);
      if (r == 0) r = tc_conv_dgrad(e->tcw, N.o1_slot, e->cfg.precision, pp.hi, pp.lo, n, 1, T, 1, 1, S.bufA, 0, st);
      if (r == 0) done = true; else if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc o1 bwd: %s", cudaGetErrorString((cudaError_t)r));
    }
    if (!done) {
      RET(conv_wgrad_simt(e, Gm, N.o1, 1, 1, io, d_out_cl, nf, 0, st));
      RET(conv_dgrad_simt(e, Pm, N.o1, 1, 1, n, 1, T, d_out_cl, nf, 0, S.bufA, 0, st));
    }
  }
  float* cur = S.bufA; float* oth = S.bufB;
  int W = T;      // W tracks the output width of the layer being differentiated
  // Fused backward (tensor-core path): a stride-1 data-gradient launch whose result is d loss / d (output of an instance-normed
  // layer) runs that layer's instance-norm (+GLU) backward in its epilogue and writes the layer's dP planes directly.  It reads
  // one plane pair while writing the other: pb[0] / pb[1]; `have` = which pair holds the dP planes of the layer differentiated next.
  const bool side_on = S.sq && S.sq->on && S.dPbhi;
  const bool fuse_ok = e->fuse_bwd && !side_on && tc_enabled(e) && S.dPhi && S.dP2hi && A.r[0].a.Yhi && A.r[0].Yrhi;
  PlanePair pb[2] = {{S.dPhi, S.dPlo}, {S.dP2hi, S.dP2lo}};
  int have = -1;
  for (int i = 1; i >= 0; --i) {
    // upsample block i: conv at width W/2 -> shuffle -> width W
    int Wc = W / 2;
    const float* Xin; const __nv_bfloat16 *Xhi, *Xlo;
    if (i == 1) { Xin = A.u[0].Y; Xhi = A.u[0].Yhi; Xlo = A.u[0].Ylo; } else { Xin = A.r[5].Yr; Xhi = A.r[5].Yrhi; Xlo = A.r[5].Yrlo; }
    const PlanePair ppu = dp_acquire(S, st);
    PostBwdParams q = post_bwd_params(e, N.u[i], cur, A.u[i], n, Wc, S, true, false, &ppu);
    CK(launch_post_bwd(q, st));
    io.x = Xin; io.xhi = Xhi; io.xlo = Xlo; io.W = Wc;
    { cudaStream_t ws = q.dp_hi ? wgrad_begin(S, st) : st;
      int rw = gated_conv_wgrad(e, N.u[i], io, q.dp, q.dp_hi, q.dp_lo, ws);
      if (q.dp_hi) wgrad_end(S);
      RET(rw); }
    bool fusedu = false;
    if (i == 0 && fuse_ok) {
      // u1's data gradient is d loss / d (residual block 6 output): run that block's h2 instance-norm backward in the epilogue
      TcBwdFuse f = in2_bwd_fuse(e, N.r[5], A.r[5].Pb, A.r[5].sb, Wc, pb[1]);
      int r = tc_conv_dgrad_fused(e->tcw, N.u[0].tc_slot, e->cfg.precision, q.dp_hi, q.dp_lo, n, 1, Wc, 1, 1, oth, 0, f, &fusedu, st);
      if (r != 0 && r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc u1 dgrad (fused): %s", cudaGetErrorString((cudaError_t)r));
      if (r != 0) fusedu = false;
      if (r == 0 && fusedu) have = 1;
      if (r != 0) RET(gated_conv_dgrad(e, N.u[i], n, 1, Wc, q.dp, q.dp_hi, q.dp_lo, oth, 0, st));
    } else {
      RET(gated_conv_dgrad(e, N.u[i], n, 1, Wc, q.dp, q.dp_hi, q.dp_lo, oth, 0, st));
    }
    float* t = cur; cur = oth; oth = t;
    W = Wc;
  }
  // residual blocks (width W = T/4); cur holds d(block output)
  for (int i = 5; i >= 0; --i) {
    const ResBlock& R = N.r[i];
    const float* Xin; const __nv_bfloat16 *Xhi, *Xlo;
    if (i == 0) { Xin = A.d[1].Y; Xhi = A.d[1].Yhi; Xlo = A.d[1].Ylo; } else { Xin = A.r[i - 1].Yr; Xhi = A.r[i - 1].Yrhi; Xlo = A.r[i - 1].Yrlo; }
    const bool tc2 = use_tc(e, R.tc_slot2) && S.dPhi && A.r[i].a.Yhi;
    // (1) dP of the h2 convolution: already in pb[have] when the previous data-gradient launch fused it, else the streaming kernels
    int b2 = have;
    if (b2 < 0) {
      if (side_on) pb[0] = dp_acquire(S, st);
      PostBwdParams q; memset(&q, 0, sizeof q);
      q.dy1 = cur; q.p = A.r[i].Pb; q.ldp = 512; q.Cc = 512; q.B = n; q.R = W; q.C = 512; q.sh = 1;
      q.beta_a = Pm + R.in2.beta; q.gamma_a = Pm + R.in2.gamma; q.has_in = 1; q.has_gate = 0; q.stats = A.r[i].sb;
      q.dp = tc2 ? nullptr : S.dP; if (tc2) { q.dp_hi = pb[0].hi; q.dp_lo = pb[0].lo; }
      q.qmode = e->cfg.precision == CGVC_PREC_F16F8;
      q.scratch = S.post;
      q.dbeta_a = Gm + R.in2.beta; q.dgamma_a = Gm + R.in2.gamma; q.dbias_a = Gm + R.h2.b;
      CK(launch_post_bwd(q, st));
      b2 = 0;
    }
    have = -1;
    ConvIO io2; io2.x = A.r[i].a.Y; io2.xhi = A.r[i].a.Yhi; io2.xlo = A.r[i].a.Ylo; io2.n = n; io2.H = 1; io2.W = W;
    // (2) h2: weight gradient, then data gradient = d loss / d (h1's GLU output), with h1's GLU + instance-norm backward fused
    bool done = false, fused1 = false;
    int b1 = 0;
    if (tc2) {
      cudaStream_t ws = side_on ? wgrad_begin(S, st) : st;
      int r = tc_conv_wgrad(e->tcw, R.tc_slot2, e->cfg.precision, io2.xhi, io2.xlo, pb[b2].hi, pb[b2].lo, n, 1, W, 1, 1,
                            Gm + R.h2.k, nullptr, nullptr, nullptr, ws);
      if (side_on) wgrad_end(S);
      if (r == 0) {
        if (fuse_ok && use_tc(e, R.h1.tc_slot)) {
          TcBwdFuse f = gated_bwd_fuse(e, R.h1, A.r[i].a, W, pb[1 - b2]);
          r = tc_conv_dgrad_fused(e->tcw, R.tc_slot2, e->cfg.precision, pb[b2].hi, pb[b2].lo, n, 1, W, 1, 1, oth, 0, f, &fused1, st);
          if (r == 0 && fused1) b1 = 1 - b2;
        } else {
          r = tc_conv_dgrad(e->tcw, R.tc_slot2, e->cfg.precision, pb[b2].hi, pb[b2].lo, n, 1, W, 1, 1, oth, 0, st);
        }
      }
      if (r == 0) done = true; else if (r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc h2 bwd: %s", cudaGetErrorString((cudaError_t)r));
    }
    if (!done) {
      RET(conv_wgrad_simt(e, Gm, R.h2, 1, 1, io2, S.dP, 512, 0, st));
      RET(conv_dgrad_simt(e, Pm, R.h2, 1, 1, n, 1, W, S.dP, 512, 0, oth, 0, st));
    }
    const float* dp1 = nullptr; const __nv_bfloat16 *dp1hi = pb[b1].hi, *dp1lo = pb[b1].lo;
    if (!fused1) {
      // writes pb[0] (free again: h2's launches are done), or the other buffer of the ping-pong when the weight gradients run aside
      const PlanePair pp1 = side_on ? dp_acquire(S, st) : pb[0];
      PostBwdParams q2 = post_bwd_params(e, R.h1, oth, A.r[i].a, n, W, S, true, false, &pp1);
      CK(launch_post_bwd(q2, st));
      dp1 = q2.dp; dp1hi = q2.dp_hi; dp1lo = q2.dp_lo; b1 = 0;
    }
    // (3) h1: weight gradient, then d_in = d_out (skip) + data gradient, in place in `cur`; that is d loss / d (previous block's
    //     output) -- or, for the first block, of the second down-sampling layer's output: fuse that layer's backward as well
    io.x = Xin; io.xhi = Xhi; io.xlo = Xlo; io.W = W;
    { cudaStream_t ws = (side_on && dp1hi) ? wgrad_begin(S, st) : st;
      int rw = gated_conv_wgrad(e, R.h1, io, dp1, dp1hi, dp1lo, ws);
      if (side_on && dp1hi) wgrad_end(S);
      RET(rw); }
    bool fused0 = false;
    if (fuse_ok && use_tc(e, R.h1.tc_slot) && dp1hi) {
      TcBwdFuse f = (i > 0) ? in2_bwd_fuse(e, N.r[i - 1], A.r[i - 1].Pb, A.r[i - 1].sb, W, pb[1 - b1])
                            : gated_bwd_fuse(e, N.d[1], A.d[1], W, pb[1 - b1]);
      int r = tc_conv_dgrad_fused(e->tcw, R.h1.tc_slot, e->cfg.precision, dp1hi, dp1lo, n, 1, W, 1, 1, cur, 1, f, &fused0, st);
      if (r != 0 && r != TC_UNSUPPORTED) return fail(e, CGVC_ERR_CUDA, "tc h1 dgrad (fused): %s", cudaGetErrorString((cudaError_t)r));
      if (r != 0) { fused0 = false; RET(gated_conv_dgrad(e, R.h1, n, 1, W, dp1, dp1hi, dp1lo, cur, 1, st)); }
      if (fused0) have = 1 - b1;
    } else {
      RET(gated_conv_dgrad(e, R.h1, n, 1, W, dp1, dp1hi, dp1lo, cur, 1, st));
    }
  }
  // downsample blocks
  for (int i = 1; i >= 0; --i) {
    const GLAct& in = (i == 1) ? A.d[0] : A.h1;
    const PlanePair ppd = (i == 1 && have >= 0) ? pb[have] : dp_acquire(S, st);
    PostBwdParams q = post_bwd_params(e, N.d[i], cur, A.d[i], n, W, S, true, false, &ppd);
    if (i == 1 && have >= 0) { q.dp = nullptr; q.dp_hi = pb[have].hi; q.dp_lo = pb[have].lo; have = -1; }   // produced by the fused epilogue of r1.h1's dgrad
    else CK(launch_post_bwd(q, st));
    io.x = in.Y; io.xhi = in.Yhi; io.xlo = in.Ylo; io.W = W * 2;
    { cudaStream_t ws = (side_on && q.dp_hi) ? wgrad_begin(S, st) : st;
      int rw = gated_conv_wgrad(e, N.d[i], io, q.dp, q.dp_hi, q.dp_lo, ws);
      if (side_on && q.dp_hi) wgrad_end(S);
      RET(rw); }
    RET(gated_conv_dgrad(e, N.d[i], n, 1, W * 2, q.dp, q.dp_hi, q.dp_lo, oth, 0, st));
    float* t = cur; cur = oth; oth = t;
    W *= 2;
  }
  // h1 (no IN)
  {
    const PlanePair pph = dp_acquire(S, st);
    PostBwdParams q = post_bwd_params(e, N.h1, cur, A.h1, n, T, S, true, false, &pph);
    CK(launch_post_bwd(q, st));
    io.x = A.x_cl; io.xhi = A.xhi; io.xlo = A.xlo; io.W = T;
    if (edge_on(e, N) && A.xchi && A.z && q.dp_hi) {
      // tap-lowered h1: the weight gradient is im2col(x)^T dP straight into the [15,24,128] kernels' GRAD ranges; the data gradient
      // (cycle passes only) is the dense dP . W^T [n*T, kw*F] followed by the tap-shifted sum
      { cudaStream_t ws = side_on ? wgrad_begin(S, st) : st;
        int rw = tc_conv_wgrad(e->tcw, N.h1c_slot, e->cfg.precision, A.xchi, A.xclo, q.dp_hi, q.dp_lo, n, 1, T, 1, 1,
                               Gm + N.h1.a.k, Gm + N.h1.g.k, nullptr, nullptr, ws);
        if (side_on) wgrad_end(S);
        if (rw != 0) return fail(e, rw == TC_UNSUPPORTED ? CGVC_ERR_UNSUPPORTED : CGVC_ERR_CUDA, "tc h1 wgrad (tap-lowered): %d", rw); }
      if (d_in_cl) {
        int r = tc_conv_dgrad(e->tcw, N.h1c_slot, e->cfg.precision, q.dp_hi, q.dp_lo, n, 1, T, 1, 1, oth, 0, st);
        if (r != 0) return fail(e, r == TC_UNSUPPORTED ? CGVC_ERR_UNSUPPORTED : CGVC_ERR_CUDA, "tc h1 dgrad (tap-lowered): %d", r);
        CK(launch_col2im_taps(oth, N.h1.a.kw * nf, (long long)n * T, T, nf, N.h1.a.kw, -1, nullptr, d_in_cl, st));
      }
      return 0;
    }
    { cudaStream_t ws = (side_on && q.dp_hi) ? wgrad_begin(S, st) : st;
      int rw = gated_conv_wgrad(e, N.h1, io, q.dp, q.dp_hi, q.dp_lo, ws);
      if (side_on && q.dp_hi) wgrad_end(S);
      RET(rw); }
    if (d_in_cl) RET(gated_conv_dgrad(e, N.h1, n, 1, T, q.dp, q.dp_hi, q.dp_lo, d_in_cl, 0, st));
  }
  return 0;
}

// ---- discriminator ---------------------------------------------------------------------------------------
static void plan_discriminator(cgvc_engine* e, Bump& ws, DiscActs& A, int n, int T) {
  const bool pl = e->cfg.precision != CGVC_PREC_FP32_SIMT;
  A.n = n; A.T = T; A.post = nullptr;
  const int H = e->cfg.num_features;
  long long r0 = (long long)n * H * (T / 2), r1 = (long long)n * (H / 2) * (T / 4), r2 = (long long)n * (H / 4) * (T / 8), r3 = (long long)n * (H / 4) * (T / 16);
  plan_gated(ws, A.h1, r0, 256, n, 128, pl, r0 * 128);
  plan_gated(ws, A.d[0], r1, 512, n, 256, pl, r1 * 256);
  plan_gated(ws, A.d[1], r2, 1024, n, 512, pl, r2 * 512);
  plan_gated(ws, A.d[2], r3, 2048, n, 1024, false, r3 * 1024);
  A.prob = ws.take<float>((size_t)r3);
}

static int discriminator_forward(cgvc_engine* e, const DiscNet& N, DiscActs& A, const float* x, cudaStream_t st, bool keep_y) {
  const int n = A.n, T = A.T, H0 = e->cfg.num_features;
  const float* Pm = e->P();
  A.x = x;
  ConvIO io; io.x = x; io.xhi = nullptr; io.xlo = nullptr; io.n = n; io.H = H0; io.W = T;
  int H = H0, W = T / 2;
  if (e->fuse_c1 && !use_tc(e, N.h1.tc_slot) && N.h1.a.cin == 1 && !N.h1.has_in && N.h1.a.kh * N.h1.a.kw <= 9 && N.h1.a.cout % 4 == 0 &&
      256 % (N.h1.a.cout / 4) == 0) {
    // input layer (one input channel, K = 9, gate without norm): convolution + GLU in one HBM-bound pass; P is kept for the backward pass
    const GatherGeom g = fwd_geom(n, H0, T, N.h1.a.kh, N.h1.a.kw, N.h1.sh, N.h1.sw);
    const PostParams q = post_params(e, N.h1, A.h1, n, H * W, keep_y, A.post);
    CK(launch_conv_c1_glu_fwd(g, x, Pm + N.h1.a.k, Pm + N.h1.g.k, Pm + N.h1.a.b, Pm + N.h1.g.b, N.h1.a.cout, A.h1.P, q.y, q.y_hi, q.y_lo, q.qmode, st));
  } else {
    RET(gated_conv_fwd(e, N.h1, io, A.h1.P, st));
    PostParams q = post_params(e, N.h1, A.h1, n, H * W, keep_y, A.post); CK(launch_post_fwd(q, st));
  }
  const GLAct* cur = &A.h1;
  for (int i = 0; i < 3; ++i) {
    io.x = (keep_y || !cur->Yhi) ? cur->Y : nullptr; io.xhi = cur->Yhi; io.xlo = cur->Ylo; io.H = H; io.W = W;
    RET(gated_conv_fwd(e, N.d[i], io, A.d[i].P, st));
    int Ho, Wo; conv_out_dims(N.d[i].a, N.d[i].sh, N.d[i].sw, H, W, Ho, Wo); H = Ho; W = Wo;
    PostParams q = post_params(e, N.d[i], A.d[i], n, H * W, keep_y, A.post);     // d3 has no planes: its fp32 output feeds the head
    CK(launch_post_fwd(q, st));
    cur = &A.d[i];
  }
  CK(launch_head_fwd(cur->Y, (long long)n * H * W, 1024, Pm + N.dense_k, Pm + N.dense_b, A.prob, st));
  if (keep_y) {
    e->taps.clear();
    e->taps["h1_glu"] = {A.h1.Y, (size_t)n * H0 * (T / 2) * 128};
    e->taps["d1"] = {A.d[0].Y, (size_t)n * (H0 / 2) * (T / 4) * 256};
    e->taps["d2"] = {A.d[1].Y, (size_t)n * (H0 / 4) * (T / 8) * 512};
    e->taps["d3"] = {A.d[2].Y, (size_t)n * (H0 / 4) * (T / 16) * 1024};
  }
  return 0;
}

// view of samples [s0, s0+ns) of a DiscActs
static DiscActs disc_view(const cgvc_engine* e, const DiscActs& A, int s0, int ns) {
  DiscActs V = A; V.n = ns;
  const int H = e->cfg.num_features, T = A.T;
  long long rows[4] = {(long long)H * (T / 2), (long long)(H / 2) * (T / 4), (long long)(H / 4) * (T / 8), (long long)(H / 4) * (T / 16)};
  const int co[4] = {128, 256, 512, 1024};
  GLAct* src[4] = {const_cast<GLAct*>(&A.h1), const_cast<GLAct*>(&A.d[0]), const_cast<GLAct*>(&A.d[1]), const_cast<GLAct*>(&A.d[2])};
  GLAct* dst[4] = {&V.h1, &V.d[0], &V.d[1], &V.d[2]};
  for (int i = 0; i < 4; ++i) {
    dst[i]->P = src[i]->P + (long long)s0 * rows[i] * 2 * co[i];
    dst[i]->stats = src[i]->stats + (long long)s0 * 4 * co[i];
    dst[i]->Y = src[i]->Y + (long long)s0 * rows[i] * co[i];
    dst[i]->Yhi = src[i]->Yhi ? src[i]->Yhi + (long long)s0 * rows[i] * co[i] : nullptr;
    dst[i]->Ylo = src[i]->Ylo ? src[i]->Ylo + (long long)s0 * rows[i] * co[i] : nullptr;
  }
  V.x = A.x + (long long)s0 * H * T;
  V.prob = A.prob + (long long)s0 * rows[3];
  return V;
}

// dY3: gradient w.r.t. the d3 GLU output [n*48, 1024].  wgrad: accumulate weight gradients.  d_in: optional [n,24,T].
static int discriminator_backward(cgvc_engine* e, const DiscNet& N, const DiscActs& A, const float* dY3, bool wgrad, float* d_in,
                                  const BwdScratch& S, cudaStream_t st) {
  const int n = A.n, T = A.T, H0 = e->cfg.num_features;
  int Hs[4] = {H0, H0 / 2, H0 / 4, H0 / 4}, Ws[4] = {T / 2, T / 4, T / 8, T / 16};   // output dims of h1, d1, d2, d3
  const float* dy = dY3;
  float* bufs[2] = {S.bufA, S.bufB};
  int flip = 0;
  ConvIO io; io.n = n;
  for (int i = 2; i >= 0; --i) {
    const GLAct& in = (i == 0) ? A.h1 : A.d[i - 1];
    const PlanePair ppd = dp_acquire(S, st);
    PostBwdParams q = post_bwd_params(e, N.d[i], dy, A.d[i], n, Hs[i + 1] * Ws[i + 1], S, wgrad, false, &ppd);
    CK(launch_post_bwd(q, st));
    io.x = in.Y; io.xhi = in.Yhi; io.xlo = in.Ylo; io.H = Hs[i]; io.W = Ws[i];
    if (wgrad) {
      const bool aside = S.sq && S.sq->on && S.dPbhi && q.dp_hi;
      cudaStream_t ws = aside ? wgrad_begin(S, st) : st;
      int rw = gated_conv_wgrad(e, N.d[i], io, q.dp, q.dp_hi, q.dp_lo, ws);
      if (aside) wgrad_end(S);
      RET(rw);
    }
    RET(gated_conv_dgrad(e, N.d[i], n, Hs[i], Ws[i], q.dp, q.dp_hi, q.dp_lo, bufs[flip], 0, st));
    dy = bufs[flip]; flip ^= 1;
  }
  // h1: one input channel (K = 9), gate without instance norm.  Fused form: the GLU backward is recomputed inside the weight-gradient /
  // data-gradient kernels, dP never goes to HBM
  if (e->fuse_c1 && N.h1.a.cout == 128 && N.h1.a.kh * N.h1.a.kw <= 9 && !N.h1.has_in) {
    if (wgrad) {
      GatherGeom g = fwd_geom(n, H0, T, 3, 3, N.h1.sh, N.h1.sw);
      CK(launch_glu_bwd_wgrad_c1(g, A.x, dy, A.h1.P, 128, e->G() + N.h1.a.k, e->G() + N.h1.g.k, e->G() + N.h1.a.b, e->G() + N.h1.g.b, st));
    }
    if (d_in) CK(launch_glu_bwd_dgrad_c1(dy, A.h1.P, 128, e->P() + N.h1.a.k, e->P() + N.h1.g.k, bufs[flip], d_in, n, H0, T, 3, 3, N.h1.sh, N.h1.sw, st));
    return 0;
  }
  PostBwdParams q = post_bwd_params(e, N.h1, dy, A.h1, n, Hs[0] * Ws[0], S, wgrad, true);
  CK(launch_post_bwd(q, st));
  if (wgrad) {
    GatherGeom g = fwd_geom(n, H0, T, 3, 3, N.h1.sh, N.h1.sw);
    CK(launch_wgrad_c1(g, A.x, S.dP, 256, 256, e->G() + N.h1.a.k, e->G() + N.h1.g.k, 128, nullptr, nullptr, st));
  }
  if (d_in) CK(launch_dgrad_c1(S.dP, 256, e->P() + N.h1.a.k, e->P() + N.h1.g.k, 128, bufs[flip], d_in, n, H0, T, 3, 3, N.h1.sh, N.h1.sw, st));
  return 0;
}

// ---- workspace sizing ---------------------------------------------------------------------------------------
// The step is two symmetric, data-independent lanes that only meet in the (atomically accumulated) gradient arena and
// loss slots:   lane 0:  G_A2B([A;B]) -> [gen_B; id_B],  G_B2A(gen_B) -> cycle_A,  D_B([B; gen_B])
//               lane 1:  G_B2A([B;A]) -> [gen_A; id_A],  G_A2B(gen_A) -> cycle_B,  D_A([A; gen_A])
// They run on two streams so that one lane's HBM-bound kernels overlap the other lane's tensor-bound kernels.
struct LanePlan {
  GenActs gfirst, gcyc;         // batch 2B and B
  DiscActs d;                   // batch 2B
  float* in;                    // channels-last generator input [2B,T,24] = [X; Y]
  float* din;                   // discriminator input [2B,24,T] = [Y_real; gen_Y]
  float* d_cyc;                 // d cycle loss / d cycle_X, channels-last [B,T,24] (later reused as transpose scratch)
  float* d_out;                 // upstream gradient of the first pass [2B,T,24] = [d gen_Y; d id_Y]
  float* d_adv;                 // d G-adv / d gen_Y, [B,24,T]
  float* dY3;                   // [2B*48, 1024]
  BwdScratch S;
};
struct TrainPlan { LanePlan lane[2]; };

static void plan_train(cgvc_engine* e, Bump& ws, TrainPlan& P, int B, int T) {
  const int nf = e->cfg.num_features;
  const bool pl = e->cfg.precision != CGVC_PREC_FP32_SIMT;
  size_t img = (size_t)B * nf * T;
  size_t n2 = 2 * (size_t)B;
  size_t buf = n2 * (size_t)nf * (T / 2) * 128;            // largest dY: discriminator h1 output
  size_t bufg = n2 * (size_t)T * 256; if (bufg > buf) buf = bufg;
  size_t dp = n2 * (size_t)nf * (T / 2) * 256;             // largest dP: discriminator h1 conv output
  size_t dpg = n2 * (size_t)T * 512; if (dpg > dp) dp = dpg;
  for (int l = 0; l < 2; ++l) {
    LanePlan& L = P.lane[l];
    L.in = ws.take<float>(2 * img); L.din = ws.take<float>(2 * img);
    L.d_cyc = ws.take<float>(img); L.d_out = ws.take<float>(2 * img); L.d_adv = ws.take<float>(img);
    L.dY3 = ws.take<float>((size_t)2 * B * (nf / 4) * (T / 16) * 1024);
    L.S.bufA = ws.take<float>(buf); L.S.bufB = ws.take<float>(buf); L.S.dP = ws.take<float>(dp);
    L.S.dPhi = L.S.dPlo = L.S.dP2hi = L.S.dP2lo = nullptr;
    if (pl) {
      L.S.dPhi = ws.take<__nv_bfloat16>(dp); L.S.dPlo = ws.take<__nv_bfloat16>(dp);
      L.S.dP2hi = ws.take<__nv_bfloat16>(dpg); L.S.dP2lo = ws.take<__nv_bfloat16>(dpg);     // generator layers only
    }
    L.S.dPbhi = L.S.dPblo = nullptr;
    if (pl) { L.S.dPbhi = ws.take<__nv_bfloat16>(dp); L.S.dPblo = ws.take<__nv_bfloat16>(dp); }
    L.S.sq = &e->sideq[l];
    L.S.post = ws.take<float>(n2 * 4 * 1024);
    plan_generator(e, ws, L.gfirst, 2 * B, T); plan_generator(e, ws, L.gcyc, B, T);
    plan_discriminator(e, ws, L.d, 2 * B, T);
    L.gfirst.post = L.gcyc.post = L.d.post = L.S.post;
  }
}

struct FwdPlan { GenActs g; DiscActs d; float* in_cl; };

static size_t work_bytes_needed(cgvc_engine* e) {
  Bump ws; ws.reset(nullptr, 0);
  size_t need = 0;
  if (e->cfg.train) { TrainPlan P; plan_train(e, ws, P, e->cfg.max_batch, e->cfg.max_frames); need = ws.off; }
  ws.reset(nullptr, 0);
  FwdPlan F; F.in_cl = ws.take<float>((size_t)e->cfg.max_batch * e->cfg.num_features * e->cfg.max_frames);
  plan_generator(e, ws, F.g, e->cfg.max_batch, e->cfg.max_frames);
  plan_discriminator(e, ws, F.d, e->cfg.max_batch, e->cfg.max_frames);
  if (ws.off > need) need = ws.off;
  return need + 4096;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int cgvc_abi_version(void) { return CGVC_ABI_VERSION; }

const char* cgvc_last_error(cgvc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cgvc_create(const cgvc_config* cfg, cgvc_handle* out) {
  if (!cfg || !out) return fail(nullptr, CGVC_ERR_ARG, "cgvc_create: null argument");
  if (cfg->num_features != 24) return fail(nullptr, CGVC_ERR_ARG, "only num_features = 24 is supported (got %d)", cfg->num_features);
  if (cfg->max_batch < 1 || cfg->max_frames < 16 || cfg->max_frames % 4 != 0)
    return fail(nullptr, CGVC_ERR_ARG, "max_batch must be >= 1 and max_frames a multiple of 4, >= 16");
  if (cfg->precision < 0 || cfg->precision > 3) return fail(nullptr, CGVC_ERR_ARG, "unknown precision %d", cfg->precision);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, CGVC_ERR_CUDA, "no CUDA device available (%s): libcgvc has no CPU fallback", cudaGetErrorString(ce));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CGVC_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
  DeviceGuard dguard;
  ce = dguard.set(cfg->device);
  if (ce != cudaSuccess) return fail(nullptr, CGVC_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce));
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10) return fail(nullptr, CGVC_ERR_CUDA, "libcgvc is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  cgvc_engine* e = new cgvc_engine();
  e->cfg = *cfg;
  TableBuilder tb{e->tensors};
  const char* gn[2] = {"generator_A2B", "generator_B2A"}; const char* dn[2] = {"discriminator_A", "discriminator_B"};
  for (int i = 0; i < 2; ++i) { tb.scope = gn[i]; build_generator(tb, e->gen[i], cfg->num_features); }
  for (int i = 0; i < 2; ++i) { tb.scope = dn[i]; build_discriminator(tb, e->disc[i]); }
  e->n_params = (tb.off + 3) & ~(size_t)3;
  e->n_real_params = tb.real;
  for (int i = 0; i < 2; ++i) {
    GenNet& g = e->gen[i];
    g.h1.tc_slot = -1; g.o1_slot = -1; g.h1c_slot = -1; g.o1f_slot = -1; for (int k = 0; k < 2; ++k) { g.d[k].tc_slot = -1; g.u[k].tc_slot = -1; }
    for (int k = 0; k < 6; ++k) { g.r[k].h1.tc_slot = -1; g.r[k].tc_slot2 = -1; }
    DiscNet& d = e->disc[i]; d.h1.tc_slot = -1; for (int k = 0; k < 3; ++k) d.d[k].tc_slot = -1;
  }
  ce = post_init_kernels();
  if (ce != cudaSuccess) { delete e; return fail(nullptr, CGVC_ERR_CUDA, "post_init_kernels: %s", cudaGetErrorString(ce)); }
  ce = cudaMalloc(&e->d_scalars, 64 * sizeof(float));
  if (ce != cudaSuccess) { delete e; return fail(nullptr, CGVC_ERR_CUDA, "cudaMalloc scalars: %s", cudaGetErrorString(ce)); }
  cudaMemset(e->d_scalars, 0, 64 * sizeof(float));
  for (int l = 0; l < 2; ++l) { cudaStreamCreateWithFlags(&e->lane_stream[l], cudaStreamNonBlocking); cudaEventCreateWithFlags(&e->ev_join[l], cudaEventDisableTiming); }
  cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
  { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); cudaStreamCreateWithPriority(&e->comm_stream, cudaStreamNonBlocking, hi); }
  cudaEventCreateWithFlags(&e->ev_grads, cudaEventDisableTiming);
  for (int k = 0; k < 4; ++k) cudaEventCreateWithFlags(&e->ev_ar[k], cudaEventDisableTiming);
  for (int l = 0; l < 2; ++l) {
    cudaStreamCreateWithFlags(&e->sideq[l].side, cudaStreamNonBlocking);
    for (int b = 0; b < 2; ++b) { cudaEventCreateWithFlags(&e->sideq[l].ready[b], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->sideq[l].done[b], cudaEventDisableTiming); }
  }
  cudaStreamCreateWithFlags(&e->graph_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&e->ev_bridge, cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_bridge2, cudaEventDisableTiming);
  if (cfg->precision != CGVC_PREC_FP32_SIMT) {
    // register every dense gated layer with the tensor-core weight store
    for (int i = 0; i < 2; ++i) {
      GenNet& g = e->gen[i];
      // the 24-channel edge layers run on the tensor cores too (channel dims zero-padded to 64 / 128 inside the planes)
      g.h1.tc_slot = tc_register(e->tcw, g.h1.a.k, g.h1.g.k, g.h1.a.b, g.h1.g.b, 1, 15, g.h1.a.cin, g.h1.a.cout, 1);
      g.o1_slot = tc_register(e->tcw, g.o1.k, 0, g.o1.b, 0, 1, 15, g.o1.cin, g.o1.cout, 0);
      // ... and, preferred (edge_lower), as dense 1 x 1 layers with the taps in the channel (h1) / column (o1) dimension: see edge_on
      if ((g.h1.a.kw * g.h1.a.cin) % 4 == 0 && g.o1.cout % 4 == 0) {
        g.h1c_slot = tc_register(e->tcw, g.h1.a.k, g.h1.g.k, g.h1.a.b, g.h1.g.b, 1, 1, g.h1.a.kw * g.h1.a.cin, g.h1.a.cout, 1);
        g.o1f_slot = tc_register(e->tcw, g.o1.k, 0, g.o1.b, 0, 1, 1, g.o1.cin, g.o1.kw * g.o1.cout, 0, 1, g.o1.kw);
      }
      for (int k = 0; k < 2; ++k) g.d[k].tc_slot = tc_register(e->tcw, g.d[k].a.k, g.d[k].g.k, g.d[k].a.b, g.d[k].g.b, 1, 5, g.d[k].a.cin, g.d[k].a.cout, 1);
      for (int k = 0; k < 6; ++k) {
        g.r[k].h1.tc_slot = tc_register(e->tcw, g.r[k].h1.a.k, g.r[k].h1.g.k, g.r[k].h1.a.b, g.r[k].h1.g.b, 1, 3, 512, 1024, 1);
        g.r[k].tc_slot2 = tc_register(e->tcw, g.r[k].h2.k, 0, g.r[k].h2.b, 0, 1, 3, 1024, 512, 0);
      }
      for (int k = 0; k < 2; ++k) g.u[k].tc_slot = tc_register(e->tcw, g.u[k].a.k, g.u[k].g.k, g.u[k].a.b, g.u[k].g.b, 1, 5, g.u[k].a.cin, g.u[k].a.cout, 1, 2);
      DiscNet& d = e->disc[i];
      for (int k = 0; k < 3; ++k) d.d[k].tc_slot = tc_register(e->tcw, d.d[k].a.k, d.d[k].g.k, d.d[k].a.b, d.d[k].g.b, d.d[k].a.kh, 3, d.d[k].a.cin, d.d[k].a.cout, 1);
    }
    e->tcw.quant = cfg->precision == CGVC_PREC_F16F8;
    e->tcw.quant_bwd = e->tcw.quant && cfg->train;           // training in that precision also needs the data-gradient planes
    e->tcw.wgrad16 = e->tcw.quant;                           // weight gradients (leaves of the graph) from the fp16 planes alone; option "wgrad_f16"
    int r = tc_alloc(e->tcw);
    if (r != 0) { std::string m = cudaGetErrorString((cudaError_t)r); cudaFree(e->d_scalars); delete e; return fail(nullptr, CGVC_ERR_CUDA, "tc_alloc: %s", m.c_str()); }
  }
  *out = e;
  return 0;
}

int cgvc_destroy(cgvc_handle e) {
  if (!e) return 0;
  DeviceGuard dguard; dguard.set(e->cfg.device);
  if (e->comm && e->nccl.CommDestroy) e->nccl.CommDestroy(e->comm);
  tc_free(e->tcw);
  for (int l = 0; l < 2; ++l) { if (e->lane_stream[l]) cudaStreamDestroy(e->lane_stream[l]); if (e->ev_join[l]) cudaEventDestroy(e->ev_join[l]); }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->comm_stream) cudaStreamDestroy(e->comm_stream);
  if (e->ev_grads) cudaEventDestroy(e->ev_grads);
  for (int k = 0; k < 4; ++k) if (e->ev_ar[k]) cudaEventDestroy(e->ev_ar[k]);
  for (int l = 0; l < 2; ++l) {
    if (e->sideq[l].side) cudaStreamDestroy(e->sideq[l].side);
    for (int b = 0; b < 2; ++b) { if (e->sideq[l].ready[b]) cudaEventDestroy(e->sideq[l].ready[b]); if (e->sideq[l].done[b]) cudaEventDestroy(e->sideq[l].done[b]); }
  }
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
  if (e->stage) cudaFree(e->stage);
  if (e->graph_stream) cudaStreamDestroy(e->graph_stream);
  if (e->ev_bridge) cudaEventDestroy(e->ev_bridge);
  if (e->ev_bridge2) cudaEventDestroy(e->ev_bridge2);
  cudaFree(e->d_scalars);
  delete e;
  return 0;
}

int cgvc_arena_bytes(cgvc_handle e, int arena, size_t* bytes) {
  if (!e || !bytes || arena < 0 || arena >= CGVC_ARENA_COUNT) return fail(e, CGVC_ERR_ARG, "cgvc_arena_bytes: bad argument");
  if (arena == CGVC_ARENA_WORK) *bytes = work_bytes_needed(e);
  else *bytes = ((e->n_params * sizeof(float)) + 255) & ~(size_t)255;
  return 0;
}

int cgvc_bind_arena(cgvc_handle e, int arena, void* p, size_t bytes) {
  if (!e || arena < 0 || arena >= CGVC_ARENA_COUNT) return fail(e, CGVC_ERR_ARG, "cgvc_bind_arena: bad argument");
  size_t need; cgvc_arena_bytes(e, arena, &need);
  if (!p || bytes < need) return fail(e, CGVC_ERR_UNBOUND, "arena %d needs %zu bytes, got %zu", arena, need, bytes);
  if ((uintptr_t)p & 255) return fail(e, CGVC_ERR_ARG, "arena %d must be 256-byte aligned", arena);
  e->arena[arena] = p; e->arena_bytes[arena] = bytes;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);   // captured graphs hold the old addresses
  e->graphs.clear();
  return 0;
}

int cgvc_param_count(cgvc_handle e, int* n_tensors, size_t* n_elements) {
  if (!e) return CGVC_ERR_ARG;
  if (n_tensors) *n_tensors = (int)e->tensors.size();
  if (n_elements) *n_elements = e->n_real_params;
  return 0;
}

int cgvc_param_info(cgvc_handle e, int index, const char** name, size_t* offset, int* ndim, int shape_out[4]) {
  if (!e || index < 0 || index >= (int)e->tensors.size()) return fail(e, CGVC_ERR_ARG, "cgvc_param_info: index out of range");
  const TensorInfo& t = e->tensors[index];
  if (name) *name = t.name.c_str();
  if (offset) *offset = t.off;
  if (ndim) *ndim = t.ndim;
  if (shape_out) for (int i = 0; i < 4; ++i) shape_out[i] = t.shape[i];
  return 0;
}

static int need_arenas(cgvc_engine* e, bool train) {
  if (!e->arena[CGVC_ARENA_PARAM] || !e->arena[CGVC_ARENA_WORK]) return fail(e, CGVC_ERR_UNBOUND, "PARAM and WORK arenas must be bound");
  if (train && (!e->arena[CGVC_ARENA_GRAD])) return fail(e, CGVC_ERR_UNBOUND, "GRAD arena must be bound");
  return 0;
}

int cgvc_params_updated(cgvc_handle e, void* stream) {
  if (!e) return CGVC_ERR_ARG;
  if (!e->arena[CGVC_ARENA_PARAM]) return fail(e, CGVC_ERR_UNBOUND, "PARAM arena must be bound");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  if (e->cfg.precision != CGVC_PREC_FP32_SIMT) {
    int r = tc_refresh_weights(e->tcw, e->P(), (cudaStream_t)stream);
    if (r != 0) return fail(e, CGVC_ERR_CUDA, "tc_refresh_weights: %s", cudaGetErrorString((cudaError_t)r));
  }
  return 0;
}

int cgvc_set_adam_step(cgvc_handle e, long long t) { if (!e || t < 0) return CGVC_ERR_ARG; e->adam_t = t; return 0; }
int cgvc_get_adam_step(cgvc_handle e, long long* t) { if (!e || !t) return CGVC_ERR_ARG; *t = e->adam_t; return 0; }

static int check_bt(cgvc_engine* e, int batch, int frames, int mult) {
  if (batch < 1 || batch > e->cfg.max_batch) return fail(e, CGVC_ERR_ARG, "batch %d outside [1, %d]", batch, e->cfg.max_batch);
  if (frames < mult || frames % mult != 0 || frames > e->cfg.max_frames)
    return fail(e, CGVC_ERR_ARG, "frames %d must be a multiple of %d in [%d, %d]", frames, mult, mult, e->cfg.max_frames);
  return 0;
}

int cgvc_generator_forward(cgvc_handle e, int direction, const float* in_dev, float* out_dev, int batch, int frames, void* stream) {
  if (!e) return CGVC_ERR_ARG;
  if (direction != 0 && direction != 1) return fail(e, CGVC_ERR_DIRECTION, "Conversion direction must be specified.");
  if (!in_dev || !out_dev) return fail(e, CGVC_ERR_ARG, "null buffer");
  RET(check_bt(e, batch, frames, 4));
  RET(need_arenas(e, false));
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  Bump ws; ws.reset(e->arena[CGVC_ARENA_WORK], e->arena_bytes[CGVC_ARENA_WORK]);
  FwdPlan F; F.in_cl = ws.take<float>((size_t)batch * e->cfg.num_features * frames);
  plan_generator(e, ws, F.g, batch, frames);
  if (ws.overflow) return fail(e, CGVC_ERR_UNBOUND, "WORK arena too small");
  CK(launch_transpose_ft(in_dev, F.in_cl, batch, e->cfg.num_features, frames, st));
  if (!e->debug_taps) e->taps.clear();
  RET(generator_forward(e, e->gen[direction], F.g, F.in_cl, st, e->debug_taps != 0, false));
  CK(launch_transpose_ft(F.g.out_cl, out_dev, batch, frames, e->cfg.num_features, st));
  return 0;
}

int cgvc_discriminator_forward(cgvc_handle e, int which, const float* in_dev, float* out_dev, int batch, int frames, void* stream) {
  if (!e) return CGVC_ERR_ARG;
  if (which != 0 && which != 1) return fail(e, CGVC_ERR_ARG, "which must be 0 (discriminator_A) or 1 (discriminator_B)");
  if (!in_dev || !out_dev) return fail(e, CGVC_ERR_ARG, "null buffer");
  RET(check_bt(e, batch, frames, 16));
  RET(need_arenas(e, false));
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  Bump ws; ws.reset(e->arena[CGVC_ARENA_WORK], e->arena_bytes[CGVC_ARENA_WORK]);
  FwdPlan F; F.in_cl = ws.take<float>((size_t)batch * e->cfg.num_features * frames);
  plan_generator(e, ws, F.g, batch, frames);   // keep layout identical to work_bytes_needed
  plan_discriminator(e, ws, F.d, batch, frames);
  if (ws.overflow) return fail(e, CGVC_ERR_UNBOUND, "WORK arena too small");
  RET(discriminator_forward(e, e->disc[which], F.d, in_dev, st, true));
  CK(cudaMemcpyAsync(out_dev, F.d.prob, (size_t)batch * (e->cfg.num_features / 4) * (frames / 16) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int cgvc_debug_activation(cgvc_handle e, const char* name, float* out_dev, size_t capacity, size_t* n_out, void* stream) {
  if (!e || !name) return CGVC_ERR_ARG;
  auto it = e->taps.find(name);
  if (it == e->taps.end()) return fail(e, CGVC_ERR_ARG, "no activation tap named '%s' (generator taps need the 'debug_taps' option set before the forward call)", name);
  if (n_out) *n_out = it->second.second;
  if (out_dev) {
    if (capacity < it->second.second) return fail(e, CGVC_ERR_ARG, "tap '%s' needs %zu elements", name, it->second.second);
    CK(cudaMemcpyAsync(out_dev, it->second.first, it->second.second * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return 0;
}

// forward + losses + backward of one training step; leaves gradients in GRAD (model.py:44-90,107-108)
// one lane of the step (see LanePlan), enqueued on stream st
static int run_lane(cgvc_engine* e, LanePlan& L, int lane, const float* Yreal_dev, int B, int T, float lambda_identity,
                    float* gen_out_dev, cudaStream_t st) {
  const int nf = e->cfg.num_features;
  const size_t img = (size_t)B * nf * T;
  float* Gm = e->G(); const float* Pm = e->P();
  float* sc = e->d_scalars; float* Ls = sc + 8;
  const GenNet& Gfirst = e->gen[lane];          // lane 0: generator_A2B ; lane 1: generator_B2A
  const GenNet& Gcyc = e->gen[1 - lane];
  const DiscNet& DN = e->disc[1 - lane];        // lane 0 judges domain B (discriminator_B); lane 1 domain A
  const float* X_cl = L.in; const float* Y_cl = L.in + img;
  // ---- forward (model.py:44-54,75-78) ----
  RET(generator_forward(e, Gfirst, L.gfirst, L.in, st, false));             // [gen_Y ; id_Y]
  const float* genY_cl = L.gfirst.out_cl; const float* idY_cl = L.gfirst.out_cl + img;
  RET(generator_forward(e, Gcyc, L.gcyc, genY_cl, st, false));              // cycle_X
  CK(cudaMemcpyAsync(L.din, Yreal_dev, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(launch_transpose_ft(genY_cl, L.din + img, B, T, nf, st));
  if (gen_out_dev) CK(cudaMemcpyAsync(gen_out_dev, L.din + img, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RET(discriminator_forward(e, DN, L.d, L.din, st, false));
  // ---- losses and their gradients (model.py:57-90) ----
  const float ls = loss_scale(e, B);                            // scales every gradient of the step (not the loss values); Adam divides it out
  CK(launch_l1_loss_grad(L.gcyc.out_cl, X_cl, (long long)img, Ls + 0, sc + 0, L.d_cyc, 0, st, ls));          // cycle term
  CK(launch_l1_loss_grad(idY_cl, Y_cl, (long long)img, Ls + 1, sc + 1, L.d_out + img, 0, st, ls));           // identity term
  const long long hrows = (long long)B * (nf / 4) * (T / 16);   // head rows per half
  const float* Y3 = L.d.d[2].Y;
  float* Dslot = Ls + (lane == 0 ? 6 : 5);                      // discriminator_loss_B / _A
  float* Gslot = Ls + (lane == 0 ? 2 : 3);                      // generator_loss_A2B / _B2A
  // discriminator loss: real half -> target 1, fake half -> target 0, each weighted 1/2 (model.py:81-88)
  CK(launch_head_loss_bwd(L.d.prob, Y3, hrows, 1024, Pm + DN.dense_k, 1.f, 0.5f, Dslot, L.dY3, Gm + DN.dense_k, Gm + DN.dense_b, st, ls));
  CK(launch_head_loss_bwd(L.d.prob + hrows, Y3 + hrows * 1024, hrows, 1024, Pm + DN.dense_k, 0.f, 0.5f, Dslot,
                          L.dY3 + hrows * 1024, Gm + DN.dense_k, Gm + DN.dense_b, st, ls));
  RET(discriminator_backward(e, DN, L.d, L.dY3, true, nullptr, L.S, st));
  // generator adversarial loss on the fake half: target 1 (model.py:68-69); the gradient flows to the fake only
  DiscActs V = disc_view(e, L.d, B, B);
  CK(launch_head_loss_bwd(V.prob, V.d[2].Y, hrows, 1024, Pm + DN.dense_k, 1.f, 1.f, Gslot, L.dY3, nullptr, nullptr, st, ls));
  RET(discriminator_backward(e, DN, V, L.dY3, false, L.d_adv, L.S, st));
  // ---- generator backward ----
  // cycle pass: G_{Y->X}(gen_Y) <- d cycle_X ; its input gradient is the first half of the first pass's upstream
  RET(generator_backward(e, Gcyc, L.gcyc, L.d_cyc, L.d_out, L.S, st));
  CK(launch_transpose_ft(L.d_adv, L.d_cyc, B, nf, T, st));      // adversarial gradient [B,24,T] -> channels-last (d_cyc is free now)
  CK(launch_add(L.d_out, L.d_cyc, L.d_out, (long long)img, st));
  if (lambda_identity == 0.f) {
    // train.py:98-99 switches the identity loss off after 10k iterations (it is still computed and logged, model.py:157):
    // its upstream gradient is then exactly zero, so only the gen_Y half of the first pass is back-propagated
    GenActs half = L.gfirst; half.n = B;
    RET(generator_backward(e, Gfirst, half, L.d_out, nullptr, L.S, st));
  } else {
    RET(generator_backward(e, Gfirst, L.gfirst, L.d_out, nullptr, L.S, st));
  }
  side_join(L.S, st);                                        // the side-stream weight gradients rejoin the lane
  return 0;
}

// forward + losses + backward of one training step; leaves gradients in GRAD (model.py:44-90,107-108)
static int forward_backward(cgvc_engine* e, const float* A_dev, const float* B_dev, int B, int T, float lc, float li,
                            float* gen_A_dev, float* gen_B_dev, float* losses_dev, cudaStream_t st) {
  const int nf = e->cfg.num_features;
  RET(check_bt(e, B, T, 16));
  if (!e->cfg.train) return fail(e, CGVC_ERR_ARG, "engine was created with train = 0");
  RET(need_arenas(e, true));
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  Bump ws; ws.reset(e->arena[CGVC_ARENA_WORK], e->arena_bytes[CGVC_ARENA_WORK]);
  TrainPlan P; plan_train(e, ws, P, B, T);
  if (ws.overflow) return fail(e, CGVC_ERR_UNBOUND, "WORK arena too small for batch %d x %d frames", B, T);
  const size_t img = (size_t)B * nf * T;
  float* sc = e->d_scalars; float* L = sc + 8;
  (void)lc;                                                  // lambdas are already in d_scalars[0..1] (set_step_scalars)
  CK(cudaMemsetAsync(L, 0, 8 * sizeof(float), st));
  CK(cudaMemsetAsync(e->G(), 0, e->n_params * sizeof(float), st));
  // channels-last copies of the real samples: lane 0 reads [A;B], lane 1 [B;A]
  CK(launch_transpose_ft(A_dev, P.lane[0].in, B, nf, T, st));
  CK(launch_transpose_ft(B_dev, P.lane[0].in + img, B, nf, T, st));
  CK(cudaMemcpyAsync(P.lane[1].in, P.lane[0].in + img, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(P.lane[1].in + img, P.lane[0].in, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < 2; ++l) {
    SideQ& q = e->sideq[l];
    q.on = e->side_wgrad && e->two_streams && !e->fuse_bwd && !tc_profile_is_on() && q.side != nullptr;
    q.used[0] = q.used[1] = false; q.cur = 0;
  }
  if (e->two_streams) {
    // fork
    CK(cudaEventRecord(e->ev_fork, st));
    for (int l = 0; l < 2; ++l) CK(cudaStreamWaitEvent(e->lane_stream[l], e->ev_fork, 0));
    RET(run_lane(e, P.lane[0], 0, B_dev, B, T, li, gen_B_dev, e->lane_stream[0]));
    RET(run_lane(e, P.lane[1], 1, A_dev, B, T, li, gen_A_dev, e->lane_stream[1]));
    // join
    for (int l = 0; l < 2; ++l) { CK(cudaEventRecord(e->ev_join[l], e->lane_stream[l])); CK(cudaStreamWaitEvent(st, e->ev_join[l], 0)); }
  } else {
    RET(run_lane(e, P.lane[0], 0, B_dev, B, T, li, gen_B_dev, st));
    RET(run_lane(e, P.lane[1], 1, A_dev, B, T, li, gen_A_dev, st));
  }
  CK(launch_finalize_losses(L, sc, st));
  if (losses_dev) CK(cudaMemcpyAsync(losses_dev, L, 8 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"

// d_scalars: [0..1] lambda_cycle, lambda_identity; [2..3] generator Adam (lr_t, grad_scale); [4..5] discriminator Adam
static int set_lambdas(cgvc_engine* e, float lc, float li, cudaStream_t st) {
  float v[6] = {lc, li, 0, 0, 0, 0};
  CK(launch_set_scalars(e->d_scalars, 0, 2, v, st));
  return 0;
}
static int set_adam_scalars(cgvc_engine* e, float lr_g, float lr_d, float grad_scale, cudaStream_t st) {
  e->adam_t += 1;                                            // both optimizers advance once per train() (Appendix A.6)
  double t = (double)e->adam_t;
  double corr = sqrt(1.0 - pow((double)ADAM_B2, t)) / (1.0 - pow((double)ADAM_B1, t));
  float v[6] = {(float)(lr_g * corr), grad_scale, (float)(lr_d * corr), grad_scale, 0, 0};
  CK(launch_set_scalars(e->d_scalars, 2, 4, v, st));
  return 0;
}
// the capturable part of the optimizer step: two Adam ranges + refresh of the tensor-core weight planes
static int adam_body(cgvc_engine* e, cudaStream_t st) {
  float* p = e->P(); float* g = e->G(); float* m = (float*)e->arena[CGVC_ARENA_ADAM_M]; float* v = (float*)e->arena[CGVC_ARENA_ADAM_V];
  size_t gend = e->gen[1].end;   // generators occupy [0, gend), discriminators [gend, n_params)  (model.py:94-95)
  CK(launch_adam(p, g, m, v, (long long)gend, e->d_scalars + 2, ADAM_B1, ADAM_B2, ADAM_EPS, st));
  CK(launch_adam(p + gend, g + gend, m + gend, v + gend, (long long)(e->n_params - gend), e->d_scalars + 4, ADAM_B1, ADAM_B2, ADAM_EPS, st));
  return cgvc_params_updated(e, (void*)st);
}

// ---- CUDA graphs: the ~650 launches of a step are captured once per (buffers, batch, frames, identity-on/off) and replayed.
// Everything inside the captured bodies reads its per-step scalars from d_scalars (written by the eager set_scalars kernel),
// so a replay is exact.  Capture needs a non-legacy stream: work arriving on the legacy default stream is bridged with events.
template <class Body>
static int run_captured(cgvc_engine* e, const GraphKey& key, cudaStream_t user, Body body) {
  if (!e->use_graphs || tc_profile_is_on()) return body(user);
  cudaStream_t st = user;
  const bool bridge = (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread);
  if (bridge) { st = e->graph_stream; CK(cudaEventRecord(e->ev_bridge, user)); CK(cudaStreamWaitEvent(st, e->ev_bridge, 0)); }
  auto it = e->graphs.find(key);
  GraphEntry ent{nullptr, 0};
  if (it != e->graphs.end()) ent = it->second;
  else {
    if (e->graphs.size() >= 16) { for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec); e->graphs.clear(); }
    cudaGraph_t graph = nullptr;
    const unsigned long long before = g_cgvc_launches;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) { cudaGetLastError(); e->use_graphs = 0; return body(user); }
    int rc = body(st);
    ce = cudaStreamEndCapture(st, &graph);
    ent.launches = g_cgvc_launches - before;                 // kernels recorded, not run: counted per replay below
    g_cgvc_launches = before;
    if (rc != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc != 0 && ce == cudaSuccess) return rc;           // the body itself refused (bad argument, arena too small): not a capture problem
      e->use_graphs = 0;                                     // fall back to eager launches for the rest of this engine's life
      return body(user);
    }
    ce = cudaGraphInstantiate(&ent.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { cudaGetLastError(); e->use_graphs = 0; return body(user); }
    e->graphs[key] = ent;
  }
  cudaGraphExec_t exec = ent.exec;
  g_cgvc_launches += ent.launches;
  CK(cudaGraphLaunch(exec, st));
  if (bridge) { CK(cudaEventRecord(e->ev_bridge2, st)); CK(cudaStreamWaitEvent(user, e->ev_bridge2, 0)); }
  return 0;
}

extern "C" {

int cgvc_compute_gradients(cgvc_handle e, const float* A_dev, const float* B_dev, int batch, int frames,
                           float lambda_cycle, float lambda_identity, float* gen_A_dev, float* gen_B_dev, float* losses_dev, void* stream) {
  if (!e || !A_dev || !B_dev) return fail(e, CGVC_ERR_ARG, "null argument");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  RET(set_lambdas(e, lambda_cycle, lambda_identity, (cudaStream_t)stream));
  RET(forward_backward(e, A_dev, B_dev, batch, frames, lambda_cycle, lambda_identity, gen_A_dev, gen_B_dev, losses_dev, (cudaStream_t)stream));
  const float ls = loss_scale(e, batch);                     // the gradients are handed out, not fed to Adam: remove the loss scale here
  if (ls != 1.f) CK(launch_scale(e->G(), (long long)e->n_params, 1.f / ls, (cudaStream_t)stream));
  return 0;
}

int cgvc_adam_step(cgvc_handle e, float lr_g, float lr_d, float grad_scale, void* stream) {
  if (!e) return CGVC_ERR_ARG;
  for (int a = 0; a < 4; ++a) if (!e->arena[a]) return fail(e, CGVC_ERR_UNBOUND, "PARAM/GRAD/ADAM_M/ADAM_V arenas must be bound");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  const long long t_before = e->adam_t;
  RET(set_adam_scalars(e, lr_g, lr_d, grad_scale, (cudaStream_t)stream));
  int r = adam_body(e, (cudaStream_t)stream);
  if (r != 0) e->adam_t = t_before;
  return r;
}

int cgvc_train_step(cgvc_handle e, const float* A_dev, const float* B_dev, int batch, int frames,
                    float lambda_cycle, float lambda_identity, float lr_g, float lr_d,
                    float* gen_A_dev, float* gen_B_dev, float* losses_dev, void* stream) {
  if (!e || !A_dev || !B_dev) return fail(e, CGVC_ERR_ARG, "null argument");
  for (int a = 0; a < 4; ++a) if (!e->arena[a]) return fail(e, CGVC_ERR_UNBOUND, "PARAM/GRAD/ADAM_M/ADAM_V arenas must be bound");
  RET(check_bt(e, batch, frames, 16));
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const float gscale = (e->comm ? 1.f / (float)e->nranks : 1.f) / loss_scale(e, batch);
  RET(set_lambdas(e, lambda_cycle, lambda_identity, st));
  // the Adam step counter advances only if the whole step was enqueued: a call that is refused further down (WORK arena too small,
  // capture failure, NCCL error) must not change the bias correction of the next one
  const long long adam_t_before = e->adam_t;
  struct Rollback { cgvc_engine* e; long long t; bool armed; ~Rollback() { if (armed) e->adam_t = t; } } rollback{e, adam_t_before, true};
  RET(set_adam_scalars(e, lr_g, lr_d, gscale, st));
  if (e->use_graphs && !tc_profile_is_on()) {
    // the graphs read the inputs from fixed staging buffers and leave the results in the WORK arena / d_scalars, so one
    // captured graph serves any caller pointers; the copies either side are eager
    const size_t img = (size_t)batch * e->cfg.num_features * frames;
    const size_t cap = (size_t)e->cfg.max_batch * e->cfg.num_features * e->cfg.max_frames;
    if (!e->stage) CK(cudaMalloc(&e->stage, 2 * cap * sizeof(float)));
    float* sA = e->stage; float* sB = e->stage + cap;
    CK(cudaMemcpyAsync(sA, A_dev, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(sB, B_dev, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GraphKey key; memset(&key, 0, sizeof key);
    key.batch = batch; key.frames = frames; key.id_off = lambda_identity == 0.f; key.lanes = e->two_streams; key.fuse = e->fuse_in | (e->fuse_bwd << 1) | (e->side_wgrad << 2) | (e->fuse_c1 << 3) | (e->edge_lower << 4); key.kind = 0;
    RET(run_captured(e, key, st, [&](cudaStream_t s) {
      return forward_backward(e, sA, sB, batch, frames, lambda_cycle, lambda_identity, nullptr, nullptr, nullptr, s);
    }));
    if (gen_A_dev || gen_B_dev) {
      Bump ws; ws.reset(e->arena[CGVC_ARENA_WORK], e->arena_bytes[CGVC_ARENA_WORK]);
      TrainPlan P; plan_train(e, ws, P, batch, frames);
      if (gen_B_dev) CK(cudaMemcpyAsync(gen_B_dev, P.lane[0].din + img, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (gen_A_dev) CK(cudaMemcpyAsync(gen_A_dev, P.lane[1].din + img, img * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (losses_dev) CK(cudaMemcpyAsync(losses_dev, e->d_scalars + 8, 8 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    RET(forward_backward(e, A_dev, B_dev, batch, frames, lambda_cycle, lambda_identity, gen_A_dev, gen_B_dev, losses_dev, st));
  }
  if (e->comm && e->pipelined_comm && e->comm_stream) {
    // one all-reduce per network (arena order), all enqueued on the communication stream behind the step's gradients; the caller's
    // stream then takes the networks one by one: wait for its all-reduce, Adam over its range, refresh of its tensor-core planes
    // the four networks in arena order; every tensor starts on a 16-byte boundary, so the ranges are cut at the aligned start of each
    // network's first tensor (the up-to-3 padding floats in front of it belong to the previous range and hold zero gradients)
    auto al4 = [](size_t v) { return (v + 3) & ~(size_t)3; };
    const size_t cut[5] = {0, al4(e->gen[1].begin), al4(e->disc[0].begin), al4(e->disc[1].begin), e->n_params};
    struct Range { size_t b, n; const float* hyper; } rg[4];
    for (int k = 0; k < 4; ++k) { rg[k].b = cut[k]; rg[k].n = cut[k + 1] - cut[k]; rg[k].hyper = e->d_scalars + (k < 2 ? 2 : 4); }
    CK(cudaEventRecord(e->ev_grads, st));
    CK(cudaStreamWaitEvent(e->comm_stream, e->ev_grads, 0));
    for (int k = 0; k < 4; ++k) {
      int r = e->nccl.AllReduce(e->G() + rg[k].b, e->G() + rg[k].b, rg[k].n, 7, 0, e->comm, e->comm_stream);     // ncclFloat32, ncclSum
      if (r != 0) return fail(e, CGVC_ERR_NCCL, "ncclAllReduce: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(r) : "?");
      CK(cudaEventRecord(e->ev_ar[k], e->comm_stream));
    }
    float* pp = e->P(); float* gg = e->G(); float* mm = (float*)e->arena[CGVC_ARENA_ADAM_M]; float* vv = (float*)e->arena[CGVC_ARENA_ADAM_V];
    for (int k = 0; k < 4; ++k) {
      CK(cudaStreamWaitEvent(st, e->ev_ar[k], 0));
      GraphKey kk; memset(&kk, 0, sizeof kk); kk.kind = 2 + k;
      const Range R = rg[k];
      RET(run_captured(e, kk, st, [&](cudaStream_t s) {
        CK(launch_adam(pp + R.b, gg + R.b, mm + R.b, vv + R.b, (long long)R.n, R.hyper, ADAM_B1, ADAM_B2, ADAM_EPS, s));
        if (e->cfg.precision != CGVC_PREC_FP32_SIMT) {
          int r = tc_refresh_weights_range(e->tcw, pp, R.b, R.b + R.n, s);
          if (r != 0) return fail(e, CGVC_ERR_CUDA, "tc_refresh_weights: %s", cudaGetErrorString((cudaError_t)r));
        }
        return 0;
      }));
    }
    rollback.armed = false;
    return 0;
  }
  if (e->comm) RET(cgvc_allreduce_grads(e, stream));
  GraphKey k2; memset(&k2, 0, sizeof k2); k2.kind = 1;
  RET(run_captured(e, k2, st, [&](cudaStream_t s) { return adam_body(e, s); }));
  rollback.armed = false;
  return 0;
}

// ---- NCCL --------------------------------------------------------------------------------------------------
static int load_nccl(cgvc_engine* e) {
  if (e->nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) return fail(e, CGVC_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  e->nccl.lib = lib;
  *(void**)&e->nccl.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
  *(void**)&e->nccl.CommInitRank = dlsym(lib, "ncclCommInitRank");
  *(void**)&e->nccl.CommDestroy = dlsym(lib, "ncclCommDestroy");
  *(void**)&e->nccl.AllReduce = dlsym(lib, "ncclAllReduce");
  *(void**)&e->nccl.GetErrorString = dlsym(lib, "ncclGetErrorString");
  if (!e->nccl.GetUniqueId || !e->nccl.CommInitRank || !e->nccl.AllReduce || !e->nccl.CommDestroy)
    return fail(e, CGVC_ERR_NCCL, "libnccl is missing required symbols");
  return 0;
}

int cgvc_comm_unique_id(cgvc_handle e, void* id128_host) {
  if (!e || !id128_host) return CGVC_ERR_ARG;
  RET(load_nccl(e));
  int r = e->nccl.GetUniqueId(id128_host);
  if (r != 0) return fail(e, CGVC_ERR_NCCL, "ncclGetUniqueId: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(r) : "?");
  return 0;
}

int cgvc_comm_init(cgvc_handle e, const void* id128_host, int rank, int nranks) {
  if (!e || !id128_host || nranks < 1 || rank < 0 || rank >= nranks) return fail(e, CGVC_ERR_ARG, "cgvc_comm_init: bad argument");
  RET(load_nccl(e));
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  Id128 id; memcpy(id.b, id128_host, 128);
  int r = e->nccl.CommInitRank(&e->comm, nranks, id, rank);
  if (r != 0) { e->comm = nullptr; return fail(e, CGVC_ERR_NCCL, "ncclCommInitRank: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(r) : "?"); }
  e->rank = rank; e->nranks = nranks;
  return 0;
}

int cgvc_comm_destroy(cgvc_handle e) {
  if (!e) return CGVC_ERR_ARG;
  if (e->comm) { e->nccl.CommDestroy(e->comm); e->comm = nullptr; e->nranks = 1; e->rank = 0; }
  return 0;
}

int cgvc_allreduce_grads(cgvc_handle e, void* stream) {
  if (!e) return CGVC_ERR_ARG;
  if (!e->comm) return fail(e, CGVC_ERR_NCCL, "no communicator attached");
  if (!e->arena[CGVC_ARENA_GRAD]) return fail(e, CGVC_ERR_UNBOUND, "GRAD arena must be bound");
  // ncclFloat32 = 7, ncclSum = 0
  int r = e->nccl.AllReduce(e->G(), e->G(), e->n_params, 7, 0, e->comm, (cudaStream_t)stream);
  if (r != 0) return fail(e, CGVC_ERR_NCCL, "ncclAllReduce: %s", e->nccl.GetErrorString ? e->nccl.GetErrorString(r) : "?");
  return 0;
}

int cgvc_set_option(cgvc_handle e, const char* name, int value) {
  if (!e || !name) return CGVC_ERR_ARG;
  if (!strcmp(name, "two_streams")) { e->two_streams = value != 0; return 0; }
  if (!strcmp(name, "fuse_in")) { e->fuse_in = value != 0; return 0; }
  if (!strcmp(name, "fuse_bwd")) { e->fuse_bwd = value != 0; return 0; }
  if (!strcmp(name, "edge_lower")) { e->edge_lower = value != 0; return 0; }
  if (!strcmp(name, "side_wgrad")) { e->side_wgrad = value != 0; return 0; }
  if (!strcmp(name, "pipelined_comm")) { e->pipelined_comm = value != 0; return 0; }
  if (!strcmp(name, "fuse_c1")) { e->fuse_c1 = value != 0; return 0; }
  if (!strcmp(name, "debug_taps")) { e->debug_taps = value != 0; return 0; }
  if (!strcmp(name, "cuda_graph")) { e->use_graphs = value != 0; return 0; }
  if (!strcmp(name, "tc_debug")) { tc_set_debug(value); return 0; }
  if (!strcmp(name, "post_onepass")) {                       // process-wide, like cta_pairs
    post_set_onepass(value);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    return 0;
  }
  if (!strcmp(name, "wgrad_f16")) {                          // F16F8 only: weight gradients from the fp16 planes alone
    e->tcw.wgrad16 = value != 0;
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    return 0;
  }
  if (!strcmp(name, "post_stream")) {                        // process-wide, like post_onepass
    post_set_stream(value);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    return 0;
  }
  if (!strcmp(name, "prep_batched")) {                       // process-wide; the captured Adam + refresh graphs hold the old kernels
    tc_set_prep_batched(value);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    return 0;
  }
  if (!strcmp(name, "cta_pairs")) {                          // process-wide switch; captured graphs hold the old kernels
    tc_set_pair(value);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
    return 0;
  }
  return fail(e, CGVC_ERR_ARG, "unknown option '%s'", name);
}
int cgvc_kernel_launches(unsigned long long* count) { if (!count) return CGVC_ERR_ARG; *count = g_cgvc_launches; return 0; }
int cgvc_profile_enable(int on) { tc_profile_enable(on); return 0; }
int cgvc_profile_launches(double* ms, double* flops, long long* meta4, int capacity, int* n_out) {
  if (!ms || !flops || !meta4 || capacity < 0) return CGVC_ERR_ARG;
  return tc_profile_launches(ms, flops, meta4, capacity, n_out) == 0 ? 0 : CGVC_ERR_CUDA;
}
int cgvc_profile_collect(double* ms3, double* flops3, long long* launches3) {
  if (!ms3 || !flops3 || !launches3) return CGVC_ERR_ARG;
  return tc_profile_collect(ms3, flops3, launches3) == 0 ? 0 : CGVC_ERR_CUDA;
}

// ---- device-resident training data ---------------------------------------------------------------------------
int cgvc_sample_plan(cgvc_handle e, const long long* offsets_A_dev, int n_A, const long long* offsets_B_dev, int n_B,
                     unsigned long long seed, long long epoch, int crop_frames, int* plan_dev, int* err_dev, void* stream) {
  if (!e || !offsets_A_dev || !offsets_B_dev || !plan_dev || !err_dev) return fail(e, CGVC_ERR_ARG, "null argument");
  if (n_A < 1 || n_B < 1 || crop_frames < 1 || epoch < 0) return fail(e, CGVC_ERR_ARG, "cgvc_sample_plan: empty corpus or bad crop / epoch");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  CK(cudaMemsetAsync(err_dev, 0, sizeof(int), (cudaStream_t)stream));
  CK(launch_sample_plan(offsets_A_dev, n_A, offsets_B_dev, n_B, seed, epoch, crop_frames, plan_dev, err_dev, (cudaStream_t)stream));
  return 0;
}

int cgvc_gather_minibatch(cgvc_handle e, const float* corpus_A_dev, const long long* offsets_A_dev, const float* corpus_B_dev,
                          const long long* offsets_B_dev, const int* plan_dev, int num_pairs, int first_pair, int batch, int crop_frames,
                          float* A_out_dev, float* B_out_dev, void* stream) {
  if (!e || !corpus_A_dev || !corpus_B_dev || !offsets_A_dev || !offsets_B_dev || !plan_dev || !A_out_dev || !B_out_dev)
    return fail(e, CGVC_ERR_ARG, "null argument");
  if (batch < 1 || first_pair < 0 || first_pair + batch > num_pairs || crop_frames < 1)
    return fail(e, CGVC_ERR_ARG, "cgvc_gather_minibatch: pairs [%d, %d) outside the epoch's %d", first_pair, first_pair + batch, num_pairs);
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  CK(launch_gather_minibatch(corpus_A_dev, offsets_A_dev, corpus_B_dev, offsets_B_dev, plan_dev, num_pairs, first_pair, batch,
                             e->cfg.num_features, crop_frames, A_out_dev, B_out_dev, (cudaStream_t)stream));
  return 0;
}

// ---- per-kernel entry points ---------------------------------------------------------------------------------
int cgvc_conv_forward(cgvc_handle e, int precision, const float* x, const float* w, const float* bias, float* y,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, void* stream) {
  if (!e || !x || !w || !y) return fail(e, CGVC_ERR_ARG, "null argument");
  if (kh * kw > CGVC_MAX_TAPS) return fail(e, CGVC_ERR_UNSUPPORTED, "at most %d filter taps", CGVC_MAX_TAPS);
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != CGVC_PREC_FP32_SIMT) {
    int r = tc_conv_fwd_adhoc(precision, x, w, bias, y, B, H, W, Cin, kh, kw, Cout, sh, sw, st);
    if (r == TC_UNSUPPORTED) return fail(e, CGVC_ERR_UNSUPPORTED, "shape not supported by the tensor-core path");
    if (r != 0) return fail(e, CGVC_ERR_CUDA, "tc conv fwd: %s", cudaGetErrorString((cudaError_t)r));
    return 0;
  }
  GatherGeom g = fwd_geom(B, H, W, kh, kw, sh, sw);
  GemmOperands op; memset(&op, 0, sizeof op);
  op.src = x; op.s_ld = Cin; op.C = Cin; op.w = w; op.w_ts = (long long)Cin * Cout; op.w_cs = Cout; op.w_ns = 1; op.N = Cout;
  op.dst = y; op.d_ld = Cout; op.bias = bias;
  CK(launch_gg_simt(g, op, st));
  return 0;
}

int cgvc_conv_backward(cgvc_handle e, int precision, const float* x, const float* w, const float* dy,
                       float* dx, float* dw, float* dbias, int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, void* stream) {
  if (!e || !x || !w || !dy) return fail(e, CGVC_ERR_ARG, "null argument");
  if (kh * kw > CGVC_MAX_TAPS) return fail(e, CGVC_ERR_UNSUPPORTED, "at most %d filter taps", CGVC_MAX_TAPS);
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != CGVC_PREC_FP32_SIMT) {
    int r = tc_conv_bwd_adhoc(precision, x, w, dy, dx, dw, dbias, B, H, W, Cin, kh, kw, Cout, sh, sw, st, e->tcw.wgrad16 ? 1 : 0);
    if (r == TC_UNSUPPORTED) return fail(e, CGVC_ERR_UNSUPPORTED, "shape not supported by the tensor-core path");
    if (r != 0) return fail(e, CGVC_ERR_CUDA, "tc conv bwd: %s", cudaGetErrorString((cudaError_t)r));
    return 0;
  }
  ConvW c; c.k = 0; c.b = 0; c.kh = kh; c.kw = kw; c.cin = Cin; c.cout = Cout;
  if (dx) RET(conv_dgrad_simt(e, w, c, sh, sw, B, H, W, dy, Cout, 0, dx, 0, st));
  if (dw) {
    GatherGeom g = fwd_geom(B, H, W, kh, kw, sh, sw);
    CK(launch_wgrad_simt(g, x, Cin, 0, Cin, dy, Cout, 0, Cout, dw, (long long)Cin * Cout, Cout, 1, st));
    if (dbias) CK(launch_colsum(dy, (long long)g.B * g.Hy * g.Wx, Cout, 0, Cout, dbias, st));
  }
  return 0;
}

int cgvc_in_glu_forward(cgvc_handle e, const float* p, const float* beta_a, const float* gamma_a, const float* beta_g, const float* gamma_g,
                        float* y, float* stats, int B, int R, int C, int shuffle, void* stream) {
  if (!e || !p || !y || !stats) return fail(e, CGVC_ERR_ARG, "null argument");
  if (C % 32 != 0 || shuffle < 1 || R % shuffle != 0) return fail(e, CGVC_ERR_UNSUPPORTED, "C must be a multiple of 32 and R of shuffle");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  PostParams q; memset(&q, 0, sizeof q);
  q.p = p; q.ldp = 2 * C * shuffle; q.Cc = C * shuffle; q.B = B; q.R = R; q.C = C; q.sh = shuffle;
  q.beta_a = beta_a; q.gamma_a = gamma_a; q.beta_g = beta_g; q.gamma_g = gamma_g; q.has_in = 1; q.has_gate = 1; q.y = y; q.stats = stats;
  CK(launch_post_fwd(q, (cudaStream_t)stream));
  return 0;
}

int cgvc_in_glu_backward(cgvc_handle e, const float* dy, const float* p, const float* stats,
                         const float* beta_a, const float* gamma_a, const float* beta_g, const float* gamma_g,
                         float* dp, float* dbeta_a, float* dgamma_a, float* dbeta_g, float* dgamma_g,
                         int B, int R, int C, int shuffle, void* stream) {
  if (!e || !dy || !p || !stats || !dp) return fail(e, CGVC_ERR_ARG, "null argument");
  if (C % 32 != 0 || shuffle < 1 || R % shuffle != 0) return fail(e, CGVC_ERR_UNSUPPORTED, "C must be a multiple of 32 and R of shuffle");
  DeviceGuard dguard; CK(dguard.set(e->cfg.device));
  PostBwdParams q; memset(&q, 0, sizeof q);
  q.dy1 = dy; q.p = p; q.ldp = 2 * C * shuffle; q.Cc = C * shuffle; q.B = B; q.R = R; q.C = C; q.sh = shuffle;
  q.beta_a = beta_a; q.gamma_a = gamma_a; q.beta_g = beta_g; q.gamma_g = gamma_g; q.has_in = 1; q.has_gate = 1; q.stats = stats;
  q.dp = dp; q.dbeta_a = dbeta_a; q.dgamma_a = dgamma_a; q.dbeta_g = dbeta_g; q.dgamma_g = dgamma_g;
  CK(launch_post_bwd(q, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
