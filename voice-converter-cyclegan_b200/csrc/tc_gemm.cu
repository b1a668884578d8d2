// tcgen05 gather-GEMM kernels of libcgvc.so (sm_100a): the dense contractions of the CycleGAN-VC hot path
// (module.py:22-64 convolutions, forward / data-gradient / weight-gradient) on the 5th-generation tensor cores.
//
// Arithmetic: every fp32 operand x is kept in HBM as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)); a product
// is evaluated as hi*hi + hi*lo + lo*hi by three tcgen05.mma (kind::f16, bf16 inputs) accumulating in fp32 in
// TMEM ("bf16x3", ~2^-16 relative error per product); CGVC_PREC_BF16 issues the first MMA only.  CGVC_PREC_F16F8 (forward
// only) keeps fp16 + two scaled e4m3 planes instead and spends 2 MMA units per product: the two cross terms as kind::f8f6f4
// MMAs first, then the fp16 hi*hi MMAs whose first one rescales the accumulator by 2^-15 (scale-input-d).
//
// Kernel anatomy (both kernels are persistent, one CTA per SM, 288 threads, tiles / work items walked with stride gridDim.x):
//   warps 0-3  producers: gather the activation rows (im2col rows, zero-filled at the TF-SAME borders) with 16-byte cp.async
//              into SWIZZLE_128B shared memory; one thread TMA-loads the weight (NT) / gradient (TN) tile; both complete on
//              the stage's "full" mbarrier.  They run ahead across tile boundaries.
//   warp  4    allocates TMEM (2 accumulator stages), issues tcgen05.mma from one elected lane, frees stages with
//              tcgen05.commit
//   warps 5-8  epilogue: tcgen05.ld the finished accumulator stage while the next tile's MMAs run; bias / accumulate, the fused
//              instance-norm (+GLU / +residual) forward epilogues, the opt-in fused backward epilogues; every global store is
//              transposed through a per-warp shared-memory patch so that 8 lanes write one row (complete 128-byte lines)
#include "tc_gemm.cuh"
#include "geom.h"
#include "kernels.cuh"
#include "im2col_map.h"

#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <stdint.h>
#include <stdio.h>

namespace {

// ------------------------------------------------------------------------------------------------ TMA descriptors (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// bf16 tensor [d2][d1][d0] (d0 contiguous), box [1][b1][64], SWIZZLE_128B: lands as b1 rows of 128 bytes
bool make_tmap3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b1) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {64, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// same for any element type: tensor [d2][d1][d0] (d0 contiguous, esize bytes), box [1][b1][b0], SWIZZLE_128B (b0 * esize == 128)
bool make_tmap3_t(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int esize, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * (uint64_t)esize, d0 * d1 * (uint64_t)esize};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its prior cp.async have completed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA tile load (rank 3) into swizzled shared memory; completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load3(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f8f6f4 (e4m3 x e4m3 here): 128 x N x 32 per instruction, same shared-memory tile layout (32 bytes per k-step)
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = A * B + D * 2^-15  (scale-input-d): folds the common 2^15 of the fp8 cross products out of the accumulator
__device__ __forceinline__ void umma_f16_rescale(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 15;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ CTA-pair (cta_group::2) PTX
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// A protocol error in a pair kernel must not hang the device: waits give up after ~4 s and trap (the launch then fails loudly).
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (uint32_t tries = 0;; ++tries) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if ((tries & 0x3FFu) == 0x3FFu) {
      uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t; else if (t - t0 > 4000000000ull) __trap();
    }
  }
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot) {        // executed by the same warp of BOTH CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[both CTAs' smem, 128 rows each] * B[both CTAs' smem, N/2 rows each]; issued by the leader only
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = A * B + D * 2^-SHIFT (scale-input-d): folds the common scale of the fp8 cross products out of the accumulator
template <int SHIFT>
__device__ __forceinline__ void umma2_f16_rescale(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, %4;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "n"(SHIFT) : "memory");
}
// the barrier at this shared-memory offset receives one arrival in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// TMA loads of a pair: data lands in the issuing CTA's shared memory, the bytes are counted on the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma2_load3(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}
__device__ __forceinline__ void tma2_im2col(uint32_t dst_smem, const CUtensorMap* map, int c, int w, int h, int n, unsigned short ow, unsigned short oh, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(dst_smem), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
}
// one arrival on the barrier at this shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): SWIZZLE_128B, version 1.
//   K-major : 8-row groups of 128-byte rows, SBO = 1024 B between groups, LBO unused (1)
//   MN-major: atoms of (64 MN-elements x 8 K-rows) = 1024 B; LBO = stride between atoms along MN, SBO = along K
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp16 x fp16 -> f32 (kind::f16) and e4m3 x e4m3 -> f32 (kind::f8f6f4) share this encoding: formats 0 / 0, K-major
__host__ __device__ constexpr uint32_t make_idesc_f0(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// exact unsigned division by a runtime constant (n < 2^31): q = (umulhi(n, mul) + n) >> shr, Granlund-Montgomery round-up form
struct FastDiv { uint32_t mul, shr, d; };
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.d = d;
  uint32_t l = 0; while ((1ull << l) < d) ++l;
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.shr = l;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (uint32_t)(((uint64_t)__umulhi(n, f.mul) + n) >> f.shr); }

// byte offset of (row r, 16-byte chunk c) inside a [rows x 128 B] SWIZZLE_128B tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// ------------------------------------------------------------------------------------------------ parameters
struct TcNTParams {                 // forward / data-gradient form: D[m,n] = sum_t sum_c A[src(m,t), c] * B[t][n][c]
  GatherGeom g;
  const __nv_bfloat16 *a_hi, *a_lo; int a_ld; int C;     // gathered operand planes, row stride (elements), contraction channels
  const __nv_bfloat16 *b_hi, *b_lo; int Nw;              // weights [slab][Nw][C]
  int N;                                                 // real output columns (stores are guarded; tiles cover Nw)
  int n_tiles;                                           // column tiles to compute (host: covers the real columns only)
  int debug;                                             // diagnostic knobs (tc_set_debug): 1 = epilogue skips its global stores, 2 = also skips
                                                         // the TMEM loads, 4 = producers skip the A gather.  Results are garbage; timing only.
  float* dst; int d_ld; const float* bias; int accumulate;
  int perm; int Cc;                                      // gated layers: weight rows / bias are stored tile-interleaved: tile j =
                                                         // [a-channels j*128..+127 | g-channels j*128..+127]; Cc = channels per branch
  // fused forward epilogue (EPI 1: instance norm + GLU, EPI 2: instance norm + residual); rows = whole samples of R positions
  int R;                                                 // positions per sample (32, 64 or 128)
  const float *gamma_a, *beta_a, *gamma_g, *beta_g;
  float* stats;                                          // [B,4,C]: mean_a, rstd_a, mean_g, rstd_g
  const float* resid;                                    // [M,C] (EPI 2)
  float* y; __nv_bfloat16 *y_hi, *y_lo;                  // [M,C] outputs (y may be null)
  int C_out;                                             // channels of y (Cc for gated, N for EPI 2); channels per branch (EPI 3, 4)
  // fused backward epilogue of the data-gradient form (EPI 3: GLU + instance-norm backward of the gated layer whose output this
  // gradient is; EPI 4: instance-norm backward of a residual block's second convolution).  dY = accumulator (+ dst when accumulate).
  const float* bp; int bp_ld;                            // that layer's saved pre-norm conv outputs [M, bp_ld] (a columns, then gate columns)
  __nv_bfloat16 *dp_hi, *dp_lo; int dp_ld;               // its dP planes, written here [M, dp_ld]
  float *dbeta_a, *dgamma_a, *dbeta_g, *dgamma_g;        // instance-norm parameter gradients (atomically accumulated; may be null)
  CUtensorMap tm_b_hi, tm_b_lo;                          // TMA maps of the weight planes, box [1][BN][64]
  // F16F8 mode (NPL == 3, forward only): a_hi / tm_b_hi are the fp16 planes; the e4m3 planes of both operands:
  const uint8_t *a8_hi, *a8_lo;                          // [rows, a_ld] bytes
  CUtensorMap tm_b8_hi, tm_b8_lo;                        // box [1][BN][128 bytes]
  uint8_t* y8;                                           // fused epilogues: q8hi plane of y [M, C_out] bytes, q8lo follows at + M * C_out (y_hi = q16)
  // CTA-pair kernel (tc_pair_nt_kernel): the gathered operand comes through TMA im2col maps of the activation planes
  // (128 pixels x 64 channels per load), each CTA of a pair loads half of the weight tile (box [1][BN/2][64])
  Im2colGeom ig;
  CUtensorMap tm_a_hi, tm_a_lo;
  CUtensorMap tm_b2_hi, tm_b2_lo;
  CUtensorMap tm_b64_hi, tm_b64_lo; int have_b64;        // the weight planes with 64-row boxes: 128-wide pair tiles chosen at launch time
  // F16F8 on CTA pairs: tm_a_hi / tm_b2_hi are then the fp16 planes; the e4m3 planes (128 channels per 128-byte line):
  CUtensorMap tm_a8_hi, tm_a8_lo;                        // im2col, 128 pixels x 128 bytes
  CUtensorMap tm_b28_hi, tm_b28_lo;                      // box [1][BN/2][128 bytes]
};

struct TcTNParams {                 // weight-gradient form: D_t[c,n] = sum_m X[src(m,t), c] * G[m, n]
  GatherGeom g;
  const __nv_bfloat16 *x_hi, *x_lo; int x_ld; int C;     // gathered operand planes [.., x_ld] (x_ld = C rounded up to 64); C = real channels (GEMM M)
  const __nv_bfloat16 *g_hi, *g_lo; int g_ld; int N;     // dense gradient planes [M, g_ld] (g_ld = N rounded up to 64); N = real columns
  float* dw_a; float* dw_g; int n_split;                 // columns [0,n_split) -> dw_a[t][c][n], rest -> dw_g[t][c][n-n_split]
  int ksplit;
  FastDiv div_hw, div_w;                                 // m -> (b, y, x) without integer division
  CUtensorMap tm_g_hi, tm_g_lo;                          // TMA maps of the gradient planes [M][g_ld], box [64 rows][64 cols]
  // CTA-pair kernel (tc_pair_tn_kernel): X through TMA im2col maps (64 pixels x 64 channels per load)
  Im2colGeom ig;
  CUtensorMap tm_x_hi, tm_x_lo;
  // F16F8 weight gradient (tc_pair_tn_q_kernel), 128 K-rows per stage: fp16 planes (tm_xq: im2col 128 pixels x 64 channels, tm_gq: box
  // [128 rows][64 columns]) and e4m3 planes (tm_x8_*: im2col 128 pixels x 128 channels, tm_g8_*: box [128 rows][128 columns])
  CUtensorMap tm_xq, tm_gq, tm_x8_hi, tm_x8_lo, tm_g8_hi, tm_g8_lo;
  int w16;                                               // 1: fp16 planes only (one MMA unit per product; weight gradients are leaves of the graph)
  int fold_n;                                            // != 0: tap-folded layer (TcLayer::fold): column = t * fold_n + n of a [taps][C][fold_n] TF kernel
};

// address of weight-gradient element (tap slab, channel c, column col) in the TF-layout kernel [taps][C][ncols]; with fold_n the layer's
// real taps live in the column dimension (col = t * fold_n + n) and the kernel in memory is [taps][C][fold_n]
__device__ __forceinline__ float* tn_dst(float* base, int slab, int C, int c, int ncols, int col, int fold_n) {
  if (fold_n) { const int t = col / fold_n; return base + ((long long)t * C + c) * fold_n + (col - t * fold_n); }
  return base + ((long long)slab * C + c) * ncols + col;
}

constexpr int kProducerThreads = 128;


template <int BN, int NPL>
struct NTCfg {
  static constexpr int A_PLANE = 128 * 128;            // bytes: 128 rows x 128 B
  static constexpr int B_PLANE = BN * 128;
  static constexpr int PLANES = NPL == 1 ? 1 : 2;       // F16F8 (NPL = 3) stages hold two 128-byte-row tiles per operand as well
  static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
  static constexpr int STAGES = (200 * 1024) / STAGE;   // 2 (BN=256,x3), 3 (128,x3), 5 (32,x3), 4 (256,x1), 6 (128,x1), 10 (32,x1)
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

// ---- fused instance-norm epilogue helpers -----------------------------------------------------------------------
// Column sums over the 32 rows of a warp (row = lane): butterfly transpose-reduce, 31 shuffles for 32 columns.
// On return lane j holds the sum of column j (in t[0]).  t is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&t)[32], int lane) {
#pragma unroll
  for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      float send = up ? t[i] : t[i + n / 2];
      float keep = up ? t[i + n / 2] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return t[0];
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }      // the 4 epilogue warps of a group (barrier 1 + group)

// combine a per-warp column statistic over the warps that share a sample (spw = R/32 warps) through shared memory
__device__ __forceinline__ float sample_sum(float v, float (*xch)[32], int q, int lane, int spw, int barid) {
  if (spw == 1) return v;
  epi_bar(barid);
  xch[q][lane] = v;
  epi_bar(barid);
  const int q0 = (q / spw) * spw;
  float r = 0.f;
  for (int w = 0; w < spw; ++w) r += xch[q0 + w][lane];
  return r;
}
__device__ __forceinline__ void bc4(const float* p, float (&v)[4]) { const float4 x = *reinterpret_cast<const float4*>(p); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// Instance-norm statistics of one 32-column chunk held as v[32] per row-thread; writes per-column scale/offset
// (norm(x) = x*sc + of) for the 32 columns into bc[0..31] / bc[32..63] of this warp's broadcast area and returns the
// column's (mean, rstd) in lane == column.
__device__ __forceinline__ void chunk_norm_coeffs(const float (&v)[32], const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  int ch, int R, float (*xch)[32], float* bc, int q, int lane, int spw, int barid, float& mean, float& rstd) {
  float t[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) t[k] = v[k];
  float s = warp_colsum32(t, lane);
  s = sample_sum(s, xch, q, lane, spw, barid);
  mean = s / (float)R;
  __syncwarp();
  bc[lane] = mean;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    float4 m4 = *reinterpret_cast<const float4*>(bc + k);
    float d0 = v[k] - m4.x, d1 = v[k + 1] - m4.y, d2 = v[k + 2] - m4.z, d3 = v[k + 3] - m4.w;
    t[k] = d0 * d0; t[k + 1] = d1 * d1; t[k + 2] = d2 * d2; t[k + 3] = d3 * d3;
  }
  float ss = warp_colsum32(t, lane);
  ss = sample_sum(ss, xch, q, lane, spw, barid);
  rstd = 1.f / sqrtf(ss / (float)R + 1e-6f);                       // module.py:11 epsilon
  const float sc = rstd * gamma[ch + lane];
  const float of = beta[ch + lane] - mean * sc;
  __syncwarp();
  bc[lane] = sc; bc[32 + lane] = of;
  __syncwarp();
}

// The same for a pixel-shuffled layer (module.py:135-146): post-shuffle channel c holds the conv channels c (v0) and c + C/2 (v1) of
// every conv row, so its statistics run over 2R values.
__device__ __forceinline__ void pair_norm_coeffs(const float (&v0)[32], const float (&v1)[32], const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 int ch, int R, float (*xch)[32], float* bc, int q, int lane, int spw, int barid, float& mean, float& rstd) {
  float t[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) t[k] = v0[k] + v1[k];
  float s = warp_colsum32(t, lane);
  s = sample_sum(s, xch, q, lane, spw, barid);
  const float inv = 1.f / (float)(2 * R);
  mean = s * inv;
  __syncwarp();
  bc[lane] = mean;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    float4 m4 = *reinterpret_cast<const float4*>(bc + k);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float d0 = v0[k + j] - m[j], d1 = v1[k + j] - m[j]; t[k + j] = fmaf(d0, d0, d1 * d1); }
  }
  float ss = warp_colsum32(t, lane);
  ss = sample_sum(ss, xch, q, lane, spw, barid);
  rstd = 1.f / sqrtf(ss * inv + 1e-6f);                              // module.py:11 epsilon
  const float sc = rstd * gamma[ch + lane];
  const float of = beta[ch + lane] - mean * sc;
  __syncwarp();
  bc[lane] = sc; bc[32 + lane] = of;
  __syncwarp();
}

// ---- warp-cooperative coalesced row stores -----------------------------------------------------------------------
// Every lane of an epilogue warp owns one tile row (TMEM lane == row).  If each lane stored its own 32 columns, one warp-wide
// 16-byte store would touch 32 different rows: 32 half-used sectors and 32 address phases in the LSU, queued in front of the
// producers' cp.async (measured: 8 % of the kernel, profiles/r01_layer_profile_v9.txt, tc_debug = 1).  Instead the warp
// transposes each 32 x 32-word block through a 4 KB shared-memory patch (16-byte chunks XOR-swizzled by the row: conflict
// free both ways) and writes it back with 8 lanes per row -- 4 complete 128-byte lines per instruction.
__device__ __forceinline__ void stage_rows(float* stg, const float (&o)[32], int lane) {
  __syncwarp();                                            // earlier readers of the patch are done
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
  __syncwarp();
}
__device__ __forceinline__ float4 staged_chunk(const float* stg, int row, int chunk) {
  return *reinterpret_cast<const float4*>(stg + row * 32 + ((chunk ^ (row & 7)) << 2));
}
// ---- half-width transposition patch (NT epilogues) -------------------------------------------------------------------------
// The forward / data-gradient epilogues run on up to 8 warps per CTA (CTA-pair kernel), and shared memory is full at 3 x 64 KB of
// pipeline stages, so their per-warp patch is 32 rows x 16 words = 2 KB: a 32-word row passes through it in two halves.  16-byte
// chunk c (0..3) of row r lives at word r*16 + ((c ^ ((r >> 1) & 3)) << 2): conflict-free for the row-wise writes (lane = row) and
// for the write-back role (chunk lane & 3 of rows (lane >> 2) + 8 i), which covers 8 rows x 64 bytes per warp instruction.
__device__ __forceinline__ uint32_t hp_off(int r, int c) { return (uint32_t)(r * 16 + ((c ^ ((r >> 1) & 3)) << 2)); }
// this lane's 32 words (its tile row) -> f(rr, w0, val): val = words [w0, w0 + 4) of tile row rr; 8 calls per lane
template <class F>
__device__ __forceinline__ void rows_out(float* stg, const float (&o)[32], int lane, F f) {
  const int hc = lane & 3, hr = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __syncwarp();                                          // earlier readers of the patch are done
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(stg + hp_off(lane, c)) = make_float4(o[16 * h + 4 * c], o[16 * h + 4 * c + 1], o[16 * h + 4 * c + 2], o[16 * h + 4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = hr + 8 * i;
      f(rr, 16 * h + 4 * hc, *reinterpret_cast<const float4*>(stg + hp_off(rr, hc)));
    }
  }
}
// the inverse: pre[4 h + i] = words [16 h + 4 (lane & 3), + 4) of tile row (lane >> 2) + 8 i (fetched by the caller, 8 rows x 64 bytes
// per warp instruction) -> v = this lane's row
__device__ __forceinline__ void rows_in(float* stg, float (&v)[32], int lane, const float4 (&pre)[8]) {
  const int hc = lane & 3, hr = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg + hp_off(hr + 8 * i, hc)) = pre[4 * h + i];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 x = *reinterpret_cast<const float4*>(stg + hp_off(lane, c));
      v[16 * h + 4 * c] = x.x; v[16 * h + 4 * c + 1] = x.y; v[16 * h + 4 * c + 2] = x.z; v[16 * h + 4 * c + 3] = x.w;
    }
  }
}
// fetch for rows_in: 32 columns [col, col + 32) of the dense rows mq .. mq + 31 of a [M, ld] fp32 matrix (rows >= M read as zero)
__device__ __forceinline__ void rows_fetch(float4 (&pre)[8], const float* base, long long ld, long long mq, long long M, int col, int lane) {
  const int hc = lane & 3, hr = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = hr + 8 * i;
      pre[4 * h + i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mq + rr < M) pre[4 * h + i] = *reinterpret_cast<const float4*>(base + (mq + rr) * ld + col + 16 * h + 4 * hc);
    }
}
__device__ __forceinline__ void hload_rows(float* stg, float (&v)[32], const float* base, long long ld, long long mq, long long M, int col, int lane) {
  float4 pre[8];
  rows_fetch(pre, base, ld, mq, M, col, lane);
  rows_in(stg, v, lane, pre);
}
// the staged block to 32 arbitrary destination rows (null = row outside the tensor), columns [col0, col0 + 32)
__device__ __forceinline__ void hwrite_rows_f32(float* stg, const float (&o)[32], float* const* rowp, int col0, int lane) {
  rows_out(stg, o, lane, [&](int rr, int w0, float4 val) {
    float* rp = rowp[rr];
    if (rp != nullptr) *reinterpret_cast<float4*>(rp + col0 + w0) = val;
  });
}
// y[32] of this lane's row -> optional fp32 copy and the bf16 hi/lo planes of the dense [M, C] activation (rows mq .. mq + 31)
// rs / ro: destination row of tile row r is r * rs + ro (pixel shuffle: the two halves of a conv row are output rows 2r and 2r + 1)
__device__ __forceinline__ void hwrite_y(float* stg, const float (&y)[32], float* yf, __nv_bfloat16* yhi, __nv_bfloat16* ylo,
                                         long long mq, long long M, int C, int ch, int lane, int rs = 1, int ro = 0) {
  if (yf)
    rows_out(stg, y, lane, [&](int rr, int w0, float4 val) {
      if (mq + rr < M) *reinterpret_cast<float4*>(yf + ((mq + rr) * rs + ro) * C + ch + w0) = val;
    });
  // planes: the row's 32 words are [16 words of hi bf16 pairs | 16 words of lo bf16 pairs]; word w of a half = columns 2w, 2w + 1
  float hl[32];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(y[2 * k], y[2 * k + 1]);
    float2 f = __bfloat1622float2(hh);
    __nv_bfloat162 ll = __floats2bfloat162_rn(y[2 * k] - f.x, y[2 * k + 1] - f.y);
    hl[k] = __uint_as_float(*reinterpret_cast<uint32_t*>(&hh)); hl[16 + k] = __uint_as_float(*reinterpret_cast<uint32_t*>(&ll));
  }
  rows_out(stg, hl, lane, [&](int rr, int w0, float4 val) {
    __nv_bfloat16* plane = (w0 < 16) ? yhi : ylo;
    if (mq + rr < M) *reinterpret_cast<float4*>(plane + ((mq + rr) * rs + ro) * C + ch + 2 * (w0 & 15)) = val;
  });
}
// F16F8 variant: the row's 32 words are [16 words of fp16 pairs | 8 words of e4m3 hi quads | 8 words of e4m3 lo quads]
__device__ __forceinline__ void hwrite_yq(float* stg, const float (&y)[32], float* yf, __nv_bfloat16* q16, uint8_t* q8,
                                          long long mq, long long M, int C, int ch, int lane, int rs = 1, int ro = 0) {
  if (yf)
    rows_out(stg, y, lane, [&](int rr, int w0, float4 val) {
      if (mq + rr < M) *reinterpret_cast<float4*>(yf + ((mq + rr) * rs + ro) * C + ch + w0) = val;
    });
  float w[32];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v[4] = {y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]};
    uint2 h; uint32_t b_hi, b_lo;
    cgvc_quant4(v, CGVC_Q_ACT_SHI, CGVC_Q_ACT_SLO, h, b_hi, b_lo);
    w[2 * k] = __uint_as_float(h.x); w[2 * k + 1] = __uint_as_float(h.y);
    w[16 + k] = __uint_as_float(b_hi); w[24 + k] = __uint_as_float(b_lo);
  }
  rows_out(stg, w, lane, [&](int rr, int w0, float4 val) {
    if (mq + rr >= M) return;
    const long long r = (mq + rr) * rs + ro;
    uint8_t* dst;
    if (w0 < 16) dst = reinterpret_cast<uint8_t*>(q16) + (r * C + ch + 2 * w0) * 2;                  // 8 fp16 columns
    else if (w0 < 24) dst = q8 + r * C + ch + 4 * (w0 - 16);                                           // 16 e4m3 columns
    else dst = q8 + M * rs * C + r * C + ch + 4 * (w0 - 24);
    *reinterpret_cast<float4*>(dst) = val;
  });
}

// ---- epilogue of one 128-row x BN-column accumulator tile (shared by the one-CTA and the CTA-pair kernels) ----------------
// Called by the epilogue warps of a CTA: 4 warps (one per TMEM lane quarter q) form a group; with ngrp = 2 groups (CTA-pair kernel)
// group grp takes the 32-column chunks grp, grp + 2, ... of the tile.  m0 = first row of this CTA's 128 rows, n0 = first column of
// the tile, tacc = TMEM address of this warp's lanes of the accumulator stage; waits for `acc_full_bar` (parity aphase).
// stg / rowp / bc: this warp's 2 KB transposition patch, destination-row table and coefficient broadcast area; epi_xch: the group's
// cross-warp exchange area (statistics of samples that span several warps; barrier 1 + grp).
template <int BN, int NPL, int EPI>
__device__ __forceinline__ void nt_tile_epilogue(const TcNTParams& p, const long long M, const int HW, const long long m0, const int n0,
                                                 const int q, const int lane, float* stg, float** rowp, float* bc, float (*epi_xch)[32],
                                                 uint64_t* acc_full_bar, const uint32_t aphase, const uint32_t tacc,
                                                 const int grp = 0, const int ngrp = 1) {
  const GatherGeom& g = p.g;
  const int barid = 1 + grp;
  const long long mq = m0 + q * 32;                      // first row of this warp
  const long long m = mq + lane;                         // TMEM lane == tile row
  float* drow = nullptr;
  if (m < M && p.dst) {                                  // dst may be null for the fused forward epilogues (inference: nothing kept for backward)
    int b = (int)(m / HW); int rem = (int)(m - (long long)b * HW);
    int y = rem / g.Wx; int x = rem - y * g.Wx;
    long long dr = ((long long)(b * g.Hd + y * g.dsy + g.doy) * g.Wd + x * g.dsx + g.dox);
    drow = p.dst + dr * p.d_ld;
  }
  __syncwarp();
  rowp[lane] = drow;                                     // destination row of every tile row, for the write-back lanes
  __syncwarp();
  mbar_wait(acc_full_bar, aphase);
  tc_fence_after();
  if (EPI == 0) {
#pragma unroll 1
    for (int cb = grp; cb < BN / 32; cb += ngrp) {
      const int nb = n0 + cb * 32;                         // column in weight-row (bias) order
      if (nb >= p.Nw) break;
      // gated layers store their weight rows tile-interleaved ([128 a | 128 g] per 256-wide tile): map back
      // (shuffled gated layers: [64 a s0 | 64 a s1 | 64 g s0 | 64 g s1] for 64 post-shuffle channels, see EPI 5)
      const int n = p.perm == 2 ? ((cb >> 2) * p.Cc + ((cb >> 1) & 1) * (p.Cc >> 1) + (n0 >> 2) + (cb & 1) * 32)
                  : p.perm      ? ((cb < 4) ? (n0 >> 1) + cb * 32 : p.Cc + (n0 >> 1) + (cb - 4) * 32) : nb;
      if (n >= p.N) { if (p.perm) continue; else break; }  // warp-uniform
      if (p.debug & 2) continue;
      float o[32];
      { uint32_t v[32]; tmem_ld32(tacc + (uint32_t)(cb * 32), v); tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(v[k]); }
      if (p.bias) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) { float4 bb = *reinterpret_cast<const float4*>(p.bias + nb + k); o[k] += bb.x; o[k + 1] += bb.y; o[k + 2] += bb.z; o[k + 3] += bb.w; }
      }
      if (!(p.debug & 1))
        rows_out(stg, o, lane, [&](int rr, int w0, float4 val) {
          float* rp = rowp[rr];
          const int col = n + w0;                          // N is a multiple of 4; padded columns are never stored
          if (rp != nullptr && col < p.N) {
            if (p.accumulate)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rp + col), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            else
              *reinterpret_cast<float4*>(rp + col) = val;
          }
        });
    }
  } else {
    // ---- fused forward epilogue: the 128 rows of the tile are whole samples of R positions (R = 32, 64, 128) ----
    const int spw = p.R >> 5;                              // warps per sample
    const long long sample = mq / p.R;
    const bool stat_writer = (q % spw) == 0 && mq < M && p.stats != nullptr;
    if (EPI == 1) {
      // gated: tile = [128 a-channels | the same 128 g-channels]; y = IN(a) * sigmoid(IN(g))   (module.py:3-20,85-98)
      const int ch0 = n0 >> 1;
#pragma unroll 1
      for (int cb = grp; cb < 4; cb += ngrp) {
        const int ch = ch0 + cb * 32;
        float va[32], vg[32];
        { uint32_t u[32]; tmem_ld32(tacc + (uint32_t)(cb * 32), u); tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) { float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + cb * 32 + k);
            va[k] = __uint_as_float(u[k]) + bb.x; va[k + 1] = __uint_as_float(u[k + 1]) + bb.y; va[k + 2] = __uint_as_float(u[k + 2]) + bb.z; va[k + 3] = __uint_as_float(u[k + 3]) + bb.w; } }
        { uint32_t u[32]; tmem_ld32(tacc + (uint32_t)(128 + cb * 32), u); tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) { float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + 128 + cb * 32 + k);
            vg[k] = __uint_as_float(u[k]) + bb.x; vg[k + 1] = __uint_as_float(u[k + 1]) + bb.y; vg[k + 2] = __uint_as_float(u[k + 2]) + bb.z; vg[k + 3] = __uint_as_float(u[k + 3]) + bb.w; } }
        // pre-norm outputs are kept for the backward pass
        if (p.dst) {
          hwrite_rows_f32(stg, va, rowp, ch, lane);
          hwrite_rows_f32(stg, vg, rowp, p.Cc + ch, lane);
        }
        float mean_a, rstd_a, mean_g, rstd_g;
        chunk_norm_coeffs(va, p.gamma_a, p.beta_a, ch, p.R, epi_xch, bc, q, lane, spw, barid, mean_a, rstd_a);
        chunk_norm_coeffs(vg, p.gamma_g, p.beta_g, ch, p.R, epi_xch, bc + 64, q, lane, spw, barid, mean_g, rstd_g);
        if (stat_writer) {
          float* st = p.stats + sample * 4 * p.C_out + ch + lane;
          st[0] = mean_a; st[p.C_out] = rstd_a; st[2 * p.C_out] = mean_g; st[3 * p.C_out] = rstd_g;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {                  // y overwrites va
          float4 sa = *reinterpret_cast<const float4*>(bc + k), oa = *reinterpret_cast<const float4*>(bc + 32 + k);
          float4 sg = *reinterpret_cast<const float4*>(bc + 64 + k), og = *reinterpret_cast<const float4*>(bc + 96 + k);
          va[k]     = fmaf(va[k],     sa.x, oa.x) * fast_sigmoid(fmaf(vg[k],     sg.x, og.x));
          va[k + 1] = fmaf(va[k + 1], sa.y, oa.y) * fast_sigmoid(fmaf(vg[k + 1], sg.y, og.y));
          va[k + 2] = fmaf(va[k + 2], sa.z, oa.z) * fast_sigmoid(fmaf(vg[k + 2], sg.z, og.z));
          va[k + 3] = fmaf(va[k + 3], sa.w, oa.w) * fast_sigmoid(fmaf(vg[k + 3], sg.w, og.w));
        }
        if (NPL == 3) hwrite_yq(stg, va, p.y, p.y_hi, p.y8, mq, M, p.C_out, ch, lane);
        else hwrite_y(stg, va, p.y, p.y_hi, p.y_lo, mq, M, p.C_out, ch, lane);
      }
    } else if (EPI == 5) {
      // gated + pixel shuffle (upsample1d_block, module.py:115-146): the tile holds, for 64 post-shuffle channels c, the conv columns
      // [a(c) | a(c + Ch) | g(c) | g(c + Ch)], 64 each (Ch = Cc / 2 post-shuffle channels).  Conv row r of a sample becomes output rows
      // 2r (columns c) and 2r + 1 (columns c + Ch); the statistics of channel c run over both: 2R positions.
      const int Ch = p.C_out;
      const int ch0 = n0 >> 2;
#pragma unroll 1
      for (int cb = grp; cb < 2; cb += ngrp) {
        const int ch = ch0 + cb * 32;
        auto load = [&](float (&v)[32], int tcol) {
          uint32_t u[32]; tmem_ld32(tacc + (uint32_t)tcol, u); tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) { float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + tcol + k);
            v[k] = __uint_as_float(u[k]) + bb.x; v[k + 1] = __uint_as_float(u[k + 1]) + bb.y; v[k + 2] = __uint_as_float(u[k + 2]) + bb.z; v[k + 3] = __uint_as_float(u[k + 3]) + bb.w; }
        };
        float mean_a, rstd_a, mean_g, rstd_g;
        {                                                  // gate branch first: statistics only, the values are re-read from TMEM below
          float g0[32], g1[32];
          load(g0, 128 + cb * 32); load(g1, 192 + cb * 32);
          if (p.dst) { hwrite_rows_f32(stg, g0, rowp, p.Cc + ch, lane); hwrite_rows_f32(stg, g1, rowp, p.Cc + Ch + ch, lane); }
          pair_norm_coeffs(g0, g1, p.gamma_g, p.beta_g, ch, p.R, epi_xch, bc + 64, q, lane, spw, barid, mean_g, rstd_g);
        }
        float a0[32], a1[32];
        load(a0, cb * 32); load(a1, 64 + cb * 32);
        if (p.dst) { hwrite_rows_f32(stg, a0, rowp, ch, lane); hwrite_rows_f32(stg, a1, rowp, Ch + ch, lane); }
        pair_norm_coeffs(a0, a1, p.gamma_a, p.beta_a, ch, p.R, epi_xch, bc, q, lane, spw, barid, mean_a, rstd_a);
        if (stat_writer) {
          float* st = p.stats + sample * 4 * Ch + ch + lane;
          st[0] = mean_a; st[Ch] = rstd_a; st[2 * Ch] = mean_g; st[3 * Ch] = rstd_g;
        }
        auto emit = [&](const float (&a)[32], const int s) {
          float gg[32];
          load(gg, 128 + s * 64 + cb * 32);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float4 sa = *reinterpret_cast<const float4*>(bc + k), oa = *reinterpret_cast<const float4*>(bc + 32 + k);
            float4 sg = *reinterpret_cast<const float4*>(bc + 64 + k), og = *reinterpret_cast<const float4*>(bc + 96 + k);
            gg[k]     = fmaf(a[k],     sa.x, oa.x) * fast_sigmoid(fmaf(gg[k],     sg.x, og.x));
            gg[k + 1] = fmaf(a[k + 1], sa.y, oa.y) * fast_sigmoid(fmaf(gg[k + 1], sg.y, og.y));
            gg[k + 2] = fmaf(a[k + 2], sa.z, oa.z) * fast_sigmoid(fmaf(gg[k + 2], sg.z, og.z));
            gg[k + 3] = fmaf(a[k + 3], sa.w, oa.w) * fast_sigmoid(fmaf(gg[k + 3], sg.w, og.w));
          }
          if (NPL == 3) hwrite_yq(stg, gg, p.y, p.y_hi, p.y8, mq, M, Ch, ch, lane, 2, s);
          else hwrite_y(stg, gg, p.y, p.y_hi, p.y_lo, mq, M, Ch, ch, lane, 2, s);
        };
        emit(a1, 1);
        emit(a0, 0);
      }
    } else if (EPI == 2) {
      // EPI 2: y = resid + IN(conv)   (residual1d_block second half, module.py:79-83); 256 independent channels per tile
#pragma unroll 1
      for (int cb = grp; cb < BN / 32; cb += ngrp) {
        const int ch = n0 + cb * 32;
        if (ch >= p.N) break;
        // the residual input (read 4 lanes per row, 8 rows per instruction) is fetched first: its global-memory latency hides behind the
        // accumulator load and the statistics of this chunk
        float4 rpre[8];
        rows_fetch(rpre, p.resid, p.C_out, mq, M, ch, lane);
        float va[32];
        { uint32_t u[32]; tmem_ld32(tacc + (uint32_t)(cb * 32), u); tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) { float4 bb = *reinterpret_cast<const float4*>(p.bias + ch + k);
            va[k] = __uint_as_float(u[k]) + bb.x; va[k + 1] = __uint_as_float(u[k + 1]) + bb.y; va[k + 2] = __uint_as_float(u[k + 2]) + bb.z; va[k + 3] = __uint_as_float(u[k + 3]) + bb.w; } }
        if (p.dst) hwrite_rows_f32(stg, va, rowp, ch, lane);
        float mean_a, rstd_a;
        chunk_norm_coeffs(va, p.gamma_a, p.beta_a, ch, p.R, epi_xch, bc, q, lane, spw, barid, mean_a, rstd_a);
        if (stat_writer) {
          float* st = p.stats + sample * 4 * p.C_out + ch + lane;
          st[0] = mean_a; st[p.C_out] = rstd_a; st[2 * p.C_out] = 0.f; st[3 * p.C_out] = 1.f;
        }
        // transpose the residual through the patch: every lane then holds its own row
        float res[32];
        rows_in(stg, res, lane, rpre);
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          float4 scl = *reinterpret_cast<const float4*>(bc + k), of = *reinterpret_cast<const float4*>(bc + 32 + k);
          va[k] = fmaf(va[k], scl.x, of.x) + res[k]; va[k + 1] = fmaf(va[k + 1], scl.y, of.y) + res[k + 1];
          va[k + 2] = fmaf(va[k + 2], scl.z, of.z) + res[k + 2]; va[k + 3] = fmaf(va[k + 3], scl.w, of.w) + res[k + 3];
        }
        if (NPL == 3) hwrite_yq(stg, va, p.y, p.y_hi, p.y8, mq, M, p.C_out, ch, lane);
        else hwrite_y(stg, va, p.y, p.y_hi, p.y_lo, mq, M, p.C_out, ch, lane);
      }
    } else {
      // ---- EPI 3 / 4: fused backward.  The tile holds dY for 256 output channels of whole samples.
      //   EPI 3 (gated layer):  y = na * sigmoid(ng), na = IN(a), ng = IN(g):  dna = dY * s, dng = dna * na * (1 - s)
      //   EPI 4 (residual h2):  y = resid + IN(a):                             dna = dY (also written back: it is the skip gradient)
      //   IN backward per (sample, channel):  dx = sc * (dn - mean_R(dn) - xhat * mean_R(dn * xhat)),  sc = gamma * rstd
      const int C = p.C_out;
      const float invR = 1.f / (float)p.R;
      const bool live = mq < M;
#pragma unroll 1
      for (int cb = grp; cb < BN / 32; cb += ngrp) {
        const int ch = n0 + cb * 32;
        if (ch >= p.N) break;
        float dy[32], xa[32];
        { uint32_t u[32]; tmem_ld32(tacc + (uint32_t)(cb * 32), u); tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) dy[k] = __uint_as_float(u[k]); }
        if (p.accumulate) {
          hload_rows(stg, xa, p.dst, p.d_ld, mq, M, ch, lane);
#pragma unroll
          for (int k = 0; k < 32; ++k) dy[k] += xa[k];
        }
        if (EPI == 4) {                                    // gradient w.r.t. the block output: the next block's skip gradient
          rows_out(stg, dy, lane, [&](int rr, int w0, float4 val) {
            if (mq + rr < M) *reinterpret_cast<float4*>(p.dst + (mq + rr) * p.d_ld + ch + w0) = val;
          });
        }
        // per-column coefficients of this warp's sample (lane == column): xhat = x * r + h ; norm = x * sc + of
        {
          const float* st = p.stats + sample * 4 * C + ch + lane;
          const float mean = live ? st[0] : 0.f, rstd = live ? st[C] : 1.f;
          const float scl = rstd * p.gamma_a[ch + lane];
          __syncwarp();
          bc[lane] = rstd; bc[32 + lane] = -mean * rstd; bc[64 + lane] = scl;
          if (EPI == 3) {
            bc[96 + lane] = p.beta_a[ch + lane] - mean * scl;
            const float mg = live ? st[2 * C] : 0.f, rg = live ? st[3 * C] : 1.f;
            const float sg = rg * p.gamma_g[ch + lane];
            bc[128 + lane] = rg; bc[160 + lane] = -mg * rg; bc[192 + lane] = sg; bc[224 + lane] = p.beta_g[ch + lane] - mg * sg;
          }
          __syncwarp();
        }
        hload_rows(stg, xa, p.bp, p.bp_ld, mq, M, ch, lane);
        float dg[32], xg[32];
        if (EPI == 3) {
          hload_rows(stg, xg, p.bp, p.bp_ld, mq, M, C + ch, lane);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float ra[4], ha[4], sa[4], oa[4], rg[4], hg[4], sg[4], og[4];
            bc4(bc + k, ra); bc4(bc + 32 + k, ha); bc4(bc + 64 + k, sa); bc4(bc + 96 + k, oa);
            bc4(bc + 128 + k, rg); bc4(bc + 160 + k, hg); bc4(bc + 192 + k, sg); bc4(bc + 224 + k, og);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float na = fmaf(xa[k + j], sa[j], oa[j]), ng = fmaf(xg[k + j], sg[j], og[j]);
              const float sgm = fast_sigmoid(ng);
              const float dna = dy[k + j] * sgm;
              dy[k + j] = dna; dg[k + j] = dna * na * (1.f - sgm);
              xa[k + j] = fmaf(xa[k + j], ra[j], ha[j]);      // xhat_a
              xg[k + j] = fmaf(xg[k + j], rg[j], hg[j]);      // xhat_g
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float ra[4], ha[4];
            bc4(bc + k, ra); bc4(bc + 32 + k, ha);
#pragma unroll
            for (int j = 0; j < 4; ++j) xa[k + j] = fmaf(xa[k + j], ra[j], ha[j]);
          }
        }
        // column sums over this warp's 32 rows (lane == column), parameter gradients, then over the whole sample
        float t[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = dy[k];
        float s1a = warp_colsum32(t, lane);
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = dy[k] * xa[k];
        float s2a = warp_colsum32(t, lane);
        float s1g = 0.f, s2g = 0.f;
        if (EPI == 3) {
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = dg[k];
          s1g = warp_colsum32(t, lane);
#pragma unroll
          for (int k = 0; k < 32; ++k) t[k] = dg[k] * xg[k];
          s2g = warp_colsum32(t, lane);
        }
        if (p.dbeta_a && live) {
          atomicAdd(p.dbeta_a + ch + lane, s1a); atomicAdd(p.dgamma_a + ch + lane, s2a);
          if (EPI == 3) { atomicAdd(p.dbeta_g + ch + lane, s1g); atomicAdd(p.dgamma_g + ch + lane, s2g); }
        }
        s1a = sample_sum(s1a, epi_xch, q, lane, spw, barid); s2a = sample_sum(s2a, epi_xch, q, lane, spw, barid);
        if (EPI == 3) { s1g = sample_sum(s1g, epi_xch, q, lane, spw, barid); s2g = sample_sum(s2g, epi_xch, q, lane, spw, barid); }
        {
          const float sca = bc[64 + lane], scg = EPI == 3 ? bc[192 + lane] : 0.f;
          __syncwarp();
          bc[256 + lane] = sca * s1a * invR; bc[288 + lane] = sca * s2a * invR;
          if (EPI == 3) { bc[320 + lane] = scg * s1g * invR; bc[352 + lane] = scg * s2g * invR; }
          __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          float c1[4], c2[4], c3[4];
          bc4(bc + 64 + k, c1); bc4(bc + 256 + k, c2); bc4(bc + 288 + k, c3);
#pragma unroll
          for (int j = 0; j < 4; ++j) dy[k + j] = fmaf(c1[j], dy[k + j], -fmaf(xa[k + j], c3[j], c2[j]));
        }
        hwrite_y(stg, dy, nullptr, p.dp_hi, p.dp_lo, mq, M, p.dp_ld, ch, lane);
        if (EPI == 3) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            float c1[4], c2[4], c3[4];
            bc4(bc + 192 + k, c1); bc4(bc + 320 + k, c2); bc4(bc + 352 + k, c3);
#pragma unroll
            for (int j = 0; j < 4; ++j) dg[k + j] = fmaf(c1[j], dg[k + j], -fmaf(xg[k + j], c3[j], c2[j]));
          }
          hwrite_y(stg, dg, nullptr, p.dp_hi, p.dp_lo, mq, M, p.dp_ld, C + ch, lane);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ NT kernel
// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ... (m fastest, so CTAs that run
// concurrently share the weight tile in L2).  288 threads:
//   warps 0-3  producers (A rows by cp.async, weight tile by TMA), running ahead across tile boundaries
//   warp  4    TMEM allocation + single-thread tcgen05.mma issue; accumulators are double-buffered in TMEM
//              (2 x BN columns) so the MMAs of tile i+1 overlap the epilogue of tile i
//   warps 5-8  epilogue: tcgen05.ld -> +bias / accumulate -> fp32 stores, then release the accumulator stage
constexpr int kNTThreads = 288;

template <int BN, int NPL, int EPI>
__global__ void __launch_bounds__(kNTThreads, 1)
tc_gg_nt_kernel(const __grid_constant__ TcNTParams p) {
  using Cfg = NTCfg<BN, NPL>;
  __shared__ float epi_xch[4][32];                           // cross-warp exchange of the fused epilogue
  __shared__ __align__(16) float epi_bc[4][(EPI == 3 || EPI == 4) ? 384 : 128];   // per-warp broadcast of per-column coefficients (32 floats per quantity)
  __shared__ __align__(16) float epi_stage[4][32 * 16];      // per-warp half-width transposition patch of the coalesced row stores
  __shared__ float* epi_rowp[4][32];                         // destination row of every tile row
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatherGeom& g = p.g;
  const long long M = (long long)g.B * g.Hy * g.Wx;
  const int HW = g.Hy * g.Wx;
  // K blocks ("stages"): 64 channels each; F16F8 walks the contraction twice with 128 channels per stage -- first the two
  // fp8 cross products (planes a8_hi x b8_lo and a8_lo x b8_hi), then the fp16 hi x hi product whose first MMA rescales D
  const int cchunks = NPL == 3 ? (p.C >> 7) : (p.C >> 6);
  const int kb_pass = g.ntaps * cchunks;
  const int num_kb = NPL == 3 ? 2 * kb_pass : kb_pass;     // > 0 (the host never launches an empty contraction)
  const int m_tiles = (int)((M + 127) / 128);
  const int n_tiles = p.n_tiles;
  const int num_tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    // full: 128 cp.async arrivals (A rows) + 1 arrive.expect_tx whose bytes the weight-tile TMA completes
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducerThreads + 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_b_hi);
    if (NPL == 2) tma_prefetch_desc(&p.tm_b_lo);
    if (NPL == 3) { tma_prefetch_desc(&p.tm_b8_hi); tma_prefetch_desc(&p.tm_b8_lo); }
  }
  if (warp == 4) tmem_alloc<2 * BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4) {
    // ===================== producers =====================
    const int t = threadIdx.x;
    const int chunk = t & 7, rsub = t >> 3;                 // 8 threads cover one 128-byte row; 16 rows per pass
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long m0 = (long long)(tile % m_tiles) * 128;
      const int n0 = (tile / m_tiles) * BN;
      int rb[8], ry[8], rx[8];                              // decoded output coordinates of this thread's 8 A rows
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        long long m = m0 + rsub + 16 * i;
        if (m < M) {
          int b = (int)(m / HW); int rem = (int)(m - (long long)b * HW);
          int y = rem / g.Wx; int x = rem - y * g.Wx;
          rb[i] = b; ry[i] = y * g.sy; rx[i] = x * g.sx;
        } else { rb[i] = -1; ry[i] = 0; rx[i] = 0; }
      }
      for (int pass = 0; pass < (NPL == 3 ? 2 : 1); ++pass)
      for (int tap = 0; tap < g.ntaps; ++tap) {
        long long aoff[8];                                   // element offset of the source row, or -1
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int yy = ry[i] + g.oy[tap], xx = rx[i] + g.ox[tap];
          bool ok = rb[i] >= 0 && yy >= 0 && yy < g.Hs && xx >= 0 && xx < g.Ws;
          aoff[i] = ok ? ((long long)(rb[i] * g.Hs + yy) * g.Ws + xx) * p.a_ld + chunk * (NPL == 3 && pass == 0 ? 16 : 8) : -1;
        }
        for (int cc = 0; cc < cchunks; ++cc) {
          const int c0 = NPL == 3 ? (cc << 7) : (cc << 6);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sA = smem_base + stage * Cfg::STAGE;
          const uint32_t sB = sA + Cfg::PLANES * Cfg::A_PLANE;
          if (t == 0) {                                      // weight tile [BN rows n][128 bytes of c] by TMA (hardware 128B swizzle)
            mbar_expect_tx(&full_bar[stage], Cfg::PLANES * Cfg::B_PLANE);
            if (NPL == 3) {
              if (pass == 0) {
                tma_load3(sB, &p.tm_b8_hi, c0, n0, g.widx[tap], &full_bar[stage]);
                tma_load3(sB + Cfg::B_PLANE, &p.tm_b8_lo, c0, n0, g.widx[tap], &full_bar[stage]);
              } else {                                       // fp16: two 64-channel tiles
                tma_load3(sB, &p.tm_b_hi, c0, n0, g.widx[tap], &full_bar[stage]);
                tma_load3(sB + Cfg::B_PLANE, &p.tm_b_hi, c0 + 64, n0, g.widx[tap], &full_bar[stage]);
              }
            } else {
              tma_load3(sB, &p.tm_b_hi, c0, n0, g.widx[tap], &full_bar[stage]);
              if (NPL == 2) tma_load3(sB + Cfg::B_PLANE, &p.tm_b_lo, c0, n0, g.widx[tap], &full_bar[stage]);
            }
          }
          if (!(p.debug & 4)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = rsub + 16 * i;
              const uint32_t so = sw128(r, chunk);
              const bool ok = aoff[i] >= 0;
              const long long off = ok ? aoff[i] + c0 : 0;
              if (NPL == 3) {
                if (pass == 0) {                             // 128 e4m3 bytes per row and plane
                  cp_async16(sA + so, p.a8_hi + off, ok ? 16u : 0u);
                  cp_async16(sA + Cfg::A_PLANE + so, p.a8_lo + off, ok ? 16u : 0u);
                } else {                                     // 2 x 64 fp16 per row
                  cp_async16(sA + so, p.a_hi + off, ok ? 16u : 0u);
                  cp_async16(sA + Cfg::A_PLANE + so, p.a_hi + off + 64, ok ? 16u : 0u);
                }
              } else {
                cp_async16(sA + so, p.a_hi + off, ok ? 16u : 0u);
                if (NPL == 2) cp_async16(sA + Cfg::A_PLANE + so, p.a_lo + off, ok ? 16u : 0u);
              }
            }
          }
          cp_async_arrive_noinc(&full_bar[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);            // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        fence_proxy_async();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_base + stage * Cfg::STAGE;
          const uint32_t sB = sA + Cfg::PLANES * Cfg::A_PLANE;
          if (NPL == 3) {
            constexpr uint32_t idq = make_idesc_f0(128, BN);
            if (kb < kb_pass) {                              // fp8 cross products, K = 32 per instruction: a8_hi x b8_lo + a8_lo x b8_hi
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f8(tmem_d, make_desc(sA + k * 32, 16, 1024), make_desc(sB + Cfg::B_PLANE + k * 32, 16, 1024), idq, (kb | k) != 0);
                umma_f8(tmem_d, make_desc(sA + Cfg::A_PLANE + k * 32, 16, 1024), make_desc(sB + k * 32, 16, 1024), idq, 1);
              }
            } else {                                         // fp16 hi x hi, two 64-channel tiles of K = 16 steps
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t a = make_desc(sA + h * Cfg::A_PLANE + k * 32, 16, 1024), b = make_desc(sB + h * Cfg::B_PLANE + k * 32, 16, 1024);
                  if (kb == kb_pass && h == 0 && k == 0) umma_f16_rescale(tmem_d, a, b, idq);     // D = A*B + D * 2^-15
                  else umma_bf16(tmem_d, a, b, idq, 1);
                }
            }
          } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {                      // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
            const uint64_t a_hi = make_desc(sA + k * 32, 16, 1024);
            const uint64_t b_hi = make_desc(sB + k * 32, 16, 1024);
            umma_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
            if (NPL == 2) {
              const uint64_t a_lo = make_desc(sA + Cfg::A_PLANE + k * 32, 16, 1024);
              const uint64_t b_lo = make_desc(sB + Cfg::B_PLANE + k * 32, 16, 1024);
              umma_bf16(tmem_d, a_hi, b_lo, idesc, 1);
              umma_bf16(tmem_d, a_lo, b_hi, idesc, 1);
            }
          }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 5..8) =====================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    float* stg = epi_stage[q];                               // this warp's 32 x 32-word transposition patch
    float** rowp = epi_rowp[q];
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const long long m0 = (long long)(tile % m_tiles) * 128;
      const int n0 = (tile / m_tiles) * BN;
      nt_tile_epilogue<BN, NPL, EPI>(p, M, HW, m0, n0, q, lane, stg, rowp, epi_bc[q], epi_xch, &tmem_full_bar[as], aphase,
                                     tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair NT kernel
// The same contraction on CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair.  CTA r of a pair owns rows
// m0 + 128 r .. + 127 -- its own gathered-operand tile and its own 128 accumulator lanes, so every epilogue is per CTA and identical
// to the one-CTA kernel's -- and loads only rows n0 + r BN/2 .. of the weight tile; the leader issues M = 256 MMAs that read both
// CTAs' shared memory.  Each SM therefore pulls (128 + BN/2) instead of (128 + BN) operand rows per K-block from L2: a third fewer
// bytes at BN = 256, which is what bounds the one-CTA kernel.  The gathered operand is no longer gathered by
// threads: one TMA im2col load per (tap, 64-channel block, plane) delivers the 128 rows, zero-filled at the TF-SAME borders and
// across sample boundaries (im2col_map.h), so a CTA is: producer warp (one lane), MMA issuer warp, and -- the thread budget the
// gather used to take -- 8 epilogue warps in two groups of 4 (one warp per TMEM lane quarter and group; group g takes the 32-column
// chunks g, g + 2, ... of a tile): with one warp per scheduler the long dependent chains of the fused instance-norm epilogues had
// no other warp to hide their latency behind.  The fused-backward epilogues (EPI 3, 4) keep one group (their coefficient tables
// would not fit twice into what 3 x 64 KB of pipeline stages leave of the shared memory).
constexpr int kPairThreads = 192;       // weight-gradient pair kernels: producer, MMA issuer, 4 epilogue warps
constexpr int kPairNTThreads = 320;     // forward / data-gradient pair kernel: producer, MMA issuer, 8 epilogue warps

template <int BN, int NPL>
struct PairCfg {
  static constexpr int A_PLANE = 128 * 128;
  static constexpr int B_PLANE = (BN / 2) * 128;
  static constexpr int PLANES = NPL == 1 ? 1 : 2;          // F16F8 (NPL = 3) stages hold two 128-byte-row tiles per operand as well
  static constexpr int STAGE = PLANES * (A_PLANE + B_PLANE);
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE;  // 3 (BN=256,x3), 4 (128,x3), 6 (256,x1), 8 (128,x1)
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

template <int BN, int NPL, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairNTThreads, 1)
tc_pair_nt_kernel(const __grid_constant__ TcNTParams p) {
  using Cfg = PairCfg<BN, NPL>;
  constexpr int S = Cfg::STAGES;
  constexpr bool kBwdEpi = EPI == 3 || EPI == 4;             // the fused backward epilogues need 1.5 KB of coefficients per warp: one group
  constexpr int NG = kBwdEpi ? 1 : 2;                        // epilogue warp groups
  __shared__ float epi_xch[NG][4][32];
  __shared__ __align__(16) float epi_bc[4 * NG][kBwdEpi ? 384 : 128];
  __shared__ __align__(16) float epi_stage[4 * NG][32 * 16];
  __shared__ float* epi_rowp[4 * NG][32];
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const GatherGeom& g = p.g;
  const long long M = (long long)g.B * g.Hy * g.Wx;
  const int HW = g.Hy * g.Wx;
  // K blocks: 64 channels each; F16F8 (NPL = 3) walks the contraction twice with 128 channels per stage -- first the two e4m3 cross
  // products, then the fp16 hi x hi product whose first MMA rescales the accumulator (same scheme as the one-CTA kernel)
  const int cchunks = NPL == 3 ? (p.C >> 7) : (p.C >> 6);
  const int kb_pass = g.ntaps * cchunks;
  const int num_kb = NPL == 3 ? 2 * kb_pass : kb_pass;
  const int m_tiles = (int)((M + 127) / 128);
  const int m_pairs = (m_tiles + 1) >> 1;
  const int num_tiles = m_pairs * p.n_tiles;

  if (threadIdx.x == 0) {
    // full (leader's is used): one arrive.expect_tx by the leader's producer, completed by the TMA bytes of BOTH CTAs;
    // empty / tmem_full: one multicast commit; tmem_empty (leader's is used): the 4 * NG epilogue warps of both CTAs
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8 * NG); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_a_hi); tma_prefetch_desc(&p.tm_b2_hi);
    if (NPL == 2) { tma_prefetch_desc(&p.tm_a_lo); tma_prefetch_desc(&p.tm_b2_lo); }
    if (NPL == 3) { tma_prefetch_desc(&p.tm_a8_hi); tma_prefetch_desc(&p.tm_a8_lo); tma_prefetch_desc(&p.tm_b28_hi); tma_prefetch_desc(&p.tm_b28_lo); }
  }
  if (warp == 1) tmem_alloc2<2 * BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // both CTAs' barriers exist before anyone signals the peer
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== producer (one lane) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const long long m0 = ((long long)(tile % m_pairs) * 2 + rank) * 128;
        const int n0 = (tile / m_pairs) * BN + (int)rank * (BN / 2);
        // base pixel of the tile's first row (rows beyond the tensor: sample index >= B, the unit fills zeros)
        const int b = (int)(m0 / HW); const int rem = (int)(m0 - (long long)b * HW);
        const int y = rem / g.Wx, x = rem - y * g.Wx;
        const int cw = p.ig.lo_w + x * g.sx, ch = p.ig.lo_h + y * g.sy;
        for (int pass = 0; pass < (NPL == 3 ? 2 : 1); ++pass)
        for (int tap = 0; tap < g.ntaps; ++tap) {
          const unsigned short ow = p.ig.off_w[tap], oh = p.ig.off_h[tap];
          const int wslab = g.widx[tap];
          for (int cc = 0; cc < cchunks; ++cc) {
            const int c0 = NPL == 3 ? (cc << 7) : (cc << 6);
            mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
            const uint32_t sA = smem_base + stage * Cfg::STAGE;
            const uint32_t sB = sA + Cfg::PLANES * Cfg::A_PLANE;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * Cfg::STAGE);
            if (NPL == 3) {
              if (pass == 0) {                               // 128 e4m3 channels per line and plane
                tma2_im2col(sA, &p.tm_a8_hi, c0, cw, ch, b, ow, oh, &full_bar[stage]);
                tma2_im2col(sA + Cfg::A_PLANE, &p.tm_a8_lo, c0, cw, ch, b, ow, oh, &full_bar[stage]);
                tma2_load3(sB, &p.tm_b28_hi, c0, n0, wslab, &full_bar[stage]);
                tma2_load3(sB + Cfg::B_PLANE, &p.tm_b28_lo, c0, n0, wslab, &full_bar[stage]);
              } else {                                       // fp16: two 64-channel tiles
                tma2_im2col(sA, &p.tm_a_hi, c0, cw, ch, b, ow, oh, &full_bar[stage]);
                tma2_im2col(sA + Cfg::A_PLANE, &p.tm_a_hi, c0 + 64, cw, ch, b, ow, oh, &full_bar[stage]);
                tma2_load3(sB, &p.tm_b2_hi, c0, n0, wslab, &full_bar[stage]);
                tma2_load3(sB + Cfg::B_PLANE, &p.tm_b2_hi, c0 + 64, n0, wslab, &full_bar[stage]);
              }
            } else {
              tma2_im2col(sA, &p.tm_a_hi, c0, cw, ch, b, ow, oh, &full_bar[stage]);
              if (NPL == 2) tma2_im2col(sA + Cfg::A_PLANE, &p.tm_a_lo, c0, cw, ch, b, ow, oh, &full_bar[stage]);
              tma2_load3(sB, &p.tm_b2_hi, c0, n0, wslab, &full_bar[stage]);
              if (NPL == 2) tma2_load3(sB + Cfg::B_PLANE, &p.tm_b2_lo, c0, n0, wslab, &full_bar[stage]);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
      // tail: every stage this CTA filled has been released (the leader's multicast commits have all landed here) before it may exit
      for (int s = 0; s < S; ++s) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(256, BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait_bounded(&tmem_empty_bar[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sA = smem_base + stage * Cfg::STAGE;
            const uint32_t sB = sA + Cfg::PLANES * Cfg::A_PLANE;
            if (NPL == 3) {
              constexpr uint32_t idq = make_idesc_f0(256, BN);
              if (kb < kb_pass) {                            // e4m3 cross products, K = 32 per instruction: a8_hi x b8_lo + a8_lo x b8_hi
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma2_f8(tmem_d, make_desc(sA + k * 32, 16, 1024), make_desc(sB + Cfg::B_PLANE + k * 32, 16, 1024), idq, (kb | k) != 0);
                  umma2_f8(tmem_d, make_desc(sA + Cfg::A_PLANE + k * 32, 16, 1024), make_desc(sB + k * 32, 16, 1024), idq, 1);
                }
              } else {                                       // fp16 hi x hi, two 64-channel tiles of K = 16 steps
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t a = make_desc(sA + h * Cfg::A_PLANE + k * 32, 16, 1024), b = make_desc(sB + h * Cfg::B_PLANE + k * 32, 16, 1024);
                    if (kb == kb_pass && h == 0 && k == 0) umma2_f16_rescale<CGVC_Q_ACC_SHIFT>(tmem_d, a, b, idq);   // D = A*B + D * 2^-15
                    else umma2_bf16(tmem_d, a, b, idq, 1);
                  }
              }
            } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_hi = make_desc(sA + k * 32, 16, 1024);
              const uint64_t b_hi = make_desc(sB + k * 32, 16, 1024);
              umma2_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
              if (NPL == 2) {
                const uint64_t a_lo = make_desc(sA + Cfg::A_PLANE + k * 32, 16, 1024);
                const uint64_t b_lo = make_desc(sB + Cfg::B_PLANE + k * 32, 16, 1024);
                umma2_bf16(tmem_d, a_hi, b_lo, idesc, 1);
                umma2_bf16(tmem_d, a_lo, b_hi, idesc, 1);
              }
            }
            }
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == num_kb - 1) umma2_commit_mc(&tmem_full_bar[as]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if ((warp - 2) < 4 * NG) {
    // ===================== epilogue (warps 2..9 in two groups; 2..5 for the fused-backward forms), this CTA's 128 rows ==========
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2, ew = 4 * grp + q;       // warp group and slot of this warp's patch / tables
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const long long m0 = ((long long)(tile % m_pairs) * 2 + rank) * 128;
      const int n0 = (tile / m_pairs) * BN;
      nt_tile_epilogue<BN, NPL, EPI>(p, M, HW, m0, n0, q, lane, epi_stage[ew], epi_rowp[ew], epi_bc[ew], epi_xch[grp], &tmem_full_bar[as], aphase,
                                     tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN), grp, NG);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // the peer may still be reading its accumulator half / shared memory
  if (warp == 1) tmem_dealloc2<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ TN kernel (wgrad)
// Tile: 128 channels (GEMM M, operand A = X, MN-major) x 256 gradient columns (GEMM N, operand B = G, MN-major);
// the contraction runs over rows m of the forward output grid, 64 rows per stage.
template <int NPL>
struct TNCfg {
  static constexpr int A_PLANE = 64 * 256;              // 64 K-rows x 128 channels x 2 B  (2 MN-atoms side by side: LBO = 8192)
  static constexpr int B_PLANE = 64 * 512;              // 64 K-rows x 256 columns x 2 B  (4 MN-atoms: LBO = 8192)
  static constexpr int STAGE = NPL * (A_PLANE + B_PLANE);
  static constexpr int STAGES = (200 * 1024) / STAGE;   // 2 (x3), 4 (x1)
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

// Persistent like the NT kernel: work items (n-tile, c-tile, tap, K-split) are walked with stride gridDim.x (n fastest, so
// concurrently running CTAs share the same rows of X and dP in L2); the accumulator is double-buffered in TMEM so the
// red.global epilogue of item i overlaps the MMAs of item i+1.
template <int NPL>
__global__ void __launch_bounds__(kNTThreads, 1)
tc_gg_tn_kernel(const __grid_constant__ TcTNParams p) {
  using Cfg = TNCfg<NPL>;
  constexpr int S = Cfg::STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float epi_stage[4][32 * 32];      // per-warp transposition patch of the coalesced row stores
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatherGeom& g = p.g;
  const long long M = (long long)g.B * g.Hy * g.Wx;
  const int HW = g.Hy * g.Wx;
  const int n_tiles = (p.g_ld + BN - 1) / BN, c_tiles = (p.x_ld + 127) / 128;
  const int num_items = n_tiles * c_tiles * g.ntaps * p.ksplit;
  long long chunk_rows = (M + p.ksplit - 1) / p.ksplit;
  chunk_rows = (chunk_rows + 63) / 64 * 64;

  struct Item { int n0, c0, tap; long long mbeg, mend; int num_kb; };
  auto decode = [&](int item) -> Item {
    Item w;
    int n_t = item % n_tiles; int t1 = item / n_tiles;
    int c_t = t1 % c_tiles; int t2 = t1 / c_tiles;
    w.tap = t2 % g.ntaps; int ks = t2 / g.ntaps;
    w.n0 = n_t * BN; w.c0 = c_t * 128;
    w.mbeg = (long long)ks * chunk_rows;
    w.mend = (w.mbeg + chunk_rows < M) ? w.mbeg + chunk_rows : M;
    w.num_kb = w.mend > w.mbeg ? (int)((w.mend - w.mbeg + 63) / 64) : 0;
    return w;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducerThreads + 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_g_hi);
    if (NPL == 2) tma_prefetch_desc(&p.tm_g_lo);
  }
  if (warp == 4) tmem_alloc<2 * BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4) {
    // ---- producers: every K-row (one output position m) is a contiguous run of channels / columns in global memory
    const int t = threadIdx.x;
    const int chunk = t & 7, rsub = t >> 3;                 // 8 threads x 16 B = one 128-byte atom row; 16 rows per pass
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item w = decode(item);
      for (int kb = 0; kb < w.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t sA = smem_base + stage * Cfg::STAGE;
        const uint32_t sB = sA + NPL * Cfg::A_PLANE;
        if (t == 0) {
          // gradient tile: 64 K-rows x 256 columns = 4 MN-atoms of [64 rows][64 cols]; rows >= M are zero-filled by TMA
          // (a stage never straddles two K-splits: split boundaries are multiples of 64 rows)
          const int row0 = (int)(w.mbeg + (long long)kb * 64);
          int natoms = 0;
#pragma unroll
          for (int a = 0; a < 4; ++a) natoms += (w.n0 + a * 64) < p.g_ld ? 1 : 0;
          mbar_expect_tx(&full_bar[stage], NPL * natoms * 8192);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            if ((w.n0 + a * 64) < p.g_ld) {
              tma_load3(sB + a * 8192, &p.tm_g_hi, w.n0 + a * 64, row0, 0, &full_bar[stage]);
              if (NPL == 2) tma_load3(sB + Cfg::B_PLANE + a * 8192, &p.tm_g_lo, w.n0 + a * 64, row0, 0, &full_bar[stage]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kr = rsub + 16 * i;                       // K-row within the stage (0..63)
          const long long m = w.mbeg + (long long)kb * 64 + kr;
          long long xoff = -1;
          if (m < w.mend) {
            const uint32_t mu = (uint32_t)m;
            int b = (int)fdiv(mu, p.div_hw); int rem = (int)(mu - (uint32_t)b * (uint32_t)HW);
            int y = (int)fdiv((uint32_t)rem, p.div_w); int x = rem - y * g.Wx;
            int yy = y * g.sy + g.oy[w.tap], xx = x * g.sx + g.ox[w.tap];
            if (yy >= 0 && yy < g.Hs && xx >= 0 && xx < g.Ws)
              xoff = ((long long)(b * g.Hs + yy) * g.Ws + xx) * p.x_ld + w.c0 + chunk * 8;
          }
          const uint32_t so = sw128(kr, chunk);               // (kr/8)*1024 + (kr%8)*128 + swizzled chunk
#pragma unroll
          for (int a = 0; a < 2; ++a) {                       // 2 channel atoms of 64
            const bool ok = xoff >= 0 && (w.c0 + a * 64) < p.x_ld;
            const long long off = ok ? xoff + a * 64 : 0;
            cp_async16(sA + a * 8192 + so, p.x_hi + off, ok ? 16u : 0u);
            if (NPL == 2) cp_async16(sA + Cfg::A_PLANE + a * 8192 + so, p.x_lo + off, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full_bar[stage]);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item w = decode(item);
      if (w.num_kb == 0) continue;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      ++it;
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < w.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        fence_proxy_async();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_base + stage * Cfg::STAGE;
          const uint32_t sB = sA + NPL * Cfg::A_PLANE;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                      // UMMA_K = 16 K-rows = two 8-row groups = 2048 bytes
            const uint64_t a_hi = make_desc(sA + k * 2048, 8192, 1024);
            const uint64_t b_hi = make_desc(sB + k * 2048, 8192, 1024);
            umma_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
            if (NPL == 2) {
              const uint64_t a_lo = make_desc(sA + Cfg::A_PLANE + k * 2048, 8192, 1024);
              const uint64_t b_lo = make_desc(sB + Cfg::B_PLANE + k * 2048, 8192, 1024);
              umma_bf16(tmem_d, a_hi, b_lo, idesc, 1);
              umma_bf16(tmem_d, a_lo, b_hi, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == w.num_kb - 1) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- epilogue (warps 5..8): atomically accumulate the tile into dW (TF layout [t][c][n]); split-K partials meet there.
    // Rows (= channels) are written back 8 lanes per row through the transposition patch, like the NT kernel's stores.
    const int q = warp & 3;
    float* stg = epi_stage[q];
    const int sc = lane & 7, sr = lane >> 3;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item w = decode(item);
      if (w.num_kb == 0) continue;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      ++it;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int cq = w.c0 + q * 32;                           // first channel row of this warp (TMEM lane == channel row)
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        const int n = w.n0 + cb * 32;
        if (n >= p.N) break;
        float o[32];
        { uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cb * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(v[k]); }
        stage_rows(stg, o, lane);
        float* base; int nn; int ncols;
        if (n < p.n_split) { base = p.dw_a; nn = n; ncols = p.n_split; } else { base = p.dw_g; nn = n - p.n_split; ncols = p.N - p.n_split; }
        if (n + 4 * sc < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = sr + 4 * i;
            const int c = cq + rr;
            if (c < p.C) {
              const float4 val = staged_chunk(stg, rr, sc);
              float* d = tn_dst(base, g.widx[w.tap], p.C, c, ncols, nn + 4 * sc, p.fold_n);
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair TN kernel
// Weight gradient on CTA pairs: one 256-channel x 256-column tile per pair and work item.  CTA r owns channels c0 + 128 r .. (its X
// tile, by two TMA im2col loads of 64 pixels x 64 channels per plane, and its 128 accumulator lanes) and loads gradient columns
// n0 + 128 r .. (two [64 x 64] boxes per plane); 64 KB instead of 96 KB of operands per SM and K-block.
template <int NPL>
struct PairTNCfg {
  static constexpr int A_PLANE = 64 * 256;              // 64 K-rows x 128 channels x 2 B (2 MN-atoms: LBO = 8192)
  static constexpr int B_PLANE = 64 * 256;              // 64 K-rows x 128 columns x 2 B
  static constexpr int STAGE = NPL * (A_PLANE + B_PLANE);
  static constexpr int STAGES = (196 * 1024) / STAGE;   // 3 (x3), 6 (x1)
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

template <int NPL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
tc_pair_tn_kernel(const __grid_constant__ TcTNParams p) {
  using Cfg = PairTNCfg<NPL>;
  constexpr int S = Cfg::STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float epi_stage[4][32 * 32];
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const GatherGeom& g = p.g;
  const long long M = (long long)g.B * g.Hy * g.Wx;
  const int HW = g.Hy * g.Wx;
  const int n_tiles = (p.g_ld + BN - 1) / BN, c_tiles = (p.x_ld + 255) / 256;
  const int num_items = n_tiles * c_tiles * g.ntaps * p.ksplit;
  long long chunk_rows = (M + p.ksplit - 1) / p.ksplit;
  chunk_rows = (chunk_rows + 63) / 64 * 64;

  struct Item { int n0, c0, tap; long long mbeg, mend; int num_kb; };
  auto decode = [&](int item) -> Item {
    Item w;
    int n_t = item % n_tiles; int t1 = item / n_tiles;
    int c_t = t1 % c_tiles; int t2 = t1 / c_tiles;
    w.tap = t2 % g.ntaps; int ks = t2 / g.ntaps;
    w.n0 = n_t * BN; w.c0 = c_t * 256;
    w.mbeg = (long long)ks * chunk_rows;
    w.mend = (w.mbeg + chunk_rows < M) ? w.mbeg + chunk_rows : M;
    w.num_kb = w.mend > w.mbeg ? (int)((w.mend - w.mbeg + 63) / 64) : 0;
    return w;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_g_hi); tma_prefetch_desc(&p.tm_x_hi);
    if (NPL == 2) { tma_prefetch_desc(&p.tm_g_lo); tma_prefetch_desc(&p.tm_x_lo); }
  }
  if (warp == 1) tmem_alloc2<2 * BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = pair; item < num_items; item += npairs) {
        const Item w = decode(item);
        const int cA = w.c0 + (int)rank * 128, nB = w.n0 + (int)rank * 128;
        const unsigned short ow = p.ig.off_w[w.tap], oh = p.ig.off_h[w.tap];
        for (int kb = 0; kb < w.num_kb; ++kb) {
          const long long mrow = w.mbeg + (long long)kb * 64;   // < M: a K-split never starts beyond the tensor
          const uint32_t mu = (uint32_t)mrow;
          const int b = (int)fdiv(mu, p.div_hw); const int rem = (int)(mu - (uint32_t)b * (uint32_t)HW);
          const int y = (int)fdiv((uint32_t)rem, p.div_w); const int x = rem - y * g.Wx;
          const int cw = p.ig.lo_w + x * g.sx, ch = p.ig.lo_h + y * g.sy;
          mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
          const uint32_t sA = smem_base + stage * Cfg::STAGE;
          const uint32_t sB = sA + NPL * Cfg::A_PLANE;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * Cfg::STAGE);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            // channels / columns beyond the tensors are zero-filled by the unit (the byte count stays the same)
            tma2_im2col(sA + a * 8192, &p.tm_x_hi, cA + a * 64, cw, ch, b, ow, oh, &full_bar[stage]);
            if (NPL == 2) tma2_im2col(sA + Cfg::A_PLANE + a * 8192, &p.tm_x_lo, cA + a * 64, cw, ch, b, ow, oh, &full_bar[stage]);
            tma2_load3(sB + a * 8192, &p.tm_g_hi, nB + a * 64, (int)mrow, 0, &full_bar[stage]);
            if (NPL == 2) tma2_load3(sB + Cfg::B_PLANE + a * 8192, &p.tm_g_lo, nB + a * 64, (int)mrow, 0, &full_bar[stage]);
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      for (int s = 0; s < S; ++s) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(256, BN, 1, 1);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = pair; item < num_items; item += npairs) {
        const Item w = decode(item);
        if (w.num_kb == 0) continue;
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        ++it;
        mbar_wait_bounded(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < w.num_kb; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sA = smem_base + stage * Cfg::STAGE;
            const uint32_t sB = sA + NPL * Cfg::A_PLANE;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                    // UMMA_K = 16 K-rows = two 8-row groups = 2048 bytes
              const uint64_t a_hi = make_desc(sA + k * 2048, 8192, 1024);
              const uint64_t b_hi = make_desc(sB + k * 2048, 8192, 1024);
              umma2_bf16(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
              if (NPL == 2) {
                const uint64_t a_lo = make_desc(sA + Cfg::A_PLANE + k * 2048, 8192, 1024);
                const uint64_t b_lo = make_desc(sB + Cfg::B_PLANE + k * 2048, 8192, 1024);
                umma2_bf16(tmem_d, a_hi, b_lo, idesc, 1);
                umma2_bf16(tmem_d, a_lo, b_hi, idesc, 1);
              }
            }
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == w.num_kb - 1) umma2_commit_mc(&tmem_full_bar[as]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ---- epilogue (warps 2..5): this CTA's 128 channel rows of the tile, atomically accumulated into dW (TF layout [t][c][n])
    const int q = warp & 3;
    float* stg = epi_stage[q];
    const int sc = lane & 7, sr = lane >> 3;
    int it = 0;
    for (int item = pair; item < num_items; item += npairs) {
      const Item w = decode(item);
      if (w.num_kb == 0) continue;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      ++it;
      mbar_wait_bounded(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int cq = w.c0 + (int)rank * 128 + q * 32;         // first channel row of this warp (TMEM lane == channel row)
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        const int n = w.n0 + cb * 32;
        if (n >= p.N || cq >= p.C) break;
        float o[32];
        { uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cb * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(v[k]); }
        stage_rows(stg, o, lane);
        float* base; int nn; int ncols;
        if (n < p.n_split) { base = p.dw_a; nn = n; ncols = p.n_split; } else { base = p.dw_g; nn = n - p.n_split; ncols = p.N - p.n_split; }
        if (n + 4 * sc < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = sr + 4 * i;
            const int c = cq + rr;
            if (c < p.C) {
              const float4 val = staged_chunk(stg, rr, sc);
              float* d = tn_dst(base, g.widx[w.tap], p.C, c, ncols, nn + 4 * sc, p.fold_n);
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ CTA-pair TN kernel, F16F8
// Weight gradient in the 2-MMA-unit precision (kernels.cuh): X and the gradient G are kept as fp16 + two scaled e4m3 planes, both
// with the activation-role scales (1, 2^12), so both cross products x8hi * g8lo and x8lo * g8hi carry 2^12.  Per work item the row
// range is walked twice, 128 K-rows per stage: first the e4m3 cross products (K = 32 per MMA, MN-major tiles: one 128-wide atom per
// CTA and plane), then the fp16 product, whose first MMA rescales the accumulator by 2^-12 (scale-input-d).  Same tile, barriers
// and epilogue as tc_pair_tn_kernel.
#define CGVC_Q_WGRAD_SHIFT 12
struct PairTNQCfg {
  static constexpr int A_BYTES = 2 * 128 * 128;         // pass 0: x8hi | x8lo (128 K-rows x 128 B each); pass 1: two 64-channel fp16 atoms
  static constexpr int B_BYTES = 2 * 128 * 128;         // pass 0: g8hi | g8lo; pass 1: two 64-column fp16 atoms
  static constexpr int STAGE = A_BYTES + B_BYTES;       // 64 KB
  static constexpr int STAGES = 3;
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
tc_pair_tn_q_kernel(const __grid_constant__ TcTNParams p) {
  using Cfg = PairTNQCfg;
  constexpr int S = Cfg::STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float epi_stage[4][32 * 32];
  __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const GatherGeom& g = p.g;
  const long long M = (long long)g.B * g.Hy * g.Wx;
  const int HW = g.Hy * g.Wx;
  const int n_tiles = (p.g_ld + BN - 1) / BN, c_tiles = (p.x_ld + 255) / 256;
  const int num_items = n_tiles * c_tiles * g.ntaps * p.ksplit;
  long long chunk_rows = (M + p.ksplit - 1) / p.ksplit;
  chunk_rows = (chunk_rows + 127) / 128 * 128;

  struct Item { int n0, c0, tap; long long mbeg, mend; int num_kb; };
  auto decode = [&](int item) -> Item {
    Item w;
    int n_t = item % n_tiles; int t1 = item / n_tiles;
    int c_t = t1 % c_tiles; int t2 = t1 / c_tiles;
    w.tap = t2 % g.ntaps; int ks = t2 / g.ntaps;
    w.n0 = n_t * BN; w.c0 = c_t * 256;
    w.mbeg = (long long)ks * chunk_rows;
    w.mend = (w.mbeg + chunk_rows < M) ? w.mbeg + chunk_rows : M;
    w.num_kb = w.mend > w.mbeg ? (int)((w.mend - w.mbeg + 127) / 128) : 0;
    return w;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_xq); tma_prefetch_desc(&p.tm_gq); tma_prefetch_desc(&p.tm_x8_hi); tma_prefetch_desc(&p.tm_x8_lo);
    tma_prefetch_desc(&p.tm_g8_hi); tma_prefetch_desc(&p.tm_g8_lo);
  }
  if (warp == 1) tmem_alloc2<2 * BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = pair; item < num_items; item += npairs) {
        const Item w = decode(item);
        const int cA = w.c0 + (int)rank * 128, nB = w.n0 + (int)rank * 128;
        const unsigned short ow = p.ig.off_w[w.tap], oh = p.ig.off_h[w.tap];
        for (int pass = p.w16; pass < 2; ++pass)
        for (int kb = 0; kb < w.num_kb; ++kb) {
          const long long mrow = w.mbeg + (long long)kb * 128;
          const uint32_t mu = (uint32_t)mrow;
          const int b = (int)fdiv(mu, p.div_hw); const int rem = (int)(mu - (uint32_t)b * (uint32_t)HW);
          const int y = (int)fdiv((uint32_t)rem, p.div_w); const int x = rem - y * g.Wx;
          const int cw = p.ig.lo_w + x * g.sx, ch = p.ig.lo_h + y * g.sy;
          mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
          const uint32_t sA = smem_base + stage * Cfg::STAGE;
          const uint32_t sB = sA + Cfg::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * Cfg::STAGE);
          if (pass == 0) {
            tma2_im2col(sA, &p.tm_x8_hi, cA, cw, ch, b, ow, oh, &full_bar[stage]);
            tma2_im2col(sA + 16384, &p.tm_x8_lo, cA, cw, ch, b, ow, oh, &full_bar[stage]);
            tma2_load3(sB, &p.tm_g8_hi, nB, (int)mrow, 0, &full_bar[stage]);
            tma2_load3(sB + 16384, &p.tm_g8_lo, nB, (int)mrow, 0, &full_bar[stage]);
          } else {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              tma2_im2col(sA + a * 16384, &p.tm_xq, cA + a * 64, cw, ch, b, ow, oh, &full_bar[stage]);
              tma2_load3(sB + a * 16384, &p.tm_gq, nB + a * 64, (int)mrow, 0, &full_bar[stage]);
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      for (int s = 0; s < S; ++s) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0) {
      // both operands MN-major; formats 0 / 0 = e4m3 (kind::f8f6f4) or fp16 (kind::f16)
      constexpr uint32_t idq = make_idesc_f0(256, BN) | (1u << 15) | (1u << 16);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = pair; item < num_items; item += npairs) {
        const Item w = decode(item);
        if (w.num_kb == 0) continue;
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        ++it;
        mbar_wait_bounded(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int pass = p.w16; pass < 2; ++pass)
        for (int kb = 0; kb < w.num_kb; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sA = smem_base + stage * Cfg::STAGE;
            const uint32_t sB = sA + Cfg::A_BYTES;
            if (pass == 0) {
              // e4m3, K = 32 K-rows = four 8-row groups = 4096 bytes per instruction; one 128-wide MN atom per CTA and plane
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma2_f8(tmem_d, make_desc(sA + k * 4096, 16384, 1024), make_desc(sB + 16384 + k * 4096, 16384, 1024), idq, (kb | k) != 0);
                umma2_f8(tmem_d, make_desc(sA + 16384 + k * 4096, 16384, 1024), make_desc(sB + k * 4096, 16384, 1024), idq, 1);
              }
            } else {
              // fp16, K = 16 K-rows = 2048 bytes per instruction; two 64-wide MN atoms per CTA (LBO = 16384)
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t a = make_desc(sA + k * 2048, 16384, 1024), b = make_desc(sB + k * 2048, 16384, 1024);
                if (kb == 0 && k == 0 && !p.w16) umma2_f16_rescale<CGVC_Q_WGRAD_SHIFT>(tmem_d, a, b, idq);
                else umma2_bf16(tmem_d, a, b, idq, (kb | k) != 0);
              }
            }
            umma2_commit_mc(&empty_bar[stage]);
            if (pass == 1 && kb == w.num_kb - 1) umma2_commit_mc(&tmem_full_bar[as]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = epi_stage[q];
    const int sc = lane & 7, sr = lane >> 3;
    int it = 0;
    for (int item = pair; item < num_items; item += npairs) {
      const Item w = decode(item);
      if (w.num_kb == 0) continue;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      ++it;
      mbar_wait_bounded(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int cq = w.c0 + (int)rank * 128 + q * 32;
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        const int n = w.n0 + cb * 32;
        if (n >= p.N || cq >= p.C) break;
        float o[32];
        { uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cb * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(v[k]); }
        stage_rows(stg, o, lane);
        float* base; int nn; int ncols;
        if (n < p.n_split) { base = p.dw_a; nn = n; ncols = p.n_split; } else { base = p.dw_g; nn = n - p.n_split; ncols = p.N - p.n_split; }
        if (n + 4 * sc < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = sr + 4 * i;
            const int c = cq + rr;
            if (c < p.C) {
              const float4 val = staged_chunk(stg, rr, sc);
              float* d = tn_dst(base, g.widx[w.tap], p.C, c, ncols, nn + 4 * sc, p.fold_n);
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ weight planes
// Row of the forward weight / bias planes that holds output channel co of a branch (noff = 0: a, != 0: gate) of a layer with cout channels
// per branch.  perm 0: branches one after the other.  perm 1 (gated): 256-row tiles [128 a | 128 g].  perm 2 (gated + pixel shuffle,
// Ch = cout / 2 post-shuffle channels): 256-row tiles [64 a(c) | 64 a(c + Ch) | 64 g(c) | 64 g(c + Ch)] for 64 post-shuffle channels c.
__host__ __device__ __forceinline__ int perm_row(int perm, int co, int cout, int noff) {
  if (perm == 1) return (co >> 7) * 256 + (noff ? 128 : 0) + (co & 127);
  if (perm == 2) { const int Ch = cout >> 1; const int s = co >= Ch ? 1 : 0; const int c = co - s * Ch; return (c >> 6) * 256 + (noff ? 128 : 0) + s * 64 + (c & 63); }
  return noff + co;
}
// element (tap, ci, co) of a TF kernel [taps][cin][cout].  fold_n != 0 (TcLayer::fold): the layer is registered as a 1 x 1 layer whose
// `cout` columns are (t, n) pairs, co = t * fold_n + n, of a kernel that lies in memory as [cout / fold_n taps][cin][fold_n]
__host__ __device__ __forceinline__ long long w_src(int tap, int ci, int co, int cin, int cout, int fold_n) {
  if (fold_n) { const int t = co / fold_n; return ((long long)t * cin + ci) * fold_n + (co - t * fold_n); }
  return ((long long)tap * cin + ci) * cout + co;
}
// TF kernel [taps][cin][cout] (fp32) -> wd[taps][cin][Ntot] (+ column offset) and wf[taps][Ntot][cin], bf16 hi/lo
__global__ void __launch_bounds__(256)
prep_weights_kernel(const float* __restrict__ w, int taps, int cin, int cout, int nt_n, int cin_k, int cin_n, int nt_k, int noff, int perm, int fold_n,
                    __nv_bfloat16* __restrict__ wf_hi, __nv_bfloat16* __restrict__ wf_lo,
                    __nv_bfloat16* __restrict__ wd_hi, __nv_bfloat16* __restrict__ wd_lo) {
  // 32x32 transposing tiles over (cin, cout) for each tap
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    int ci = ci0 + r, co = co0 + tx;
    float v = 0.f;
    if (ci < cin && co < cout) {
      v = w[w_src(tap, ci, co, cin, cout, fold_n)];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      long long o = ((long long)tap * cin_n + ci) * nt_k + noff + co;
      wd_hi[o] = h; wd_lo[o] = l;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int co = co0 + r, ci = ci0 + tx;
    if (ci < cin && co < cout) {
      float v = tile[tx][r];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      // forward rows of gated layers are tile-interleaved: row = (co/128)*256 + branch*128 + co%128 (branch = noff/cout)
      const int nrow = perm_row(perm, co, cout, noff);
      long long o = ((long long)tap * nt_n + nrow) * cin_k + ci;
      wf_hi[o] = h; wf_lo[o] = l;
    }
  }
}

__global__ void copy_bias_kernel(const float* __restrict__ b, float* __restrict__ dst, int n, int noff, int perm) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[perm_row(perm, i, n, noff)] = b[i];
}

// TF kernel [taps][cin][cout] (fp32) -> F16F8 forward planes wq[taps][nt_n][cin_q] (weight scales); 4 input channels per thread
__global__ void __launch_bounds__(256)
prep_weights_q_kernel(const float* __restrict__ w, int taps, int cin, int cout, int nt_n, int cin_q, int noff, int perm, int fold_n,
                      __half* __restrict__ q16, uint8_t* __restrict__ q8hi, uint8_t* __restrict__ q8lo) {
  const int cq = cin / 4;
  const long long total = (long long)taps * cout * cq;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c4 = (int)(i % cq); long long r = i / cq;
    const int co = (int)(r % cout); const int tap = (int)(r / cout);
    const int ci = c4 * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = w[w_src(tap, ci + k, co, cin, cout, fold_n)];
    const int nrow = perm_row(perm, co, cout, noff);
    const long long o = ((long long)tap * nt_n + nrow) * cin_q + ci;
    uint2 hh; uint32_t b_hi, b_lo;
    cgvc_quant4(v, CGVC_Q_W_SHI, CGVC_Q_W_SLO, hh, b_hi, b_lo);
    *reinterpret_cast<uint2*>(q16 + o) = hh;
    *reinterpret_cast<uint32_t*>(q8hi + o) = b_hi;
    *reinterpret_cast<uint32_t*>(q8lo + o) = b_lo;
  }
}

// TF kernel [taps][cin][cout] (fp32) -> F16F8 data-gradient planes wdq[taps][cin_n][nt_q] (K = output columns contiguous, weight
// scales); 4 output columns per thread.  Column of (branch, co) = noff + co, like the bf16 wd planes.
__global__ void __launch_bounds__(256)
prep_weights_qd_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cin_n, int nt_q, int noff, int fold_n,
                       __half* __restrict__ q16, uint8_t* __restrict__ q8hi, uint8_t* __restrict__ q8lo) {
  const int cq = cout / 4;
  const long long total = (long long)taps * cin * cq;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c4 = (int)(i % cq); long long r = i / cq;
    const int ci = (int)(r % cin); const int tap = (int)(r / cin);
    const float4 v4 = *reinterpret_cast<const float4*>(w + w_src(tap, ci, c4 * 4, cin, cout, fold_n));      // (fold_n % 4 == 0: a quad never straddles two taps)
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    const long long o = ((long long)tap * cin_n + ci) * nt_q + noff + c4 * 4;
    uint2 hh; uint32_t b_hi, b_lo;
    cgvc_quant4(v, CGVC_Q_W_SHI, CGVC_Q_W_SLO, hh, b_hi, b_lo);
    *reinterpret_cast<uint2*>(q16 + o) = hh;
    *reinterpret_cast<uint32_t*>(q8hi + o) = b_hi;
    *reinterpret_cast<uint32_t*>(q8lo + o) = b_lo;
  }
}

// ---- F16F8 planes of every layer in ONE launch (per refresh, or per network range) --------------------------------------------------
// The per-layer kernels above cost ~210 launches per train step (forward planes, data-gradient planes and bias of 70 layer branches) and
// prep_weights_q_kernel reads its source with a stride of `cout` floats between neighbouring threads.  This kernel walks a job table --
// one job per layer branch -- in 32-channel x 64-column tiles: the tile is read once, coalesced along the output columns, written to the
// data-gradient planes in that orientation, transposed through shared memory and written to the forward planes as whole 32-byte sectors
// (8 consecutive input channels per thread); the first tile of a job also copies the bias.  Same arithmetic (cgvc_quant4, weight scales),
// bit-identical planes (tests/test_gpu_model.py::test_batched_weight_planes_match_per_layer_kernels).
struct PrepJob {                               // one branch (a or g) of one layer
  long long w_off, b_off;                      // offsets of its TF kernel / bias in the PARAM arena (b_off < 0: no bias, tap-folded layers)
  int taps, cin, cout, nt_n, cin_q, cin_n, nt_q, noff, perm, fold_n;
  __half* q16; uint8_t *q8hi, *q8lo;           // forward planes [taps][nt_n][cin_q]
  __half* dq16; uint8_t *dq8hi, *dq8lo;        // data-gradient planes [taps][cin_n][nt_q] (null: forward-only engine)
  float* bias;                                 // [nt_n], weight-row order
  int tiles_ci, tiles_co;                      // tiles per tap
  int first_block;                             // blocks of the jobs before this one
};

__global__ void __launch_bounds__(256)
prep_weights_q_all_kernel(const float* __restrict__ params, const PrepJob* __restrict__ jobs, int j0, int j1, int block0) {
  __shared__ float tile[32][65];
  const int blk = (int)blockIdx.x + block0;
  int lo = j0, hi = j1 - 1;                     // last job whose first block is <= blk (block-uniform)
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].first_block <= blk) lo = mid; else hi = mid - 1; }
  const PrepJob& J = jobs[lo];
  int r = blk - J.first_block;
  const int tco = r % J.tiles_co; r /= J.tiles_co;
  const int tci = r % J.tiles_ci; const int tap = r / J.tiles_ci;
  const int ci0 = tci * 32, co0 = tco * 64;
  const float* __restrict__ w = params + J.w_off;
  const int t = threadIdx.x;
  {
    const int c4 = (t & 15) * 4;
    for (int rr = t >> 4; rr < 32; rr += 16) {
      const int ci = ci0 + rr, co = co0 + c4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (ci < J.cin && co < J.cout) {           // cout % 4 == 0 (layer_ok): whole quads
        const float4 q = *reinterpret_cast<const float4*>(w + w_src(tap, ci, co, J.cin, J.cout, J.fold_n));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        if (J.dq16) {
          const long long o = ((long long)tap * J.cin_n + ci) * J.nt_q + J.noff + co;
          uint2 hh; uint32_t b_hi, b_lo;
          cgvc_quant4(v, CGVC_Q_W_SHI, CGVC_Q_W_SLO, hh, b_hi, b_lo);
          *reinterpret_cast<uint2*>(J.dq16 + o) = hh;
          *reinterpret_cast<uint32_t*>(J.dq8hi + o) = b_hi;
          *reinterpret_cast<uint32_t*>(J.dq8lo + o) = b_lo;
        }
      }
      tile[rr][c4] = v[0]; tile[rr][c4 + 1] = v[1]; tile[rr][c4 + 2] = v[2]; tile[rr][c4 + 3] = v[3];
    }
  }
  __syncthreads();
  {
    const int lane = t & 31, warp = t >> 5;
    const int cig = (lane & 3) * 8;               // 8 consecutive input channels: 16 bytes of q16, 8 of each e4m3 plane
    const int col = warp * 8 + (lane >> 2);
    const int co = co0 + col;
    if (co < J.cout && ci0 + cig < J.cin_q) {     // channels in [cin, cin_q) are written as zeros (they are zero in the tile)
      float va[4], vb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { va[k] = tile[cig + k][col]; vb[k] = tile[cig + 4 + k][col]; }
      const int nrow = perm_row(J.perm, co, J.cout, J.noff);
      const long long o = ((long long)tap * J.nt_n + nrow) * J.cin_q + ci0 + cig;
      uint2 ha, hb; uint32_t a_hi, a_lo, b_hi, b_lo;
      cgvc_quant4(va, CGVC_Q_W_SHI, CGVC_Q_W_SLO, ha, a_hi, a_lo);
      cgvc_quant4(vb, CGVC_Q_W_SHI, CGVC_Q_W_SLO, hb, b_hi, b_lo);
      *reinterpret_cast<uint4*>(J.q16 + o) = make_uint4(ha.x, ha.y, hb.x, hb.y);
      *reinterpret_cast<uint2*>(J.q8hi + o) = make_uint2(a_hi, b_hi);
      *reinterpret_cast<uint2*>(J.q8lo + o) = make_uint2(a_lo, b_lo);
    }
  }
  if (tap == 0 && tci == 0 && tco == 0 && J.b_off >= 0)
    for (int i = t; i < J.cout; i += 256) J.bias[perm_row(J.perm, i, J.cout, J.noff)] = params[J.b_off + i];
}

// opt-in to > 48 KB dynamic shared memory, once per kernel (never inside a stream capture: see tc_init_kernels)
template <class K>
cudaError_t set_smem(K kernel, int bytes) {
  static std::vector<const void*> done;
  const void* key = (const void*)kernel;
  for (const void* d : done) if (d == key) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.push_back(key);
  return e;
}

// ---- optional per-launch event timing (bench.py roofline): class 0 = NT (fwd/dgrad, plain epilogue), 1 = TN (wgrad),
//      2 = NT with the fused instance-norm epilogue
struct ProfRec { cudaEvent_t a, b; double flops; int cls; long long M; int N, K; };
int g_tc_debug = 0;
int g_tc_pair = 1;            // CTA-pair kernels (cta_group::2 + TMA im2col) where the shape allows; 0: one-CTA kernels only
int g_tc_prep_batched = 1;    // F16F8 weight planes of all layers by prep_weights_q_all_kernel (one launch); 0: the per-layer kernels
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
void prof_begin(cudaStream_t st, double flops, int cls, long long M = 0, int N = 0, int K = 0) {
  if (!g_prof_on) return;
  ProfRec r; r.flops = flops; r.cls = cls; r.M = M; r.N = N; r.K = K;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { if (g_prof_on && !g_prof.empty()) cudaEventRecord(g_prof.back().b, st); }

// output-column tile width: 256 where the padded width allows, 32 for the 24-column edge layers (G.o1 forward, G.h1 data
// gradient: a 128-wide tile would spend 5x the MMAs on zero padding), else 128
inline int tile_rows(int n_real, int n_padded) { return (n_padded % 256 == 0) ? 256 : (n_real <= 32 ? 32 : 128); }

cudaError_t launch_nt(TcNTParams p, int precision, cudaStream_t st, int epi, bool pair_ok = false) {
  const long long M = (long long)p.g.B * p.g.Hy * p.g.Wx;
  if (M == 0) return cudaSuccess;
  const bool x3 = precision == 1;
  if (p.g.ntaps == 0) return cudaErrorInvalidValue;      // empty contractions are the caller's business
  const int bn = tile_rows(p.N, p.Nw);
  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
  p.n_tiles = bn == 32 ? 1 : p.Nw / bn;                  // the 32-wide tile only ever covers the (<= 32) real columns
  const long long tiles = ((M + 127) / 128) * p.n_tiles;
  dim3 grid((unsigned)(tiles < num_sms ? tiles : num_sms));
  cudaError_t e;
  ++g_cgvc_launches;
  p.debug = g_tc_debug;
  prof_begin(st, 2.0 * (double)M * p.N * p.g.ntaps * p.C, (epi == 1 || epi == 2 || epi == 5) ? 2 : 0, M, p.N, p.g.ntaps * p.C);
  if (pair_ok && g_tc_pair && bn != 32 && M < (1ll << 31)) {
    // CTA pairs: 256 x bn tile per cluster of 2, persistent over min(#pair tiles, #SM pairs) clusters
    const long long npairs = num_sms / 2;
    const long long m_pairs = ((M + 127) / 128 + 1) / 2;
    int pbn = bn;
    if (epi == 0 && bn == 256 && p.have_b64 && !p.perm) {     // (gated forward layers keep their [128 a | 128 g] 256-wide tiles)
      // wave quantisation: a launch whose 256-wide tiles fill the last round of the persistent grid badly runs 128-wide tiles
      // instead (twice the tiles, 1.5x the operand bytes per FLOP: worth it only for a clearly better fill)
      const long long t256 = m_pairs * (p.Nw / 256), t128 = m_pairs * (p.Nw / 128);
      const double e256 = (double)t256 / (double)(((t256 + npairs - 1) / npairs) * npairs);
      const double e128 = (double)t128 / (double)(((t128 + npairs - 1) / npairs) * npairs);
      if (0.9 * e128 > e256) { pbn = 128; p.n_tiles = p.Nw / 128; p.tm_b2_hi = p.tm_b64_hi; p.tm_b2_lo = p.tm_b64_lo; }
    }
    const long long ptiles = m_pairs * p.n_tiles;
    dim3 pgrid((unsigned)(2 * (ptiles < npairs ? ptiles : npairs)));
#define LAUNCH_PAIR(BN_, NPL_, EPI_)                                                              \
  do {                                                                                            \
    e = set_smem(tc_pair_nt_kernel<BN_, NPL_, EPI_>, PairCfg<BN_, NPL_>::SMEM);                   \
    if (e != cudaSuccess) return e;                                                               \
    tc_pair_nt_kernel<BN_, NPL_, EPI_><<<pgrid, kPairNTThreads, PairCfg<BN_, NPL_>::SMEM, st>>>(p); \
  } while (0)
    if (epi != 0 && bn != 256) return cudaErrorInvalidValue;
    if (precision == 3) {                                   // F16F8 (no fused backward epilogues in this precision)
      if (epi == 1)        LAUNCH_PAIR(256, 3, 1);
      else if (epi == 2)   LAUNCH_PAIR(256, 3, 2);
      else if (epi == 5)   LAUNCH_PAIR(256, 3, 5);
      else if (epi != 0)   return cudaErrorInvalidValue;
      else if (pbn == 256) LAUNCH_PAIR(256, 3, 0);
      else                 LAUNCH_PAIR(128, 3, 0);
    }
    else if (epi == 1)  { if (x3) LAUNCH_PAIR(256, 2, 1); else LAUNCH_PAIR(256, 1, 1); }
    else if (epi == 2)  { if (x3) LAUNCH_PAIR(256, 2, 2); else LAUNCH_PAIR(256, 1, 2); }
    else if (epi == 3)  { if (x3) LAUNCH_PAIR(256, 2, 3); else LAUNCH_PAIR(256, 1, 3); }
    else if (epi == 4)  { if (x3) LAUNCH_PAIR(256, 2, 4); else LAUNCH_PAIR(256, 1, 4); }
    else if (epi == 5)  { if (x3) LAUNCH_PAIR(256, 2, 5); else LAUNCH_PAIR(256, 1, 5); }
    else if (pbn == 256) { if (x3) LAUNCH_PAIR(256, 2, 0); else LAUNCH_PAIR(256, 1, 0); }
    else                 { if (x3) LAUNCH_PAIR(128, 2, 0); else LAUNCH_PAIR(128, 1, 0); }
#undef LAUNCH_PAIR
    prof_end(st);
    return cudaGetLastError();
  }
#define LAUNCH_NT(BN_, NPL_, EPI_)                                                                \
  do {                                                                                            \
    e = set_smem(tc_gg_nt_kernel<BN_, NPL_, EPI_>, NTCfg<BN_, NPL_>::SMEM);                       \
    if (e != cudaSuccess) return e;                                                               \
    tc_gg_nt_kernel<BN_, NPL_, EPI_><<<grid, kNTThreads, NTCfg<BN_, NPL_>::SMEM, st>>>(p);        \
  } while (0)
  if (epi != 0 && bn != 256) return cudaErrorInvalidValue;
  if (precision == 3) {                                     // F16F8: forward form only
    if (epi == 1)       LAUNCH_NT(256, 3, 1);
    else if (epi == 2)  LAUNCH_NT(256, 3, 2);
    else if (epi == 5)  LAUNCH_NT(256, 3, 5);
    else if (epi != 0)  return cudaErrorInvalidValue;
    else if (bn == 256) LAUNCH_NT(256, 3, 0);
    else if (bn == 128) LAUNCH_NT(128, 3, 0);
    else                LAUNCH_NT(32, 3, 0);
  }
  else if (epi == 1)  { if (x3) LAUNCH_NT(256, 2, 1); else LAUNCH_NT(256, 1, 1); }
  else if (epi == 2)  { if (x3) LAUNCH_NT(256, 2, 2); else LAUNCH_NT(256, 1, 2); }
  else if (epi == 3)  { if (x3) LAUNCH_NT(256, 2, 3); else LAUNCH_NT(256, 1, 3); }
  else if (epi == 4)  { if (x3) LAUNCH_NT(256, 2, 4); else LAUNCH_NT(256, 1, 4); }
  else if (epi == 5)  { if (x3) LAUNCH_NT(256, 2, 5); else LAUNCH_NT(256, 1, 5); }
  else if (bn == 256) { if (x3) LAUNCH_NT(256, 2, 0); else LAUNCH_NT(256, 1, 0); }
  else if (bn == 128) { if (x3) LAUNCH_NT(128, 2, 0); else LAUNCH_NT(128, 1, 0); }
  else                { if (x3) LAUNCH_NT(32, 2, 0);  else LAUNCH_NT(32, 1, 0); }
#undef LAUNCH_NT
  prof_end(st);
  return cudaGetLastError();
}

cudaError_t launch_tn(TcTNParams p, int precision, cudaStream_t st, bool pair_ok = false) {
  const long long M = (long long)p.g.B * p.g.Hy * p.g.Wx;
  if (M == 0) return cudaSuccess;
  const bool x3 = precision == 1;
  if (M >= (1ll << 31)) return cudaErrorInvalidValue;
  // CTA pairs (256-channel x 256-column tiles) where both extents fill them
  const bool pair = pair_ok && g_tc_pair && precision != 3 && p.x_ld % 256 == 0 && p.g_ld % 256 == 0;
  int tiles = ((p.g_ld + 255) / 256) * ((p.x_ld + (pair ? 255 : 127)) / (pair ? 256 : 128)) * p.g.ntaps;
  static int num_sms_dev = 0;
  if (!num_sms_dev) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms_dev, cudaDevAttrMultiProcessorCount, dev); }
  const int num_sms = pair ? num_sms_dev / 2 : num_sms_dev;   // persistent CTAs (or CTA pairs) the work items are spread over
  // split the row (contraction) range so that the work items fill whole rounds of the persistent grid; every item keeps
  // >= 16 stages so that its red.global epilogue hides behind the next item's MMAs
  long long maxsplit = M / 1024; if (maxsplit < 1) maxsplit = 1; if (maxsplit > 32) maxsplit = 32;
  int ksplit = 1; double best = 0.0;
  for (int ks = 1; ks <= (int)maxsplit; ++ks) {
    long long items = (long long)tiles * ks;
    double eff = (double)items / (double)(((items + num_sms - 1) / num_sms) * num_sms);
    if (items < num_sms) eff *= 0.5;                       // a single partial round: prefer more, smaller items
    if (eff > best + 0.02) { best = eff; ksplit = ks; }
  }
  p.ksplit = ksplit;
  p.div_hw = make_fastdiv((uint32_t)(p.g.Hy * p.g.Wx)); p.div_w = make_fastdiv((uint32_t)p.g.Wx);
  const long long items = (long long)tiles * ksplit;
  dim3 grid((unsigned)(items < num_sms ? items : num_sms));
  cudaError_t e;
  ++g_cgvc_launches;
  prof_begin(st, 2.0 * (double)M * p.N * p.g.ntaps * p.C, 1, M, p.N, p.g.ntaps * p.C);
  if (pair) {
    dim3 pgrid((unsigned)(2 * (items < num_sms ? items : num_sms)));     // num_sms already counts pairs here
    if (x3) {
      e = set_smem(tc_pair_tn_kernel<2>, PairTNCfg<2>::SMEM); if (e != cudaSuccess) return e;
      tc_pair_tn_kernel<2><<<pgrid, kPairThreads, PairTNCfg<2>::SMEM, st>>>(p);
    } else {
      e = set_smem(tc_pair_tn_kernel<1>, PairTNCfg<1>::SMEM); if (e != cudaSuccess) return e;
      tc_pair_tn_kernel<1><<<pgrid, kPairThreads, PairTNCfg<1>::SMEM, st>>>(p);
    }
    prof_end(st);
    return cudaGetLastError();
  }
  if (x3) {
    e = set_smem(tc_gg_tn_kernel<2>, TNCfg<2>::SMEM); if (e != cudaSuccess) return e;
    tc_gg_tn_kernel<2><<<grid, kNTThreads, TNCfg<2>::SMEM, st>>>(p);
  } else {
    e = set_smem(tc_gg_tn_kernel<1>, TNCfg<1>::SMEM); if (e != cudaSuccess) return e;
    tc_gg_tn_kernel<1><<<grid, kNTThreads, TNCfg<1>::SMEM, st>>>(p);
  }
  prof_end(st);
  return cudaGetLastError();
}

cudaError_t launch_tn_q(TcTNParams p, cudaStream_t st) {
  const long long M = (long long)p.g.B * p.g.Hy * p.g.Wx;
  if (M == 0) return cudaSuccess;
  static int num_sms_dev = 0;
  if (!num_sms_dev) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms_dev, cudaDevAttrMultiProcessorCount, dev); }
  const int npairs = num_sms_dev / 2;
  const int tiles = ((p.g_ld + 255) / 256) * ((p.x_ld + 255) / 256) * p.g.ntaps;
  long long maxsplit = M / 2048; if (maxsplit < 1) maxsplit = 1; if (maxsplit > 32) maxsplit = 32;     // every item keeps >= 16 stages of 128 rows
  int ksplit = 1; double best = 0.0;
  for (int ks = 1; ks <= (int)maxsplit; ++ks) {
    long long items = (long long)tiles * ks;
    double eff = (double)items / (double)(((items + npairs - 1) / npairs) * npairs);
    if (items < npairs) eff *= 0.5;
    if (eff > best + 0.02) { best = eff; ksplit = ks; }
  }
  p.ksplit = ksplit;
  p.div_hw = make_fastdiv((uint32_t)(p.g.Hy * p.g.Wx)); p.div_w = make_fastdiv((uint32_t)p.g.Wx);
  const long long items = (long long)tiles * ksplit;
  dim3 pgrid((unsigned)(2 * (items < npairs ? items : npairs)));
  ++g_cgvc_launches;
  prof_begin(st, 2.0 * (double)M * p.N * p.g.ntaps * p.C, 1, M, p.N, p.g.ntaps * p.C);
  cudaError_t e = set_smem(tc_pair_tn_q_kernel, PairTNQCfg::SMEM); if (e != cudaSuccess) return e;
  tc_pair_tn_q_kernel<<<pgrid, kPairThreads, PairTNQCfg::SMEM, st>>>(p);
  prof_end(st);
  return cudaGetLastError();
}

inline int ru(int v, int m) { return (v + m - 1) / m * m; }
inline int Ntot(const TcLayer& L) { return L.cout * (L.gated ? 2 : 1); }
// padded extents: *_k = as a contraction dimension (multiple of 64), *_n = as an output-tile dimension (multiple of 128)
inline int cin_k(const TcLayer& L) { return ru(L.cin, 64); }
inline int cin_n(const TcLayer& L) { return ru(L.cin, 128); }
inline int nt_k(const TcLayer& L) { return ru(Ntot(L), 64); }
inline int nt_n(const TcLayer& L) { return ru(Ntot(L), 128); }
inline int cin_q(const TcLayer& L) { return ru(L.cin, 128); }
inline int nt_q(const TcLayer& L) { return ru(Ntot(L), 128); }      // output columns as an F16F8 contraction dimension (128-byte e4m3 lines)
inline size_t wq_elems(const TcLayer& L) { return (size_t)L.kh * L.kw * nt_n(L) * cin_q(L); }
inline size_t wdq_elems(const TcLayer& L) { return (size_t)L.kh * L.kw * cin_n(L) * nt_q(L); }
inline size_t wf_elems(const TcLayer& L) { return (size_t)L.kh * L.kw * nt_n(L) * cin_k(L); }
inline size_t wd_elems(const TcLayer& L) { return (size_t)L.kh * L.kw * cin_n(L) * nt_k(L); }
// gated layers whose branch width is a multiple of 128 keep their forward weight rows tile-interleaved (see TcNTParams::perm)
inline int layer_perm(const TcLayer& L) {
  if (L.gated && L.shuffle == 2 && L.cout % 128 == 0) return 2;      // see perm_row
  return (L.gated && L.cout % 128 == 0) ? 1 : 0;
}
inline bool layer_ok(const TcLayer& L) {
  if (L.fold && (L.gated || L.kh * L.kw != 1 || L.cout % L.fold || (L.cout / L.fold) % 4)) return false;      // see TcLayer::fold
  return L.kh * L.kw <= CGVC_MAX_TAPS && L.cin % 4 == 0 && Ntot(L) % 4 == 0;
}

// TMA descriptors of a layer's weight planes (call after wf_/wd_ pointers are set)
bool make_layer_maps(TcLayer& L) {
  const int taps = L.kh * L.kw;
  const int bf = tile_rows(Ntot(L), nt_n(L)), bd = tile_rows(L.cin, cin_n(L));   // must match launch_nt's choice of BN
  // pair kernels: each CTA of a pair loads half of the weight tile (tiles there are 256 or 128 rows wide, never 32)
  const int bf2 = (bf == 256 ? 256 : 128) / 2, bd2 = (bd == 256 ? 256 : 128) / 2;
  return make_tmap3(&L.tm_f_hi, L.wf_hi, cin_k(L), nt_n(L), taps, bf) &&
         make_tmap3(&L.tm_f_lo, L.wf_lo, cin_k(L), nt_n(L), taps, bf) &&
         make_tmap3(&L.tm_d_hi, L.wd_hi, nt_k(L), cin_n(L), taps, bd) &&
         make_tmap3(&L.tm_d_lo, L.wd_lo, nt_k(L), cin_n(L), taps, bd) &&
         make_tmap3(&L.tm_f2_hi, L.wf_hi, cin_k(L), nt_n(L), taps, bf2) &&
         make_tmap3(&L.tm_f2_lo, L.wf_lo, cin_k(L), nt_n(L), taps, bf2) &&
         make_tmap3(&L.tm_d2_hi, L.wd_hi, nt_k(L), cin_n(L), taps, bd2) &&
         make_tmap3(&L.tm_d2_lo, L.wd_lo, nt_k(L), cin_n(L), taps, bd2) &&
         make_tmap3(&L.tm_f64_hi, L.wf_hi, cin_k(L), nt_n(L), taps, 64) &&
         make_tmap3(&L.tm_f64_lo, L.wf_lo, cin_k(L), nt_n(L), taps, 64) &&
         make_tmap3(&L.tm_d64_hi, L.wd_hi, nt_k(L), cin_n(L), taps, 64) &&
         make_tmap3(&L.tm_d64_lo, L.wd_lo, nt_k(L), cin_n(L), taps, 64);
}

// TMA im2col maps of the gathered operand planes (pair kernels); false if the geometry cannot be expressed
bool make_gather_maps(TcNTParams& p, int precision = 1) {
  p.ig = im2col_geom(p.g);
  if (!p.ig.ok) return false;
  if (precision == 3) {                                      // F16F8: fp16 plane + two e4m3 planes
    return make_im2col_map(&p.tm_a_hi, p.a_hi, p.g, p.ig, p.C, p.a_ld, 128, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2) &&
           make_im2col_map(&p.tm_a8_hi, p.a8_hi, p.g, p.ig, p.C, p.a_ld, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1) &&
           make_im2col_map(&p.tm_a8_lo, p.a8_lo, p.g, p.ig, p.C, p.a_ld, 128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  }
  if (!make_im2col_map(&p.tm_a_hi, p.a_hi, p.g, p.ig, p.C, p.a_ld, 128)) return false;
  if (p.a_lo && !make_im2col_map(&p.tm_a_lo, p.a_lo, p.g, p.ig, p.C, p.a_ld, 128)) return false;
  return true;
}
bool make_layer_maps_q(TcLayer& L) {
  const int taps = L.kh * L.kw;
  const int bf = tile_rows(Ntot(L), nt_n(L)), bd = tile_rows(L.cin, cin_n(L));
  const int bf2 = (bf == 256 ? 256 : 128) / 2, bd2 = (bd == 256 ? 256 : 128) / 2;      // half tiles of the pair kernels
  const CUtensorMapDataType F16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16, U8 = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  bool ok = make_tmap3_t(&L.tm_q16, L.wq16, F16, 2, cin_q(L), nt_n(L), taps, 64, bf) &&
            make_tmap3_t(&L.tm_q8hi, L.wq8hi, U8, 1, cin_q(L), nt_n(L), taps, 128, bf) &&
            make_tmap3_t(&L.tm_q8lo, L.wq8lo, U8, 1, cin_q(L), nt_n(L), taps, 128, bf) &&
            make_tmap3_t(&L.tm_q16h, L.wq16, F16, 2, cin_q(L), nt_n(L), taps, 64, bf2) &&
            make_tmap3_t(&L.tm_q8hih, L.wq8hi, U8, 1, cin_q(L), nt_n(L), taps, 128, bf2) &&
            make_tmap3_t(&L.tm_q8loh, L.wq8lo, U8, 1, cin_q(L), nt_n(L), taps, 128, bf2);
  if (ok && L.wdq16)
    ok = make_tmap3_t(&L.tm_dq16, L.wdq16, F16, 2, nt_q(L), cin_n(L), taps, 64, bd) &&
         make_tmap3_t(&L.tm_dq8hi, L.wdq8hi, U8, 1, nt_q(L), cin_n(L), taps, 128, bd) &&
         make_tmap3_t(&L.tm_dq8lo, L.wdq8lo, U8, 1, nt_q(L), cin_n(L), taps, 128, bd) &&
         make_tmap3_t(&L.tm_dq16h, L.wdq16, F16, 2, nt_q(L), cin_n(L), taps, 64, bd2) &&
         make_tmap3_t(&L.tm_dq8hih, L.wdq8hi, U8, 1, nt_q(L), cin_n(L), taps, 128, bd2) &&
         make_tmap3_t(&L.tm_dq8loh, L.wdq8lo, U8, 1, nt_q(L), cin_n(L), taps, 128, bd2);
  return ok;
}

int refresh_layer(TcLayer& L, const float* ka, const float* kg, const float* ba, const float* bg, cudaStream_t st) {
  const int taps = L.kh * L.kw;
  dim3 grid((L.cout + 31) / 32, (L.cin + 31) / 32, taps);
  g_cgvc_launches += L.gated ? 4 : 2;
  const int perm = layer_perm(L);
  const int fn = L.fold ? L.cout / L.fold : 0;          // tap-folded layer: columns are (t, n) pairs of a [fold][cin][fn] kernel; no bias of its own
  // (an engine in the F16F8 precision never reads the bf16 planes: only the quantised planes below are refreshed)
  if (!L.wq16) prep_weights_kernel<<<grid, 256, 0, st>>>(ka, taps, L.cin, L.cout, nt_n(L), cin_k(L), cin_n(L), nt_k(L), 0, perm, fn, L.wf_hi, L.wf_lo, L.wd_hi, L.wd_lo);
  if (!L.fold) copy_bias_kernel<<<(L.cout + 255) / 256, 256, 0, st>>>(ba, L.bias, L.cout, 0, perm);
  if (L.gated) {
    if (!L.wq16) prep_weights_kernel<<<grid, 256, 0, st>>>(kg, taps, L.cin, L.cout, nt_n(L), cin_k(L), cin_n(L), nt_k(L), L.cout, perm, fn, L.wf_hi, L.wf_lo, L.wd_hi, L.wd_lo);
    copy_bias_kernel<<<(L.cout + 255) / 256, 256, 0, st>>>(bg, L.bias, L.cout, L.cout, perm);
  }
  if (L.wq16) {
    long long tot = (long long)taps * L.cout * (L.cin / 4); long long nb = (tot + 255) / 256; if (nb > 148 * 32) nb = 148 * 32;
    g_cgvc_launches += L.gated ? 2 : 1;
    prep_weights_q_kernel<<<(unsigned)nb, 256, 0, st>>>(ka, taps, L.cin, L.cout, nt_n(L), cin_q(L), 0, perm, fn, (__half*)L.wq16, L.wq8hi, L.wq8lo);
    if (L.gated) prep_weights_q_kernel<<<(unsigned)nb, 256, 0, st>>>(kg, taps, L.cin, L.cout, nt_n(L), cin_q(L), L.cout, perm, fn, (__half*)L.wq16, L.wq8hi, L.wq8lo);
    if (L.wdq16) {
      g_cgvc_launches += L.gated ? 2 : 1;
      prep_weights_qd_kernel<<<(unsigned)nb, 256, 0, st>>>(ka, taps, L.cin, L.cout, cin_n(L), nt_q(L), 0, fn, (__half*)L.wdq16, L.wdq8hi, L.wdq8lo);
      if (L.gated) prep_weights_qd_kernel<<<(unsigned)nb, 256, 0, st>>>(kg, taps, L.cin, L.cout, cin_n(L), nt_q(L), L.cout, fn, (__half*)L.wdq16, L.wdq8hi, L.wdq8lo);
    }
  }
  return (int)cudaGetLastError();
}

// x planes: [n,H,W,cin_k] (channels beyond cin are zero)
int layer_fwd(const TcLayer& L, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo, int n, int H, int W, int sh, int sw,
              float* P, cudaStream_t st, const TcFuse* fuse = nullptr, bool* fused_out = nullptr) {
  if (!layer_ok(L)) return TC_UNSUPPORTED;
  TcNTParams p; memset(&p, 0, sizeof p);
  p.g = fwd_geom(n, H, W, L.kh, L.kw, sh, sw);
  p.a_hi = xhi; p.a_lo = xlo; p.a_ld = cin_k(L); p.C = cin_k(L);
  p.b_hi = L.wf_hi; p.b_lo = L.wf_lo; p.Nw = nt_n(L); p.N = Ntot(L);
  p.dst = P; p.d_ld = Ntot(L); p.bias = L.bias; p.accumulate = 0;
  p.perm = layer_perm(L); p.Cc = L.cout;
  p.tm_b_hi = L.tm_f_hi; p.tm_b_lo = L.tm_f_lo;
  if (precision == 3) {
    // F16F8: x planes are [rows, cin_q] -- xhi = q16, xlo = q8hi followed by q8lo (kernels.cuh)
    if (!L.wq16) return TC_UNSUPPORTED;
    p.a_ld = cin_q(L); p.C = cin_q(L); p.a_lo = nullptr;
    p.a8_hi = reinterpret_cast<const uint8_t*>(xlo); p.a8_lo = p.a8_hi + (size_t)n * H * W * cin_q(L);
    p.tm_b_hi = L.tm_q16; p.tm_b8_hi = L.tm_q8hi; p.tm_b8_lo = L.tm_q8lo;
  }
  int epi = 0;
  if (fuse && fuse->R > 0) {
    // fused instance-norm epilogue: 1-D layer, whole samples per 128-row tile, 256-wide tiles
    const bool shape_ok = H == 1 && (fuse->R == 32 || fuse->R == 64 || fuse->R == 128) && p.g.Wx == fuse->R && nt_n(L) % 256 == 0 && Ntot(L) == nt_n(L);
    if (shape_ok && L.gated && p.perm == 1) epi = 1; else if (shape_ok && L.gated && p.perm == 2) epi = 5; else if (shape_ok && !L.gated) epi = 2;
    if (epi) {
      p.R = fuse->R; p.gamma_a = fuse->gamma_a; p.beta_a = fuse->beta_a; p.gamma_g = fuse->gamma_g; p.beta_g = fuse->beta_g;
      p.stats = fuse->stats; p.resid = fuse->resid; p.y = fuse->y; p.y_hi = fuse->y_hi; p.y_lo = fuse->y_lo;
      p.y8 = reinterpret_cast<uint8_t*>(fuse->y_lo);
      p.C_out = epi == 5 ? L.cout / 2 : (L.gated ? L.cout : Ntot(L));
      if (!p.y_hi || (epi == 2 && !p.resid)) epi = 0;
    }
  }
  if (fused_out) *fused_out = epi != 0;
  if (!epi && !P) return (int)cudaErrorInvalidValue;          // only the fused epilogues can do without the pre-norm output
  bool pair_ok = false;
  if (g_tc_pair && tile_rows(p.N, p.Nw) != 32) {
    if (precision == 3) { p.tm_b2_hi = L.tm_q16h; p.tm_b28_hi = L.tm_q8hih; p.tm_b28_lo = L.tm_q8loh; }
    else {
      p.tm_b2_hi = L.tm_f2_hi; p.tm_b2_lo = L.tm_f2_lo;
      p.tm_b64_hi = L.tm_f64_hi; p.tm_b64_lo = L.tm_f64_lo; p.have_b64 = 1;
    }
    pair_ok = make_gather_maps(p, precision);
  }
  return (int)launch_nt(p, precision, st, epi, pair_ok);
}

// dP planes: [rows_out, nt_k]
int layer_dgrad(const TcLayer& L, int precision, const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, int n, int H, int W, int sh, int sw,
                float* dx, int accumulate, cudaStream_t st, const TcBwdFuse* fuse = nullptr, bool* fused_out = nullptr) {
  if (fused_out) *fused_out = false;
  if (!layer_ok(L) || (precision == 3 && !L.wdq16)) return TC_UNSUPPORTED;     // F16F8 needs the data-gradient planes (training engines)
  std::vector<GatherGeom> gs = dgrad_geoms(n, H, W, L.kh, L.kw, sh, sw);
  for (const GatherGeom& g : gs) if (g.ntaps == 0) return TC_UNSUPPORTED;     // (never the case for this model's layers)
  // fused backward epilogue: stride-1 1-D layer (one geometry, dense rows), whole samples per 128-row tile, 256-wide tiles
  int epi = 0;
  if (fuse && fuse->R > 0 && gs.size() == 1 && H == 1 && sh == 1 && sw == 1 && (fuse->R == 32 || fuse->R == 64 || fuse->R == 128) &&
      gs[0].Wx == fuse->R && cin_n(L) % 256 == 0 && L.cin == cin_n(L) && fuse->bp && fuse->stats && fuse->dp_hi && fuse->dp_lo && fuse->gamma_a &&
      (fuse->gated ? (fuse->gamma_g && fuse->beta_a && fuse->beta_g) : (dx != nullptr)) && precision != 3)
    epi = fuse->gated ? 3 : 4;
  for (const GatherGeom& g : gs) {
    TcNTParams p; memset(&p, 0, sizeof p);
    p.g = g;
    p.a_hi = dPhi; p.a_lo = dPlo; p.a_ld = nt_k(L); p.C = nt_k(L);
    p.b_hi = L.wd_hi; p.b_lo = L.wd_lo; p.Nw = cin_n(L); p.N = L.cin;
    p.dst = dx; p.d_ld = L.cin; p.bias = nullptr; p.accumulate = accumulate;
    p.tm_b_hi = L.tm_d_hi; p.tm_b_lo = L.tm_d_lo;
    if (precision == 3) {
      // F16F8: dP planes are [rows, nt_q] -- dPhi = q16, dPlo = q8hi followed by q8lo (activation-role scales); weights from the wdq planes
      p.a_ld = nt_q(L); p.C = nt_q(L); p.a_lo = nullptr;
      p.a8_hi = reinterpret_cast<const uint8_t*>(dPlo); p.a8_lo = p.a8_hi + (size_t)g.B * g.Hs * g.Ws * nt_q(L);
      p.tm_b_hi = L.tm_dq16; p.tm_b8_hi = L.tm_dq8hi; p.tm_b8_lo = L.tm_dq8lo;
    }
    if (epi) {
      p.R = fuse->R; p.C_out = L.cin; p.stats = const_cast<float*>(fuse->stats);
      p.gamma_a = fuse->gamma_a; p.beta_a = fuse->beta_a; p.gamma_g = fuse->gamma_g; p.beta_g = fuse->beta_g;
      p.bp = fuse->bp; p.bp_ld = fuse->bp_ld; p.dp_hi = fuse->dp_hi; p.dp_lo = fuse->dp_lo; p.dp_ld = fuse->dp_ld;
      p.dbeta_a = fuse->dbeta_a; p.dgamma_a = fuse->dgamma_a; p.dbeta_g = fuse->dbeta_g; p.dgamma_g = fuse->dgamma_g;
    }
    bool pair_ok = false;
    if (g_tc_pair && tile_rows(p.N, p.Nw) != 32) {
      if (precision == 3) { p.tm_b2_hi = L.tm_dq16h; p.tm_b28_hi = L.tm_dq8hih; p.tm_b28_lo = L.tm_dq8loh; }
      else {
        p.tm_b2_hi = L.tm_d2_hi; p.tm_b2_lo = L.tm_d2_lo;
        p.tm_b64_hi = L.tm_d64_hi; p.tm_b64_lo = L.tm_d64_lo; p.have_b64 = 1;
      }
      pair_ok = make_gather_maps(p, precision);
    }
    cudaError_t e = launch_nt(p, precision, st, epi, pair_ok);
    if (e != cudaSuccess) return (int)e;
  }
  if (fused_out) *fused_out = epi != 0;
  return 0;
}

int layer_wgrad(const TcLayer& L, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, int n, int H, int W, int sh, int sw,
                float* dwa, float* dwg, cudaStream_t st, int w16 = 0) {
  if (!layer_ok(L)) return TC_UNSUPPORTED;
  TcTNParams p; memset(&p, 0, sizeof p);
  p.w16 = (precision == 3 && w16) ? 1 : 0;
  p.fold_n = L.fold ? L.cout / L.fold : 0;
  p.g = fwd_geom(n, H, W, L.kh, L.kw, sh, sw);
  if (precision == 3) {
    // F16F8: x planes [rows_in, cin_q] and dP planes [M, nt_q], each q16 + (q8hi | q8lo); always the CTA-pair kernel
    const long long M = (long long)p.g.B * p.g.Hy * p.g.Wx, rows_in = (long long)n * H * W;
    if (M >= (1ll << 31)) return (int)cudaErrorInvalidValue;
    p.x_ld = cin_q(L); p.C = L.cin; p.g_ld = nt_q(L); p.N = Ntot(L);
    p.dw_a = dwa; p.dw_g = dwg; p.n_split = L.cout;
    const uint8_t* x8 = reinterpret_cast<const uint8_t*>(xlo); const uint8_t* g8 = reinterpret_cast<const uint8_t*>(dPlo);
    p.ig = im2col_geom(p.g);
    const CUtensorMapDataType F16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16, U8 = CU_TENSOR_MAP_DATA_TYPE_UINT8;
    if (!p.ig.ok ||
        !make_im2col_map(&p.tm_xq, xhi, p.g, p.ig, p.x_ld, p.x_ld, 128, F16, 2) ||
        !make_im2col_map(&p.tm_x8_hi, x8, p.g, p.ig, p.x_ld, p.x_ld, 128, U8, 1) ||
        !make_im2col_map(&p.tm_x8_lo, x8 + rows_in * p.x_ld, p.g, p.ig, p.x_ld, p.x_ld, 128, U8, 1) ||
        !make_tmap3_t(&p.tm_gq, dPhi, F16, 2, (uint64_t)p.g_ld, (uint64_t)M, 1, 64, 128) ||
        !make_tmap3_t(&p.tm_g8_hi, g8, U8, 1, (uint64_t)p.g_ld, (uint64_t)M, 1, 128, 128) ||
        !make_tmap3_t(&p.tm_g8_lo, g8 + M * p.g_ld, U8, 1, (uint64_t)p.g_ld, (uint64_t)M, 1, 128, 128))
      return TC_UNSUPPORTED;
    return (int)launch_tn_q(p, st);
  }
  p.x_hi = xhi; p.x_lo = xlo; p.x_ld = cin_k(L); p.C = L.cin;
  p.g_hi = dPhi; p.g_lo = dPlo; p.g_ld = nt_k(L); p.N = Ntot(L);
  p.dw_a = dwa; p.dw_g = dwg; p.n_split = L.cout;
  const long long M = (long long)p.g.B * p.g.Hy * p.g.Wx;
  if (!make_tmap3(&p.tm_g_hi, dPhi, (uint64_t)nt_k(L), (uint64_t)M, 1, 64) || !make_tmap3(&p.tm_g_lo, dPlo, (uint64_t)nt_k(L), (uint64_t)M, 1, 64))
    return (int)cudaErrorInvalidValue;
  bool pair_ok = false;
  if (g_tc_pair && p.x_ld % 256 == 0 && p.g_ld % 256 == 0) {
    p.ig = im2col_geom(p.g);
    pair_ok = p.ig.ok && make_im2col_map(&p.tm_x_hi, xhi, p.g, p.ig, p.x_ld, p.x_ld, 64) && make_im2col_map(&p.tm_x_lo, xlo, p.g, p.ig, p.x_ld, p.x_ld, 64);
  }
  return (int)launch_tn(p, precision, st, pair_ok);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ public API
int tc_register(TcWeights& w, size_t ka, size_t kg, size_t ba, size_t bg, int kh, int kw, int cin, int cout, int gated, int shuffle, int fold) {
  TcLayer L{}; L.ka = ka; L.kg = kg; L.ba = ba; L.bg = bg; L.kh = kh; L.kw = kw; L.cin = cin; L.cout = cout; L.gated = gated; L.shuffle = shuffle; L.fold = fold;
  w.layers.push_back(L);
  return (int)w.layers.size() - 1;
}

static cudaError_t tc_init_kernels();

int tc_alloc(TcWeights& w) {
  { cudaError_t ie = tc_init_kernels(); if (ie != cudaSuccess) return (int)ie; }
  size_t total = 0;
  auto rnd = [](size_t b) { return (b + 255) & ~(size_t)255; };
  for (TcLayer& L : w.layers) {
    total += 2 * rnd(wf_elems(L) * sizeof(__nv_bfloat16)) + 2 * rnd(wd_elems(L) * sizeof(__nv_bfloat16)) + rnd((size_t)nt_n(L) * sizeof(float));
    if (w.quant && layer_ok(L)) total += rnd(wq_elems(L) * 2) + 2 * rnd(wq_elems(L));
    if (w.quant && w.quant_bwd && layer_ok(L)) total += rnd(wdq_elems(L) * 2) + 2 * rnd(wdq_elems(L));
  }
  cudaError_t err = cudaMalloc(&w.pool, total);
  if (err != cudaSuccess) return (int)err;
  err = cudaMemset(w.pool, 0, total);               // padded rows / channels / bias entries stay zero forever
  if (err != cudaSuccess) return (int)err;
  w.pool_bytes = total;
  char* p = (char*)w.pool;
  for (TcLayer& L : w.layers) {
    size_t ef = rnd(wf_elems(L) * sizeof(__nv_bfloat16)), ed = rnd(wd_elems(L) * sizeof(__nv_bfloat16));
    L.wf_hi = (__nv_bfloat16*)p; p += ef; L.wf_lo = (__nv_bfloat16*)p; p += ef;
    L.wd_hi = (__nv_bfloat16*)p; p += ed; L.wd_lo = (__nv_bfloat16*)p; p += ed;
    L.bias = (float*)p; p += rnd((size_t)nt_n(L) * sizeof(float));
    if (!make_layer_maps(L)) return (int)cudaErrorInvalidValue;
    L.wq16 = nullptr; L.wq8hi = L.wq8lo = nullptr; L.wdq16 = nullptr; L.wdq8hi = L.wdq8lo = nullptr;
    if (w.quant && layer_ok(L)) {
      L.wq16 = p; p += rnd(wq_elems(L) * 2);
      L.wq8hi = (uint8_t*)p; p += rnd(wq_elems(L)); L.wq8lo = (uint8_t*)p; p += rnd(wq_elems(L));
      if (w.quant_bwd) {
        L.wdq16 = p; p += rnd(wdq_elems(L) * 2);
        L.wdq8hi = (uint8_t*)p; p += rnd(wdq_elems(L)); L.wdq8lo = (uint8_t*)p; p += rnd(wdq_elems(L));
      }
      if (!make_layer_maps_q(L)) return (int)cudaErrorInvalidValue;
    }
  }
  // job table of prep_weights_q_all_kernel: one job per branch of every layer that has F16F8 planes, in layer order
  w.job_first.clear(); w.job_ka.clear();
  if (w.quant) {
    std::vector<PrepJob> jobs;
    int blocks = 0;
    for (TcLayer& L : w.layers) {
      if (!L.wq16) continue;
      for (int br = 0; br < (L.gated ? 2 : 1); ++br) {
        PrepJob J; memset(&J, 0, sizeof J);
        J.w_off = (long long)(br ? L.kg : L.ka); J.b_off = L.fold ? -1 : (long long)(br ? L.bg : L.ba);
        J.taps = L.kh * L.kw; J.cin = L.cin; J.cout = L.cout; J.nt_n = nt_n(L); J.cin_q = cin_q(L); J.cin_n = cin_n(L); J.nt_q = nt_q(L);
        J.noff = br ? L.cout : 0; J.perm = layer_perm(L); J.fold_n = L.fold ? L.cout / L.fold : 0;
        J.q16 = (__half*)L.wq16; J.q8hi = L.wq8hi; J.q8lo = L.wq8lo;
        J.dq16 = (__half*)L.wdq16; J.dq8hi = L.wdq8hi; J.dq8lo = L.wdq8lo;
        J.bias = L.bias;
        J.tiles_ci = (L.cin + 31) / 32; J.tiles_co = (L.cout + 63) / 64;
        J.first_block = blocks;
        blocks += J.taps * J.tiles_ci * J.tiles_co;
        jobs.push_back(J); w.job_first.push_back(J.first_block); w.job_ka.push_back(L.ka);
      }
    }
    w.job_first.push_back(blocks);
    if (!jobs.empty()) {
      err = cudaMalloc(&w.prep_jobs, jobs.size() * sizeof(PrepJob));
      if (err != cudaSuccess) return (int)err;
      err = cudaMemcpy(w.prep_jobs, jobs.data(), jobs.size() * sizeof(PrepJob), cudaMemcpyHostToDevice);
      if (err != cudaSuccess) return (int)err;
    }
  }
  w.ready = false;
  return 0;
}

// configure every tensor-core kernel instantiation up front (so that no attribute call happens during graph capture)
static cudaError_t tc_init_kernels() {
  cudaError_t e;
#define INIT_NT(BN_, NPL_, EPI_) if ((e = set_smem(tc_gg_nt_kernel<BN_, NPL_, EPI_>, NTCfg<BN_, NPL_>::SMEM)) != cudaSuccess) return e;
  INIT_NT(256, 2, 0) INIT_NT(256, 1, 0) INIT_NT(128, 2, 0) INIT_NT(128, 1, 0) INIT_NT(32, 2, 0) INIT_NT(32, 1, 0)
  INIT_NT(256, 2, 1) INIT_NT(256, 1, 1) INIT_NT(256, 2, 2) INIT_NT(256, 1, 2)
  INIT_NT(256, 2, 3) INIT_NT(256, 1, 3) INIT_NT(256, 2, 4) INIT_NT(256, 1, 4)
  INIT_NT(256, 3, 0) INIT_NT(128, 3, 0) INIT_NT(32, 3, 0) INIT_NT(256, 3, 1) INIT_NT(256, 3, 2)
#undef INIT_NT
  if ((e = set_smem(tc_gg_tn_kernel<2>, TNCfg<2>::SMEM)) != cudaSuccess) return e;
  if ((e = set_smem(tc_gg_tn_kernel<1>, TNCfg<1>::SMEM)) != cudaSuccess) return e;
#define INIT_PAIR(BN_, NPL_, EPI_) if ((e = set_smem(tc_pair_nt_kernel<BN_, NPL_, EPI_>, PairCfg<BN_, NPL_>::SMEM)) != cudaSuccess) return e;
  INIT_PAIR(256, 2, 0) INIT_PAIR(256, 1, 0) INIT_PAIR(128, 2, 0) INIT_PAIR(128, 1, 0)
  INIT_PAIR(256, 2, 1) INIT_PAIR(256, 1, 1) INIT_PAIR(256, 2, 2) INIT_PAIR(256, 1, 2)
  INIT_PAIR(256, 2, 3) INIT_PAIR(256, 1, 3) INIT_PAIR(256, 2, 4) INIT_PAIR(256, 1, 4)
#undef INIT_PAIR
  if ((e = set_smem(tc_pair_tn_kernel<2>, PairTNCfg<2>::SMEM)) != cudaSuccess) return e;
  if ((e = set_smem(tc_pair_tn_kernel<1>, PairTNCfg<1>::SMEM)) != cudaSuccess) return e;
#define INIT_PAIR(BN_, NPL_, EPI_) if ((e = set_smem(tc_pair_nt_kernel<BN_, NPL_, EPI_>, PairCfg<BN_, NPL_>::SMEM)) != cudaSuccess) return e;
  INIT_PAIR(256, 3, 0) INIT_PAIR(128, 3, 0) INIT_PAIR(256, 3, 1) INIT_PAIR(256, 3, 2)
#undef INIT_PAIR
  if ((e = set_smem(tc_pair_tn_q_kernel, PairTNQCfg::SMEM)) != cudaSuccess) return e;
  return cudaSuccess;
}

void tc_free(TcWeights& w) {
  if (w.pool) cudaFree(w.pool);
  if (w.prep_jobs) cudaFree(w.prep_jobs);
  w.pool = nullptr; w.prep_jobs = nullptr; w.ready = false;
}

// the jobs [j0, j1) of the batched F16F8 plane kernel (contiguous: the layers of a network are registered together)
static int refresh_jobs(TcWeights& w, const float* params, int j0, int j1, cudaStream_t st) {
  if (j1 <= j0) return 0;
  const int b0 = w.job_first[j0], nb = w.job_first[j1] - b0;
  ++g_cgvc_launches;
  prep_weights_q_all_kernel<<<(unsigned)nb, 256, 0, st>>>(params, (const PrepJob*)w.prep_jobs, j0, j1, b0);
  return (int)cudaGetLastError();
}

int tc_refresh_weights(TcWeights& w, const float* params, cudaStream_t st) {
  if (!w.pool) return 0;
  const bool batched = g_tc_prep_batched && w.prep_jobs;
  for (TcLayer& L : w.layers) {
    if (batched && L.wq16) continue;
    int r = refresh_layer(L, params + L.ka, params + L.kg, params + L.ba, params + L.bg, st);
    if (r != 0) return r;
  }
  if (batched) { int r = refresh_jobs(w, params, 0, (int)w.job_ka.size(), st); if (r != 0) return r; }
  w.ready = true;
  return 0;
}

// the layers whose kernels live in [begin, end) of the parameter arena (one network)
int tc_refresh_weights_range(TcWeights& w, const float* params, size_t begin, size_t end, cudaStream_t st) {
  if (!w.pool) return 0;
  const bool batched = g_tc_prep_batched && w.prep_jobs;
  for (TcLayer& L : w.layers) {
    if (L.ka < begin || L.ka >= end) continue;
    if (batched && L.wq16) continue;
    int r = refresh_layer(L, params + L.ka, params + L.kg, params + L.ba, params + L.bg, st);
    if (r != 0) return r;
  }
  if (batched) {
    const int nj = (int)w.job_ka.size();
    int j = 0;
    while (j < nj) {                             // maximal runs of jobs inside the range (one run per network in practice)
      if (w.job_ka[j] < begin || w.job_ka[j] >= end) { ++j; continue; }
      int k = j; while (k < nj && w.job_ka[k] >= begin && w.job_ka[k] < end) ++k;
      int r = refresh_jobs(w, params, j, k, st); if (r != 0) return r;
      j = k;
    }
  }
  return 0;
}

void tc_set_prep_batched(int v) { g_tc_prep_batched = v != 0; }

int tc_conv_fwd(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                int n, int H, int W, int sh, int sw, float* P, cudaStream_t st) {
  return layer_fwd(w.layers[slot], precision, xhi, xlo, n, H, W, sh, sw, P, st);
}

int tc_conv_fwd_fused(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                      int n, int H, int W, int sh, int sw, float* P, const TcFuse& fuse, bool* fused, cudaStream_t st) {
  return layer_fwd(w.layers[slot], precision, xhi, xlo, n, H, W, sh, sw, P, st, &fuse, fused);
}

int tc_conv_dgrad(TcWeights& w, int slot, int precision, const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo,
                  int n, int H, int W, int sh, int sw, float* dx, int accumulate, cudaStream_t st) {
  return layer_dgrad(w.layers[slot], precision, dPhi, dPlo, n, H, W, sh, sw, dx, accumulate, st);
}

int tc_conv_dgrad_fused(TcWeights& w, int slot, int precision, const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo,
                        int n, int H, int W, int sh, int sw, float* dx, int accumulate, const TcBwdFuse& fuse, bool* fused, cudaStream_t st) {
  return layer_dgrad(w.layers[slot], precision, dPhi, dPlo, n, H, W, sh, sw, dx, accumulate, st, &fuse, fused);
}

int tc_conv_wgrad(TcWeights& w, int slot, int precision, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo,
                  const __nv_bfloat16* dPhi, const __nv_bfloat16* dPlo, int n, int H, int W, int sh, int sw,
                  float* dwa, float* dwg, float* dba, float* dbg, cudaStream_t st) {
  (void)dba; (void)dbg;   // bias gradients are column sums of the fp32 dP: done by the caller (launch_colsum)
  return layer_wgrad(w.layers[slot], precision, xhi, xlo, dPhi, dPlo, n, H, W, sh, sw, dwa, dwg, st, w.wgrad16 ? 1 : 0);
}

bool tc_profile_is_on() { return g_prof_on; }

void tc_profile_enable(int on) {
  for (ProfRec& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = on != 0;
}

// sums per class over everything recorded since tc_profile_enable(1); synchronises the device
int tc_profile_collect(double ms[3], double flops[3], long long launches[3]) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  for (int c = 0; c < 3; ++c) { ms[c] = 0; flops[c] = 0; launches[c] = 0; }
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    ms[r.cls] += t; flops[r.cls] += r.flops; launches[r.cls] += 1;
  }
  return 0;
}

// every recorded launch in order: ms / flops per launch, meta = (class, M rows, N columns, K = taps * channels) x 4 ints
int tc_profile_launches(double* ms, double* flops, long long* meta4, int capacity, int* n_out) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  int n = 0;
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    if (n < capacity) { ms[n] = t; flops[n] = r.flops; meta4[4 * n] = r.cls; meta4[4 * n + 1] = r.M; meta4[4 * n + 2] = r.N; meta4[4 * n + 3] = r.K; }
    ++n;
  }
  if (n_out) *n_out = n;
  return 0;
}
void tc_set_debug(int v) { g_tc_debug = v; }
void tc_set_pair(int v) { g_tc_pair = v != 0; }

// ---- self-contained versions for the unit tests: fp32 in/out, temporary planes ----
namespace {
struct Temp {
  std::vector<void*> ptrs;
  ~Temp() { for (void* p : ptrs) cudaFree(p); }
  template <class T> T* get(size_t n) { void* p = nullptr; if (cudaMalloc(&p, n * sizeof(T) + 256) != cudaSuccess) return nullptr; ptrs.push_back(p); return (T*)p; }
};
}  // namespace

static int adhoc_layer_q(Temp& T, TcLayer& L, cudaStream_t st) {
  L.wq16 = T.get<uint16_t>(wq_elems(L)); L.wq8hi = T.get<uint8_t>(wq_elems(L)); L.wq8lo = T.get<uint8_t>(wq_elems(L));
  if (!L.wq16 || !L.wq8hi || !L.wq8lo) return (int)cudaErrorMemoryAllocation;
  cudaMemsetAsync(L.wq16, 0, wq_elems(L) * 2, st); cudaMemsetAsync(L.wq8hi, 0, wq_elems(L), st); cudaMemsetAsync(L.wq8lo, 0, wq_elems(L), st);
  L.wdq16 = T.get<uint16_t>(wdq_elems(L)); L.wdq8hi = T.get<uint8_t>(wdq_elems(L)); L.wdq8lo = T.get<uint8_t>(wdq_elems(L));
  if (!L.wdq16 || !L.wdq8hi || !L.wdq8lo) return (int)cudaErrorMemoryAllocation;
  cudaMemsetAsync(L.wdq16, 0, wdq_elems(L) * 2, st); cudaMemsetAsync(L.wdq8hi, 0, wdq_elems(L), st); cudaMemsetAsync(L.wdq8lo, 0, wdq_elems(L), st);
  return make_layer_maps_q(L) ? 0 : (int)cudaErrorInvalidValue;
}

static int adhoc_layer(Temp& T, TcLayer& L, cudaStream_t st) {
  L.wf_hi = T.get<__nv_bfloat16>(wf_elems(L)); L.wf_lo = T.get<__nv_bfloat16>(wf_elems(L));
  L.wd_hi = T.get<__nv_bfloat16>(wd_elems(L)); L.wd_lo = T.get<__nv_bfloat16>(wd_elems(L));
  L.bias = T.get<float>(nt_n(L));
  if (!L.wf_hi || !L.wf_lo || !L.wd_hi || !L.wd_lo || !L.bias) return (int)cudaErrorMemoryAllocation;
  cudaMemsetAsync(L.wf_hi, 0, wf_elems(L) * 2, st); cudaMemsetAsync(L.wf_lo, 0, wf_elems(L) * 2, st);
  cudaMemsetAsync(L.wd_hi, 0, wd_elems(L) * 2, st); cudaMemsetAsync(L.wd_lo, 0, wd_elems(L) * 2, st);
  cudaMemsetAsync(L.bias, 0, nt_n(L) * sizeof(float), st);
  if (!make_layer_maps(L)) return (int)cudaErrorInvalidValue;
  return 0;
}

int tc_conv_fwd_adhoc(int precision, const float* x, const float* w, const float* bias, float* y,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, cudaStream_t st) {
  TcLayer L{}; L.kh = kh; L.kw = kw; L.cin = Cin; L.cout = Cout; L.gated = 0;
  if (!layer_ok(L)) return TC_UNSUPPORTED;
  Temp T;
  int r = adhoc_layer(T, L, st); if (r) return r;
  size_t rows = (size_t)B * H * W;
  const int cpad = precision == 3 ? cin_q(L) : cin_k(L);
  if (precision == 3) { r = adhoc_layer_q(T, L, st); if (r) return r; }
  __nv_bfloat16* xhi = T.get<__nv_bfloat16>(rows * cpad); __nv_bfloat16* xlo = T.get<__nv_bfloat16>(rows * cpad);
  float* zero = T.get<float>(Cout);
  if (!xhi || !xlo || !zero) return (int)cudaErrorMemoryAllocation;
  cudaMemsetAsync(zero, 0, Cout * sizeof(float), st);
  r = refresh_layer(L, w, nullptr, bias ? bias : zero, nullptr, st); if (r) return r;
  cudaError_t e = precision == 3 ? launch_pad_split_q(x, (long long)rows, Cin, Cin, cpad, xhi, xlo, st)
                                 : launch_pad_split(x, (long long)rows, Cin, Cin, cpad, xhi, xlo, st);
  if (e != cudaSuccess) return (int)e;
  r = layer_fwd(L, precision, xhi, xlo, B, H, W, sh, sw, y, st); if (r) return r;
  return (int)cudaStreamSynchronize(st);
}

int tc_conv_bwd_adhoc(int precision, const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
                      int B, int H, int W, int Cin, int kh, int kw, int Cout, int sh, int sw, cudaStream_t st, int w16) {
  TcLayer L{}; L.kh = kh; L.kw = kw; L.cin = Cin; L.cout = Cout; L.gated = 0;
  if (!layer_ok(L)) return TC_UNSUPPORTED;
  Temp T;
  int r = adhoc_layer(T, L, st); if (r) return r;
  if (precision == 3) { r = adhoc_layer_q(T, L, st); if (r) return r; }
  GatherGeom g = fwd_geom(B, H, W, kh, kw, sh, sw);
  size_t rows = (size_t)B * H * W, orows = (size_t)g.B * g.Hy * g.Wx;
  const int xpad = precision == 3 ? cin_q(L) : cin_k(L), gpad = precision == 3 ? nt_q(L) : nt_k(L);
  __nv_bfloat16* xhi = T.get<__nv_bfloat16>(rows * xpad); __nv_bfloat16* xlo = T.get<__nv_bfloat16>(rows * xpad);
  __nv_bfloat16* ghi = T.get<__nv_bfloat16>(orows * gpad); __nv_bfloat16* glo = T.get<__nv_bfloat16>(orows * gpad);
  float* zero = T.get<float>(Cout);
  if (!xhi || !xlo || !ghi || !glo || !zero) return (int)cudaErrorMemoryAllocation;
  cudaMemsetAsync(zero, 0, Cout * sizeof(float), st);
  r = refresh_layer(L, w, nullptr, zero, nullptr, st); if (r) return r;
  cudaError_t e = precision == 3 ? launch_pad_split_q(x, (long long)rows, Cin, Cin, xpad, xhi, xlo, st)
                                 : launch_pad_split(x, (long long)rows, Cin, Cin, xpad, xhi, xlo, st);
  if (e != cudaSuccess) return (int)e;
  e = precision == 3 ? launch_pad_split_q(dy, (long long)orows, Cout, Cout, gpad, ghi, glo, st)
                     : launch_pad_split(dy, (long long)orows, Cout, Cout, gpad, ghi, glo, st);
  if (e != cudaSuccess) return (int)e;
  if (dx) { r = layer_dgrad(L, precision, ghi, glo, B, H, W, sh, sw, dx, 0, st); if (r) return r; }
  if (dw) {
    r = layer_wgrad(L, precision, xhi, xlo, ghi, glo, B, H, W, sh, sw, dw, nullptr, st, w16); if (r) return r;
    if (dbias) { e = launch_colsum(dy, (long long)orows, Cout, 0, Cout, dbias, st); if (e != cudaSuccess) return (int)e; }
  }
  return (int)cudaStreamSynchronize(st);
}
