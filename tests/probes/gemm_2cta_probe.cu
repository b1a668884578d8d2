// Standalone probe (not part of libcgvc.so, not a test): does halving the weight-operand bytes each SM pulls from L2 lift a
// tcgen05 GEMM off the L2 -> SM operand wall (DESIGN.md sections 7, 10)?  Dense bf16 GEMM D[M,N] = A[M,K] * B[N,K]^T, both operands by
// TMA into SWIZZLE_128B tiles, one output tile per CTA (group), run twice:
//   CTAS = 1   128 x 256 tile per CTA, tcgen05.mma.cta_group::1            (0.0117 operand bytes per FLOP from L2)
//   CTAS = 2   256 x 256 tile per CTA PAIR (cluster of 2), cta_group::2: each CTA loads its 128 rows of A and HALF of the B tile,
//              the leader issues M = 256 MMAs that read both CTAs' shared memory                 (0.0078 bytes per FLOP)
// Prints TFLOP/s and checks sampled outputs against a CPU reference.  Every wait is bounded (a faulting MMA cannot hang the box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/probes/bin/gemm_2cta_probe tests/probes/gemm_2cta_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {          // K-major SWIZZLE_128B, SBO = 1024
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int tries = 0; tries < (1 << 24); ++tries) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CTAS>
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  if (CTAS == 1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
  } else {
    // the transaction bytes are signalled on the LEADER CTA's barrier: clear the peer bit of the barrier's shared-window address
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  if (CTAS == 1)
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
  else
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CTAS == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else {
    const uint16_t mask = 3;                              // arrive on this barrier in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
  }
}

constexpr int BK = 64;                                    // bf16 elements per stage = one 128-byte swizzle row
template <int CTAS> struct Cfg {
  static constexpr int A_BYTES = 128 * 128;               // 128 rows per CTA
  static constexpr int B_ROWS = 256 / CTAS;               // each CTA of a pair holds half of the 256-wide B tile
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = CTAS == 1 ? 4 : 6;
  static constexpr int SMEM = STAGES * STAGE + 1024;
};

template <int CTAS>
__global__ void __launch_bounds__(192, 1) gemm_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                      float* __restrict__ D, int M, int N, int K, int* status) {
  using C = Cfg<CTAS>;
  constexpr int S = C::STAGES;
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t full[S], empty[S], tmem_full;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 1 ? 0u : cluster_rank();
  const int group = blockIdx.x / CTAS;
  const int tiles_n = N / 256;
  const int m0 = (group / tiles_n) * 128 * CTAS + (int)rank * 128;     // this CTA's 128 rows of A / D
  const int n0 = (group % tiles_n) * 256;
  const int num_kb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tmem_full)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CTAS == 2) cluster_sync();                          // both CTAs' barriers are initialised before anyone signals the peer
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  bool ok = true;

  if (warp == 0 && lane == 0) {
    // ---- producer: this CTA's A rows and its share of the B tile; bytes are counted on the leader's full barrier
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb && ok; ++kb) {
      ok = mbar_wait(&empty[stage], phase ^ 1);
      const uint32_t sA = base + stage * C::STAGE, sB = sA + C::A_BYTES;
      if (rank == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"((uint32_t)(C::STAGE * CTAS)) : "memory");
      tma_load<CTAS>(sA, &tmA, kb * BK, m0, &full[stage]);
      tma_load<CTAS>(sB, &tmB, kb * BK, n0 + (int)rank * C::B_ROWS, &full[stage]);
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && rank == 0) {
    // ---- MMA issuer (leader CTA only): 128*CTAS x 256 x 16 per instruction
    constexpr uint32_t id = idesc_bf16(128 * CTAS, 256);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb && ok; ++kb) {
      ok = mbar_wait(&full[stage], phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0 && ok) {
        const uint32_t sA = base + stage * C::STAGE, sB = sA + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma<CTAS>(tm, make_desc(sA + k * 32), make_desc(sB + k * 32), id, (kb | k) != 0);
        commit<CTAS>(&empty[stage]);
        if (kb == num_kb - 1) commit<CTAS>(&tmem_full);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ---- epilogue: this CTA's 128 accumulator lanes -> D rows m0 .. m0 + 127
    const int q = warp & 3;
    ok = mbar_wait(&tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* drow = D + (size_t)(m0 + q * 32 + lane) * N + n0;
    // only the first 32 columns of every tile are written back: the probe measures the operand pipeline, not the epilogue
    for (int c = 0; c < 32 && ok; c += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(drow + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  if (!ok) atomicExch(status, -1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CTAS == 2) cluster_sync();                          // the peer may still be reading its accumulator half
  if (warp == 1) {
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
  }
}

// ---- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (EncodeTiledFn)p; }
  cuuint64_t dims[2] = {K, rows}; cuuint64_t strides[1] = {K * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
  return enc && enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }

template <int CTAS>
static int run(const __nv_bfloat16* dA, const __nv_bfloat16* dB, float* dD, int* dstat, int M, int N, int K,
               const std::vector<float>& hA, const std::vector<float>& hB) {
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, dA, M, K, 128) || !make_map(&tmB, dB, N, K, 256 / CTAS)) { printf("tensor map encoding failed\n"); return 1; }
  cudaFuncSetAttribute(gemm_probe<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CTAS>::SMEM);
  const int groups = (M / (128 * CTAS)) * (N / 256);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CTAS); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = Cfg<CTAS>::SMEM;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaMemset(dstat, 0, sizeof(int)); cudaMemset(dD, 0, (size_t)M * N * sizeof(float));
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_probe<CTAS>, tmA, tmB, dD, M, N, K, dstat);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (le != cudaSuccess || e != cudaSuccess) { printf("CTAS=%d: CUDA error %s / %s\n", CTAS, cudaGetErrorString(le), cudaGetErrorString(e)); return 2; }
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    int st = 0; cudaMemcpy(&st, dstat, sizeof(int), cudaMemcpyDeviceToHost);
    if (st != 0) { printf("CTAS=%d: a barrier wait timed out (status %d)\n", CTAS, st); return 3; }
  }
  // sampled check
  std::vector<float> row(N);
  int bad = 0; double maxrel = 0;
  const int rows_to_check[6] = {0, 1, 127, 128, 255, M - 1};
  for (int ri = 0; ri < 6; ++ri) {
    const int m = rows_to_check[ri];
    cudaMemcpy(row.data(), dD + (size_t)m * N, N * sizeof(float), cudaMemcpyDeviceToHost);
    for (int n = 0; n < N; n += 256 + 0) for (int j = 0; j < 32; j += 5) {
      const int nn = n + j;
      double s = 0; for (int k = 0; k < K; ++k) s += (double)hA[(size_t)m * K + k] * hB[(size_t)nn * K + k];
      // fp32 tensor-core accumulation over K / 16 partial sums: compare against the magnitude of the terms, not of the (possibly tiny) sum
      double d = fabs(row[nn] - s), rel = d / (fabs(s) + 0.05 * sqrt((double)K));
      if (rel > maxrel) maxrel = rel;
      if (rel > 1e-3) ++bad;
    }
  }
  const double tf = 2.0 * M * N * K / (best * 1e-3) / 1e12;
  printf("CTAS=%d  %d x %d x %d  tile %d x 256 per CTA%s: %.3f ms  %.1f TFLOP/s  operand stream %.2f TB/s  sampled check: %s (max err / term scale %.2e)\n",
         CTAS, M, N, K, 128 * CTAS, CTAS == 2 ? " pair" : "", best, tf, tf * (CTAS == 1 ? 0.01171875 : 0.0078125), bad ? "MISMATCH" : "OK", maxrel);
  return bad ? 4 : 0;
}

int main() {
  const int M = 8192, N = 8192, K = 8192;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K);
  std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = bf16_round(rnd()); bA[i] = __float2bfloat16(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = bf16_round(rnd()); bB[i] = __float2bfloat16(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; int* dstat;
  cudaMalloc(&dA, bA.size() * 2); cudaMalloc(&dB, bB.size() * 2); cudaMalloc(&dD, (size_t)M * N * sizeof(float)); cudaMalloc(&dstat, sizeof(int));
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  int r1 = run<1>(dA, dB, dD, dstat, M, N, K, hA, hB);
  if (r1 == 2) return 1;                                   // context lost
  int r2 = run<2>(dA, dB, dD, dstat, M, N, K, hA, hB);
  printf("exit codes: 1-CTA %d, 2-CTA %d\n", r1, r2);
  return 0;
}
