// Standalone probe (not part of libcgvc.so, not a test): issue rate of tcgen05.mma from ONE CTA per SM with both operands
// in shared memory (cta_group::1, SS form), for the MMA kinds a 2-unit split-precision scheme would mix:
//   kind::f16     bf16 x bf16  128 x N x 16
//   kind::f8f6f4  e4m3 x e4m3  128 x N x 32      (same bytes per instruction; 2x the FLOPs)
//   alternating f16 / f8f6f4 into the SAME fp32 TMEM accumulator
// Prints achieved dense TFLOP/s over all SMs and checks one accumulator value of each kind (operands are all ones).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/probes/bin/mma_rate_probe tests/probes/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: gives up after ~2^22 polls so that a faulting MMA cannot hang the GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int tries = 0; tries < (1 << 22); ++tries) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// mode 0: f16 only, 1: f8 only, 2: alternate (one f16 + two f8 per k-step: the 2-unit scheme's issue mix)
template <int N>
__global__ void __launch_bounds__(128, 1) probe(int mode, int iters, long long* cycles, float* sample) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  // A16: 128 rows x 128 B of bf16 1.0 (0x3F80); A8: 128 x 128 B of e4m3 1.0 (0x38); B likewise with N rows
  const uint32_t oA16 = 0, oB16 = 16384, oA8 = 16384 + N * 128, oB8 = oA8 + 16384, total = oB8 + N * 128;
  for (uint32_t i = threadIdx.x; i < total / 2; i += blockDim.x) {
    uint32_t byte = i * 2;
    bool is8 = byte >= oA8;
    ((uint16_t*)g)[i] = is8 ? 0x3838 : 0x3F80;
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t id16 = idesc(128, N, 1, 1), id8 = idesc(128, N, 0, 0);
    uint32_t phase = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {                         // one 64-wide (bf16) / 128-wide (fp8) K block = 4 k-steps of 32 bytes
        const uint64_t a16 = make_desc(base + oA16 + k * 32, 16, 1024), b16 = make_desc(base + oB16 + k * 32, 16, 1024);
        const uint64_t a8 = make_desc(base + oA8 + k * 32, 16, 1024), b8 = make_desc(base + oB8 + k * 32, 16, 1024);
        if (mode == 0) { mma_f16(tm, a16, b16, id16, (it | k) != 0); mma_f16(tm, a16, b16, id16, 1); mma_f16(tm, a16, b16, id16, 1); }
        else if (mode == 1) { mma_f8(tm, a8, b8, id8, (it | k) != 0); mma_f8(tm, a8, b8, id8, 1); mma_f8(tm, a8, b8, id8, 1); }
        else { mma_f16(tm, a16, b16, id16, (it | k) != 0); mma_f8(tm, a8, b8, id8, 1); mma_f8(tm, a8, b8, id8, 1); }
      }
      if ((it & 15) == 15 || it == iters - 1) { commit(&bar); if (!mbar_wait(&bar, phase)) { cycles[blockIdx.x] = -1; break; } phase ^= 1; }
    }
    t1 = clock64();
    if (cycles[blockIdx.x] != -1) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tm));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) *sample = __uint_as_float(v);
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N>
void run(int sms, int clock_khz) {
  const int smem = 120 * 1024;                          // > half of the SM's shared memory: exactly one CTA per SM
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* cyc; float* smp; cudaMalloc(&cyc, sms * sizeof(long long)); cudaMalloc(&smp, sizeof(float));
  const char* names[3] = {"f16 (bf16) only  ", "f8f6f4 (e4m3)    ", "1 f16 + 2 f8 mix "};
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 2048;
    cudaMemset(cyc, 0, sms * sizeof(long long));
    probe<N><<<sms, 128, smem>>>(mode, 64, cyc, smp);                     // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<N><<<sms, 128, smem>>>(mode, iters, cyc, smp);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("N=%d mode %d: CUDA error %s\n", N, mode, cudaGetErrorString(err)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    float s; cudaMemcpy(&s, smp, sizeof(float), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < sms; ++i) { if (h[i] < 0) { printf("N=%d mode %d: MMA completion never arrived (timeout)\n", N, mode); exit(2); } if (h[i] > mx) mx = h[i]; }
    // FLOPs: per iteration 12 MMAs; f16: 2*128*N*16 each, f8: 2*128*N*32 each
    double f16 = 2.0 * 128 * N * 16, f8 = 2.0 * 128 * N * 32;
    double per_it = mode == 0 ? 12 * f16 : mode == 1 ? 12 * f8 : 4 * f16 + 8 * f8;
    double flops = per_it * iters * sms;
    double expect = mode == 0 ? 12.0 * 16 * iters : mode == 1 ? 12.0 * 32 * iters : (4.0 * 16 + 8.0 * 32) * iters;
    printf("N=%3d %s: %8.3f ms  %7.1f TFLOP/s dense-equivalent  %6.1f clk per MMA (max over SMs)  D[0,0]=%.0f (expect %.0f) %s\n",
           N, names[mode], ms, flops / (ms * 1e-3) / 1e12, (double)mx / (12.0 * iters), s, expect, s == (float)expect ? "OK" : "MISMATCH");
  }
  cudaFree(cyc); cudaFree(smp);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<256>(p.multiProcessorCount, 0);
  run<128>(p.multiProcessorCount, 0);
  return 0;
}
