// Standalone correctness probe (not part of libcgvc.so, not a test) for the pieces a 2-unit split-precision scheme needs on
// sm_100a: does tcgen05.mma kind::f8f6f4 take the SAME shared-memory tile layouts the bf16 kernels use, and
// does `scale-input-d` do what the two-pass accumulation needs?  One CTA, one 128 x 256 tile, exact small-integer data,
// CPU reference in the same binary.
//   T1  e4m3 x e4m3, both operands K-major  SWIZZLE_128B (forward / data-gradient kernel layout), K = 128 (4 MMAs of K = 32)
//   T2  e4m3 x e4m3, both operands MN-major SWIZZLE_128B (weight-gradient kernel layout),       K = 64  (2 MMAs of K = 32)
//   T3  e5m2 (A) x e4m3 (B), K-major: the a_format field
//   T4  mixed kinds into one accumulator with scale-input-d:  D = bf16 MMA  +  2^-7 * (e4m3 MMA)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/probes/bin/mma_f8_layout_probe tests/probes/mma_f8_layout_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int TM = 128, TN = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int afmt, int bfmt, int amn, int bmn) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_scaled7(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {      // D = A*B + D * 2^-7
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, 1, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 7;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int tries = 0; tries < (1 << 22); ++tries) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// img: [A tile | B tile | A16 tile | B16 tile] byte images already in their shared-memory layouts; sizes in bytes
__global__ void __launch_bounds__(128, 1) probe(int test, const uint8_t* img, int bytes, float* out, int* status) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(img)[i];
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    if (test == 1 || test == 3) {              // K-major: A 128 rows x 128 B, B 256 rows x 128 B; k-step = 32 bytes
      const uint32_t id = idesc(TM, TN, test == 3 ? 1 : 0, 0, 0, 0);
      for (int k = 0; k < 4; ++k)
        mma_f8(tm, make_desc(base + k * 32, 16, 1024), make_desc(base + 16384 + k * 32, 16, 1024), id, k != 0);
    } else if (test == 2) {                    // MN-major: per 128-wide MN atom, 64 K-rows x 128 B = 8192 B (8-row groups 1024 B apart)
      const uint32_t id = idesc(TM, TN, 0, 0, 1, 1);
      for (int k = 0; k < 2; ++k)              // one MMA = 32 K-rows = 4 groups = 4096 B; B has two atoms 8192 B apart (LBO)
        mma_f8(tm, make_desc(base + k * 4096, 8192, 1024), make_desc(base + 8192 + k * 4096, 8192, 1024), id, k != 0);
    } else {                                   // test 4: fp8 part first (K = 128), then the bf16 part (K = 64) whose first MMA scales D by 2^-7
      const uint32_t id8 = idesc(TM, TN, 0, 0, 0, 0), id16 = idesc(TM, TN, 1, 1, 0, 0);
      for (int k = 0; k < 4; ++k)
        mma_f8(tm, make_desc(base + k * 32, 16, 1024), make_desc(base + 16384 + k * 32, 16, 1024), id8, k != 0);
      const uint32_t a16 = base + 16384 + 32768, b16 = a16 + 16384;
      mma_f16_scaled7(tm, make_desc(a16, 16, 1024), make_desc(b16, 16, 1024), id16);
      for (int k = 1; k < 4; ++k) mma_f16(tm, make_desc(a16 + k * 32, 16, 1024), make_desc(b16 + k * 32, 16, 1024), id16, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    if (!mbar_wait(&bar, 0)) *status = -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < TN; c += 32) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * TN + c + j] = __uint_as_float(v[j]);
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}

// ---- host: encoders and layouts
static uint8_t e4m3(int v) {            // exact for |v| <= 8
  if (v == 0) return 0;
  uint8_t s = v < 0 ? 0x80 : 0; int a = abs(v);
  int e = 0; while ((1 << (e + 1)) <= a) ++e;                   // a in [2^e, 2^(e+1))
  int m = ((a << 3) >> e) & 7;                                   // 3 mantissa bits (exact when a * 8 / 2^e is an integer)
  return s | (uint8_t)((e + 7) << 3) | (uint8_t)m;
}
static uint8_t e5m2(int v) {            // exact for |v| in {0,1,2,3,4,6,8}
  if (v == 0) return 0;
  uint8_t s = v < 0 ? 0x80 : 0; int a = abs(v);
  int e = 0; while ((1 << (e + 1)) <= a) ++e;
  int m = ((a << 2) >> e) & 3;
  return s | (uint8_t)((e + 15) << 2) | (uint8_t)m;
}
static uint16_t bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static inline size_t sw128(int row, int byte) { return (size_t)(row / 8) * 1024 + (row % 8) * 128 + ((((byte / 16) ^ (row % 8)) * 16) + byte % 16); }

static int Aval(int m, int k) { return ((k * 7 + m * 3) % 9) - 4; }
static int Bval(int n, int k) { return ((k * 5 + n) % 7) - 3; }
static int A5val(int m, int k) { static const int t[7] = {-4, -2, -1, 0, 1, 2, 3}; return t[(k * 3 + m) % 7]; }

int main() {
  uint8_t* dimg; float* dout; int* dstat;
  const int IMG = 16384 + 32768 + 16384 + 32768;
  cudaMalloc(&dimg, IMG); cudaMalloc(&dout, TM * TN * sizeof(float)); cudaMalloc(&dstat, sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 2048);
  std::vector<float> out(TM * TN), ref(TM * TN);
  int failures = 0;
  const int order[4] = {1, 3, 4, 2};          // the MN-major probe last: if the hardware rejects it, the context is lost
  for (int oi = 0; oi < 4; ++oi) {
    const int test = order[oi];
    std::vector<uint8_t> img(IMG, 0);
    std::fill(ref.begin(), ref.end(), 0.f);
    if (test == 1 || test == 3 || test == 4) {
      const int K = 128;
      for (int m = 0; m < TM; ++m) for (int k = 0; k < K; ++k) img[sw128(m, k)] = test == 3 ? e5m2(A5val(m, k)) : e4m3(Aval(m, k));
      for (int n = 0; n < TN; ++n) for (int k = 0; k < K; ++k) img[16384 + sw128(n, k)] = e4m3(Bval(n, k));
      for (int m = 0; m < TM; ++m) for (int n = 0; n < TN; ++n) {
        float s = 0; for (int k = 0; k < K; ++k) s += (float)((test == 3 ? A5val(m, k) : Aval(m, k)) * Bval(n, k));
        ref[m * TN + n] = s;
      }
      if (test == 4) {                         // bf16 part, K = 64: values v / 4 so that the two parts are distinguishable
        uint16_t* a16 = reinterpret_cast<uint16_t*>(img.data() + 16384 + 32768); uint16_t* b16 = reinterpret_cast<uint16_t*>(img.data() + 16384 + 32768 + 16384);
        for (int m = 0; m < TM; ++m) for (int k = 0; k < 64; ++k) a16[sw128(m, 2 * k) / 2] = bf16(0.25f * Aval(m, k + 1));
        for (int n = 0; n < TN; ++n) for (int k = 0; k < 64; ++k) b16[sw128(n, 2 * k) / 2] = bf16((float)Bval(n, k + 2));
        for (int m = 0; m < TM; ++m) for (int n = 0; n < TN; ++n) {
          float s = 0; for (int k = 0; k < 64; ++k) s += 0.25f * Aval(m, k + 1) * (float)Bval(n, k + 2);
          ref[m * TN + n] = s + ref[m * TN + n] / 128.f;
        }
      }
    } else {                                   // MN-major: element (k, mn) at atom (mn / 128): group (k / 8) * 1024 + (k % 8) * 128 + swizzled (mn % 128)
      const int K = 64;
      for (int k = 0; k < K; ++k) for (int m = 0; m < TM; ++m) img[sw128(k, m)] = e4m3(Aval(m, k));
      for (int k = 0; k < K; ++k) for (int n = 0; n < TN; ++n) img[8192 + (n / 128) * 8192 + sw128(k, n % 128)] = e4m3(Bval(n, k));
      for (int m = 0; m < TM; ++m) for (int n = 0; n < TN; ++n) {
        float s = 0; for (int k = 0; k < K; ++k) s += (float)(Aval(m, k) * Bval(n, k));
        ref[m * TN + n] = s;
      }
    }
    cudaMemcpy(dimg, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaMemset(dstat, 0, sizeof(int)); cudaMemset(dout, 0, TM * TN * sizeof(float));
    probe<<<1, 128, IMG + 2048>>>(test, dimg, IMG, dout, dstat);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T%d: CUDA error %s\n", test, cudaGetErrorString(e)); return 1; }
    int st = 0; cudaMemcpy(&st, dstat, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dout, TM * TN * sizeof(float), cudaMemcpyDeviceToHost);
    int bad = 0; double maxd = 0; int fm = -1, fn = -1;
    for (int i = 0; i < TM * TN; ++i) { double d = fabs((double)out[i] - ref[i]); if (d > 1e-3) { if (!bad) { fm = i / TN; fn = i % TN; } ++bad; } if (d > maxd) maxd = d; }
    const char* names[5] = {"", "e4m3 x e4m3, K-major SW128", "e4m3 x e4m3, MN-major SW128", "e5m2 x e4m3, K-major SW128", "bf16 MMA + 2^-7 * e4m3 MMA (scale-input-d)"};
    printf("T%d %-46s: %s  mismatches %d / %d  max |diff| %.4g  D[0,0]=%g (ref %g) D[5,77]=%g (ref %g)%s\n", test, names[test],
           (bad == 0 && st == 0) ? "OK  " : "FAIL", bad, TM * TN, maxd, out[0], ref[0], out[5 * TN + 77], ref[5 * TN + 77], st ? "  [MMA completion timed out]" : "");
    if (bad) printf("    first mismatch at (m=%d, n=%d): got %g, ref %g\n", fm, fn, out[fm * TN + fn], ref[fm * TN + fn]);
    failures += (bad != 0 || st != 0);
  }
  printf("%d of 4 probes failed\n", failures);
  return 0;
}
