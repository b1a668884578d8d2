// Standalone probe (not part of libcgvc.so): does a TMA im2col load reproduce the gather of the gather-GEMM kernels for
// every convolution geometry of the CycleGAN-VC hot path (TF 'SAME' padding, strides 1 / 2, 1-D and 2-D, forward and the
// per-parity data-gradient forms), including tiles that straddle sample boundaries, the ragged last tile and a tile that
// lies entirely beyond the tensor?  Uses the library's own map builder (csrc/im2col_map.h).  Checked forms:
//   CTAS = 1   cp.async.bulk.tensor.4d...im2col                      one CTA per (tile, tap, channel block)
//   CTAS = 2   ...im2col.cta_group::2, both CTAs of a pair signal the leader's mbarrier (what the CTA-pair GEMM does)
// Every value of the bf16 source tensor is a distinct-ish non-zero integer, so zero fill and misplaced rows both show.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I voice-converter-cyclegan_b200/csrc -o tests/probes/bin/tma_im2col_probe tests/probes/tma_im2col_probe.cu
#include "geom.h"
#include "im2col_map.h"

#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int tries = 0; tries < (1 << 22); ++tries) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct ProbeParams {
  GatherGeom g; Im2colGeom ig;
  int pixels;              // rows per load (128: NT kernel, 64: TN kernel)
  int cblocks;             // 64-channel blocks
  long long M;
};

template <int CTAS>
__global__ void __launch_bounds__(128, 1) im2col_probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ ProbeParams p,
                                                        __nv_bfloat16* __restrict__ out, int* status) {
  __shared__ __align__(1024) uint8_t tile[128 * 128];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = CTAS == 1 ? 0u : cluster_rank();
  const int tile_idx = blockIdx.x;                         // CTAS == 2: consecutive CTAs of a pair own consecutive tiles
  const int tap = blockIdx.y, cb = blockIdx.z;
  const GatherGeom& g = p.g;
  const long long m0 = (long long)tile_idx * p.pixels;
  const int HW = g.Hy * g.Wx;
  const int b = (int)(m0 / HW); const int rem = (int)(m0 - (long long)b * HW);
  const int y = rem / g.Wx, x = rem - y * g.Wx;
  const uint32_t bytes = (uint32_t)p.pixels * 128u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CTAS == 2) cluster_sync();
  bool ok = true;
  if (threadIdx.x == 0) {
    const int cw = p.ig.lo_w + x * g.sx, ch = p.ig.lo_h + y * g.sy;
    const unsigned short ow = p.ig.off_w[tap], oh = p.ig.off_h[tap];
    if (CTAS == 1) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                   ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(&bar)), "r"(cb * 64), "r"(cw), "r"(ch), "r"(b), "h"(ow), "h"(oh) : "memory");
      ok = mbar_wait(&bar, 0);
    } else {
      if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(2u * bytes) : "memory");
      const uint32_t mbar = smem_u32(&bar) & 0xFEFFFFFFu;    // the leader's barrier counts both CTAs' bytes
      asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                   ::"r"(smem_u32(tile)), "l"(&tm), "r"(mbar), "r"(cb * 64), "r"(cw), "r"(ch), "r"(b), "h"(ow), "h"(oh) : "memory");
      if (rank == 0) ok = mbar_wait(&bar, 0);
    }
    if (!ok) atomicExch(status, -1);
  }
  __syncthreads();
  if (CTAS == 2) cluster_sync();                            // the leader has seen both loads complete
  // de-swizzle: row r, 16-byte chunk c lives at r*128 + ((c ^ (r & 7)) << 4)
  for (int i = threadIdx.x; i < p.pixels * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
    const long long orow = ((long long)(tap * p.cblocks + cb) * gridDim.x + tile_idx) * p.pixels + r;
    *reinterpret_cast<uint4*>(out + orow * 64 + c * 8) = v;
  }
}

static float srcval(int b, int y, int x, int c) { return (float)(((b * 131 + y * 31 + x * 7 + c * 3) % 255) + 1); }

static int run_case(const char* name, const GatherGeom& g, int C, int pixels, int ctas) {
  Im2colGeom ig = im2col_geom(g);
  if (!ig.ok) { printf("%-34s im2col_geom refused\n", name); return 1; }
  const size_t n_src = (size_t)g.B * g.Hs * g.Ws * C;
  std::vector<__nv_bfloat16> hsrc(n_src);
  for (int b = 0; b < g.B; ++b) for (int y = 0; y < g.Hs; ++y) for (int x = 0; x < g.Ws; ++x) for (int c = 0; c < C; ++c)
    hsrc[(((size_t)b * g.Hs + y) * g.Ws + x) * C + c] = __float2bfloat16(srcval(b, y, x, c));
  const long long M = (long long)g.B * g.Hy * g.Wx;
  int tiles = (int)((M + pixels - 1) / pixels);
  if (ctas == 2) tiles = (tiles + 1) / 2 * 2 + 2;            // pairs, plus one pair entirely beyond the tensor
  else tiles += 1;                                           // plus one tile entirely beyond the tensor
  const int cblocks = C / 64;
  const size_t n_out = (size_t)g.ntaps * cblocks * tiles * pixels * 64;
  __nv_bfloat16 *dsrc, *dout; int* dstat;
  cudaMalloc(&dsrc, n_src * 2); cudaMalloc(&dout, n_out * 2); cudaMalloc(&dstat, 4);
  cudaMemcpy(dsrc, hsrc.data(), n_src * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xFF, n_out * 2); cudaMemset(dstat, 0, 4);
  CUtensorMap tm;
  if (!make_im2col_map(&tm, dsrc, g, ig, C, C, pixels)) { printf("%-34s map encoding failed (lo %d,%d up %d,%d)\n", name, ig.lo_w, ig.lo_h, ig.up_w, ig.up_h); return 1; }
  ProbeParams p; p.g = g; p.ig = ig; p.pixels = pixels; p.cblocks = cblocks; p.M = M;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, g.ntaps, cblocks); cfg.blockDim = dim3(128);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t le = ctas == 1 ? cudaLaunchKernelEx(&cfg, im2col_probe<1>, tm, p, dout, dstat) : cudaLaunchKernelEx(&cfg, im2col_probe<2>, tm, p, dout, dstat);
  cudaError_t se = cudaDeviceSynchronize();
  if (le != cudaSuccess || se != cudaSuccess) { printf("%-34s CUDA error %s / %s\n", name, cudaGetErrorString(le), cudaGetErrorString(se)); exit(2); }
  int st = 0; cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
  std::vector<__nv_bfloat16> hout(n_out);
  cudaMemcpy(hout.data(), dout, n_out * 2, cudaMemcpyDeviceToHost);
  long long bad = 0, checked = 0; long long first_bad = -1;
  for (int t = 0; t < g.ntaps; ++t) for (int cb = 0; cb < cblocks; ++cb) for (long long m = 0; m < (long long)tiles * pixels; ++m) {
    bool inside = m < M;
    int b = 0, y = 0, x = 0, yy = 0, xx = 0;
    if (inside) {
      b = (int)(m / (g.Hy * g.Wx)); int rem = (int)(m % (g.Hy * g.Wx)); y = rem / g.Wx; x = rem % g.Wx;
      yy = y * g.sy + g.oy[t]; xx = x * g.sx + g.ox[t];
      inside = yy >= 0 && yy < g.Hs && xx >= 0 && xx < g.Ws;
    }
    for (int c = 0; c < 64; ++c) {
      const float want = inside ? srcval(b, yy, xx, cb * 64 + c) : 0.f;
      const float got = __bfloat162float(hout[(((size_t)(t * cblocks + cb) * tiles) * pixels + m) * 64 + c]);
      ++checked;
      if (got != want) { if (first_bad < 0) first_bad = m; ++bad; }
    }
  }
  printf("%-34s B%d %dx%d -> %dx%d s%d,%d taps %2d  lo(%d,%d) up(%d,%d) pix %3d ctas %d tiles %3d : %s (%lld / %lld wrong%s)\n", name, g.B, g.Hs, g.Ws, g.Hy, g.Wx,
         g.sy, g.sx, g.ntaps, ig.lo_w, ig.lo_h, ig.up_w, ig.up_h, pixels, ctas, tiles, (bad == 0 && st == 0) ? "OK" : "MISMATCH", bad, checked,
         st ? ", barrier wait timed out" : "");
  if (bad) printf("    first wrong row m = %lld\n", first_bad);
  cudaFree(dsrc); cudaFree(dout); cudaFree(dstat);
  return (bad || st) ? 1 : 0;
}

int main() {
  struct Case { const char* name; int B, H, W, kh, kw, sh, sw, C; };
  const Case cases[] = {
      {"G.d1  1x5 s2  W128", 3, 1, 128, 1, 5, 1, 2, 128},
      {"G.d2  1x5 s2  W64", 5, 1, 64, 1, 5, 1, 2, 64},
      {"G.res 1x3 s1  W32", 5, 1, 32, 1, 3, 1, 1, 128},
      {"G.u1  1x5 s1  W32", 3, 1, 32, 1, 5, 1, 1, 64},
      {"G.h1  1x15 s1 W128", 2, 1, 128, 1, 15, 1, 1, 64},
      {"G.res 1x3 s1  W129 (T=516)", 2, 1, 129, 1, 3, 1, 1, 64},
      {"G.d2  1x5 s2  W258 (T=516)", 2, 1, 258, 1, 5, 1, 2, 64},
      {"G.d1  1x5 s2  W1400", 1, 1, 1400, 1, 5, 1, 2, 64},
      {"D.d1  3x3 s2,2 24x64", 2, 24, 64, 3, 3, 2, 2, 128},
      {"D.d2  3x3 s2,2 12x32", 3, 12, 32, 3, 3, 2, 2, 64},
      {"D.d3  6x3 s1,2 6x16", 5, 6, 16, 6, 3, 1, 2, 64},
      {"D.d3  6x3 s1,2 6x88 (T=1408)", 2, 6, 88, 6, 3, 1, 2, 64},
  };
  int fails = 0, n = 0;
  for (const Case& c : cases) {
    GatherGeom f = fwd_geom(c.B, c.H, c.W, c.kh, c.kw, c.sh, c.sw);
    for (int ctas = 1; ctas <= 2; ++ctas) for (int pixels = 128; pixels >= 64; pixels -= 64) {
      char nm[96]; snprintf(nm, sizeof nm, "fwd   %s", c.name);
      fails += run_case(nm, f, c.C, pixels, ctas); ++n;
    }
    std::vector<GatherGeom> gs = dgrad_geoms(c.B, c.H, c.W, c.kh, c.kw, c.sh, c.sw);
    int k = 0;
    for (const GatherGeom& g : gs) {
      if (g.ntaps == 0) continue;
      char nm[96]; snprintf(nm, sizeof nm, "dgrad%d %s", k++, c.name);
      fails += run_case(nm, g, c.C, 128, 1 + (k & 1)); ++n;
    }
  }
  printf("%d of %d cases failed\n", fails, n);
  return fails ? 1 : 0;
}
