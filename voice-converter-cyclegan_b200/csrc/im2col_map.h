// TMA im2col tensor maps for the gather operand of the gather-GEMM kernels (host side).
//
// A GatherGeom row m = (b, y, x) reads source pixel (b, y*sy + oy[t], x*sx + ox[t]) of a channels-last tensor [B,Hs,Ws,C],
// zero outside (TF 'SAME' padding; module.py:22-64 of the reference).  That is exactly what a TMA im2col load walks: the
// "base pixel" of row m is (b, lo_h + y*sy, lo_w + x*sx) with lo = min over taps of the offset, the instruction's 16-bit
// offsets carry (oy[t] - lo_h, ox[t] - lo_w), and the bounding box [lo, lo + (n_out-1)*s] per dimension tells the unit where
// a row of base pixels ends, so one load of P pixels crosses row and sample boundaries the way the row enumeration m does
// and zero-fills the halo.  One instruction per (tap, 64-channel block, plane) replaces the per-thread cp.async gather.
#pragma once
#include "kernels.cuh"
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct Im2colGeom {
  int lo_w, lo_h, up_w, up_h;                               // pixelBoxLowerCorner / pixelBoxUpperCorner (W, H)
  unsigned short off_w[CGVC_MAX_TAPS], off_h[CGVC_MAX_TAPS]; // per-tap im2col offsets
  bool ok;
};

static inline Im2colGeom im2col_geom(const GatherGeom& g) {
  Im2colGeom r; r.ok = g.ntaps > 0;
  int lw = 1 << 30, lh = 1 << 30;
  for (int t = 0; t < g.ntaps; ++t) { if (g.ox[t] < lw) lw = g.ox[t]; if (g.oy[t] < lh) lh = g.oy[t]; }
  if (!r.ok) { lw = lh = 0; }
  r.lo_w = lw; r.lo_h = lh;
  // last base pixel of a row = lo + (n_out - 1) * stride = (n_src - 1) + upper
  r.up_w = lw + (g.Wx - 1) * g.sx + 1 - g.Ws;
  r.up_h = lh + (g.Hy - 1) * g.sy + 1 - g.Hs;
  for (int t = 0; t < g.ntaps; ++t) { r.off_w[t] = (unsigned short)(g.ox[t] - lw); r.off_h[t] = (unsigned short)(g.oy[t] - lh); }
  // rank-4 maps keep 8 bits per corner; the traversal stride field holds 1..8
  if (r.lo_w < -128 || r.lo_w > 127 || r.lo_h < -128 || r.lo_h > 127 || r.up_w < -128 || r.up_w > 127 || r.up_h < -128 || r.up_h > 127) r.ok = false;
  if (g.sx < 1 || g.sx > 8 || g.sy < 1 || g.sy > 8) r.ok = false;
  return r;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeIm2colFn get_im2col_encoder() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeIm2colFn)p;
  }
  return fn;
}

// Tensor [B][Hs][Ws][ld] of `esize`-byte elements (ld = row stride in elements; rows are 16-byte aligned), SWIZZLE_128B, one 128-byte
// line per pixel (64 two-byte or 128 one-byte channels) x `pixels` rows per load: the rows land as `pixels` consecutive 128-byte lines,
// swizzled exactly like a tiled [pixels][128 B] box.
static inline bool make_im2col_map(CUtensorMap* m, const void* base, const GatherGeom& g, const Im2colGeom& ig, int channels, int ld, int pixels,
                                   CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, int esize = 2) {
  EncodeIm2colFn enc = get_im2col_encoder();
  if (!enc || !ig.ok || pixels < 1 || pixels > 1024 || (esize != 1 && esize != 2)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)g.Ws, (cuuint64_t)g.Hs, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * esize, (cuuint64_t)ld * esize * g.Ws, (cuuint64_t)ld * esize * g.Ws * g.Hs};
  int lower[2] = {ig.lo_w, ig.lo_h}, upper[2] = {ig.up_w, ig.up_h};
  cuuint32_t es[4] = {1, (cuuint32_t)g.sx, (cuuint32_t)g.sy, 1};
  CUresult r = enc(m, dt, 4, const_cast<void*>(base), dims, strides, lower, upper, (cuuint32_t)(128 / esize), (cuuint32_t)pixels, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  // drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KB (one descriptor bit must be cleared; the
  // same correction the CUTLASS im2col descriptor builder applies)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010) {
    const unsigned long long bytes = (unsigned long long)strides[2] * (unsigned long long)g.B;
    if (bytes < 131072ull) reinterpret_cast<unsigned long long*>(m)[1] &= ~(1ull << 21);
  }
  return true;
}
