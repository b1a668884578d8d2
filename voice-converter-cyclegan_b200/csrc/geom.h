// Host-side geometry builders for the gather-GEMM primitive (TF 'SAME' padding).
#pragma once
#include "kernels.cuh"
#include <string.h>
#include <vector>

static inline void same_pad(int n, int k, int s, int& before, int& out) {
  out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n; if (total < 0) total = 0;
  before = total / 2;                       // the extra element of an odd total goes AFTER (TF convention)
}

static inline GatherGeom fwd_geom(int B, int H, int W, int kh, int kw, int sh, int sw) {
  GatherGeom g; memset(&g, 0, sizeof g);
  int ph, pw, Ho, Wo; same_pad(H, kh, sh, ph, Ho); same_pad(W, kw, sw, pw, Wo);
  g.B = B; g.Hy = Ho; g.Wx = Wo; g.Hs = H; g.Ws = W; g.sy = sh; g.sx = sw;
  g.Hd = Ho; g.Wd = Wo; g.dsy = 1; g.dsx = 1; g.doy = 0; g.dox = 0;
  g.ntaps = 0;
  for (int i = 0; i < kh; ++i) for (int j = 0; j < kw; ++j) {
    g.oy[g.ntaps] = (short)(i - ph); g.ox[g.ntaps] = (short)(j - pw); g.widx[g.ntaps] = (short)(i * kw + j); g.ntaps++;
  }
  return g;
}

static inline bool divisible(int v, int s) { return ((v % s) + s) % s == 0; }

// data-gradient geometries: one per output parity class (input position h = y*sh + py)
static inline std::vector<GatherGeom> dgrad_geoms(int B, int H, int W, int kh, int kw, int sh, int sw) {
  std::vector<GatherGeom> out;
  int ph, pw, Ho, Wo; same_pad(H, kh, sh, ph, Ho); same_pad(W, kw, sw, pw, Wo);
  for (int py = 0; py < sh; ++py) for (int px = 0; px < sw; ++px) {
    GatherGeom g; memset(&g, 0, sizeof g);
    g.B = B; g.Hy = (H - py + sh - 1) / sh; g.Wx = (W - px + sw - 1) / sw;
    if (g.Hy <= 0 || g.Wx <= 0) continue;
    g.Hs = Ho; g.Ws = Wo; g.sy = 1; g.sx = 1;
    g.Hd = H; g.Wd = W; g.dsy = sh; g.dsx = sw; g.doy = py; g.dox = px;
    g.ntaps = 0;
    for (int i = 0; i < kh; ++i) for (int j = 0; j < kw; ++j) {
      if (!divisible(py + ph - i, sh) || !divisible(px + pw - j, sw)) continue;
      g.oy[g.ntaps] = (short)((py + ph - i) / sh); g.ox[g.ntaps] = (short)((px + pw - j) / sw);
      g.widx[g.ntaps] = (short)(i * kw + j); g.ntaps++;
    }
    out.push_back(g);
  }
  return out;
}
