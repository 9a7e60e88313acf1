// Training-side kernels: weighted cross-entropy (losses.py:13-27), backward passes of BN / LeakyReLU / bilinear
// resize / ConvLSTM cell, bias and weight gradients, Keras Adam (train2D.py:61,87-93).
// Data gradients of the convolutions run through the same table-driven implicit-GEMM kernel as the forward
// (transposed weight packing, LU_EPI_GRAD epilogue); see lu_train_host.inl.
#pragma once
#include "lu_defs.h"
#include "lu_conv.cuh"
#include "lu_elem.cuh"

struct TrainState {
  size_t off_loss_acc = 0;     // double[2]: sum(weighted ce), sum(valid)
};

// Keras (TF2 OptimizerV2) Adam, epsilon outside the bias-corrected sqrt (train2D.py:61,93)
struct LuAdam {
  float* p; const float* g; float* m; float* v; float lr_t, b1, b2, eps;
  LU_HD void operator()(int64_t i) const {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
};

// ---- weighted cross entropy over the raw logits (N,Hp,Wp,raw_cpad) ------------------------------------------------
// labels: (B,T,1,H,W) == (B,T,H,W,1) floats in {-1,0,1,2}; -1 = ignore (one_hot(-1) = 0 and valid = 0).
struct LuCeReduce {      // item = chunk of pixels of the un-padded frames
  const float* raw; const float* labels; double* acc;
  int64_t npix; int H, W, Hp, Wp, py0, px0, raw_cpad, chunk; float w0, w1, w2;
  LU_HD void operator()(int64_t it) const {
    int64_t p0 = it * chunk, p1 = p0 + chunk; if (p1 > npix) p1 = npix;
    float s = 0.f, nv = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      const int x = (int)(p % W); int64_t q = p / W; const int y = (int)(q % H); const int64_t f = q / H;
      const float lab = labels[p];
      if (!(lab > -1.f)) continue;
      nv += 1.f;
      const int li = (int)lab;
      const float* r = raw + ((f * Hp + y + py0) * Wp + x + px0) * (int64_t)raw_cpad;
      const float mx = fmaxf(r[0], fmaxf(r[1], r[2]));
      const float lse = mx + logf(expf(r[0] - mx) + expf(r[1] - mx) + expf(r[2] - mx));
      const float w = li == 0 ? w0 : (li == 1 ? w1 : w2);
      s += (lse - r[li < 0 ? 0 : (li > 2 ? 2 : li)]) * w;
    }
    lu_atomic_add(&acc[0], (double)s);
    lu_atomic_add(&acc[1], (double)nv);
  }
};
// loss = acc0 / (acc1 + 1e-5); dlogits -> bf16 planes buffer (N,Hp,Wp,planes*cpad), zero in the padding border
struct LuCeGrad {        // item = pixel of the PADDED frame
  const float* raw; const float* labels; const double* acc; float* loss_out; uint16_t* g;
  int H, W, Hp, Wp, py0, px0, raw_cpad, cpad, planes; float w0, w1, w2;
  float rank_scale;        // 1, or the number of ranks when acc[1] holds the valid-pixel count of ALL ranks (the mean
                           // over the ranks of loss and gradients is then the single-device value)
  LU_HD void operator()(int64_t p) const {
    const int xx = (int)(p % Wp); int64_t q = p / Wp; const int yy = (int)(q % Hp); const int64_t f = q / Hp;
    if (p == 0) loss_out[0] = (float)(acc[0] * (double)rank_scale / (acc[1] + 0.00001));
    float d[3] = {0.f, 0.f, 0.f};
    const int y = yy - py0, x = xx - px0;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const float lab = labels[(f * H + y) * W + x];
      if (lab > -1.f) {
        const int li = (int)lab;
        const float* r = raw + p * (int64_t)raw_cpad;
        const float mx = fmaxf(r[0], fmaxf(r[1], r[2]));
        const float e0 = expf(r[0] - mx), e1 = expf(r[1] - mx), e2 = expf(r[2] - mx);
        const float inv = 1.f / (e0 + e1 + e2);
        const float w = (li == 0 ? w0 : (li == 1 ? w1 : w2)) * rank_scale / (float)(acc[1] + 0.00001);
        d[0] = w * (e0 * inv - (li == 0 ? 1.f : 0.f));
        d[1] = w * (e1 * inv - (li == 1 ? 1.f : 0.f));
        d[2] = w * (e2 * inv - (li == 2 ? 1.f : 0.f));
      }
    }
    uint16_t* o = g + p * (int64_t)(cpad * planes);
    for (int c = 0; c < 3; ++c) {
      uint16_t hi, lo; lu_split(d[c], hi, lo);
      o[c] = hi;
      if (planes == 2) o[cpad + c] = lo;
    }
  }
};

struct LuCeLossOnly {
  const double* acc; float* loss_out; float rank_scale;
  LU_HD void operator()(int64_t) const { loss_out[0] = (float)(acc[0] * (double)rank_scale / (acc[1] + 0.00001)); }
};

LU_HDI float lu_ldplanes(const uint16_t* p, int cpad, int planes) {
  float v = lu_bf2f(p[0]);
  if (planes == 2) v += lu_bf2f(p[cpad]);
  return v;
}
LU_HDI void lu_stplanes(uint16_t* p, int cpad, int planes, float v) {
  uint16_t hi, lo; lu_split(v, hi, lo);
  p[0] = hi;
  if (planes == 2) p[cpad] = lo;
}

// 8 consecutive channels of a hi[/lo] planes buffer
LU_HDI void lu_ld8planes(const uint16_t* p, int cpad, int planes, float* v) {
  lu_load8_bf16(p, v);
  if (planes == 2) {
    float t[8];
    lu_load8_bf16(p + cpad, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t[j];
  }
}
LU_HDI void lu_st8planes(uint16_t* p, int cpad, int planes, const float* v) {
  uint16_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) lu_split(v[j], hi[j], lo[j]);
  lu_store8_bf16(p, hi);
  if (planes == 2) lu_store8_bf16(p + cpad, lo);
}
// ---- BatchNorm (training) + LeakyReLU backward ------------------------------------------------------------------
// (row-loop kernels, lu_elem.cuh: a thread keeps the per-channel constants of its 8 channels in registers)
// g = dA * lrelu'(bn_out); sums: [0:cpad) = sum g, [cpad:2cpad) = sum g * xhat
struct LuBnBwdReduce {
  const uint16_t* dA; const float* raw; const float* scale; const float* shift; const float* mean; const float* invstd;
  double* sums; int cpad, planes, raw_cpad, c_real; float alpha;
  struct State { float sc[8], sh[8], mu[8], is[8], s[8], sx[8]; };
  static constexpr int NSUM = 16;
  LU_HD void begin(int g, State& st) const {
    const int c = g * 8;
    lu_ld8f(scale + c, st.sc); lu_ld8f(shift + c, st.sh); lu_ld8f(mean + c, st.mu); lu_ld8f(invstd + c, st.is);
#pragma unroll
    for (int j = 0; j < 8; ++j) { st.s[j] = 0.f; st.sx[j] = 0.f; }
  }
  LU_HD void pixel(int64_t p, int g, State& st) const {
    const int c = g * 8;
    float r[8], d[8];
    lu_ld8f(raw + p * raw_cpad + c, r);
    lu_ld8planes(dA + p * (int64_t)(cpad * planes) + c, cpad, planes, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bn = r[j] * st.sc[j] + st.sh[j];
      const float gg = d[j] * (bn > 0.f ? 1.f : alpha);
      st.s[j] += gg; st.sx[j] += gg * (r[j] - st.mu[j]) * st.is[j];
    }
  }
  LU_HD void partials(const State& st, float* part) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) { part[j] = st.s[j]; part[8 + j] = st.sx[j]; }
  }
  LU_HD void flush(int g, const float* part) const {
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      if (c >= c_real) continue;
      lu_atomic_add(&sums[c], (double)part[j]);
      lu_atomic_add(&sums[raw_cpad + c], (double)part[8 + j]);
    }
  }
};
// dRaw = scale * (g - sum_g/n - xhat * sum_gx/n), written IN PLACE over dA; means = [mean g | mean g*xhat] per channel
struct LuBnBwdApply {
  uint16_t* dA; const float* raw; const float* scale; const float* shift; const float* mean; const float* invstd;
  const float* means; int cpad, planes, raw_cpad, c_real; float alpha;
  struct State { float sc[8], sh[8], mu[8], is[8], mg[8], mx[8]; };
  static constexpr int NSUM = 0;
  LU_HD void begin(int g, State& st) const {
    const int c = g * 8;
    if (c < raw_cpad) {
      lu_ld8f(scale + c, st.sc); lu_ld8f(shift + c, st.sh); lu_ld8f(mean + c, st.mu); lu_ld8f(invstd + c, st.is);
      lu_ld8f(means + c, st.mg); lu_ld8f(means + raw_cpad + c, st.mx);
    }
  }
  LU_HD void pixel(int64_t p, int g, State& st) const {
    const int c = g * 8;
    uint16_t* o = dA + p * (int64_t)(cpad * planes) + c;
    float gv[8], d[8], r[8];
    if (c < raw_cpad) {
      lu_ld8planes(o, cpad, planes, gv);
      lu_ld8f(raw + p * raw_cpad + c, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float bn = r[j] * st.sc[j] + st.sh[j];
        const float gg = gv[j] * (bn > 0.f ? 1.f : alpha);
        d[j] = (c + j < c_real) ? st.sc[j] * (gg - st.mg[j] - (r[j] - st.mu[j]) * st.is[j] * st.mx[j]) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = 0.f;
    }
    lu_st8planes(o, cpad, planes, d);
  }
  LU_HD void partials(const State&, float*) const {}
  LU_HD void flush(int, const float*) const {}
};
struct LuBnBwdParams {   // dgamma = sum g*xhat, dbeta = sum g, per-channel means for the apply pass; item = channel
  const double* sums; float* dgamma; float* dbeta; float* means; int64_t npix; int raw_cpad, c_real;
  int write_grads;         // 0: only the means (second call of the synchronised-BN path, on the sums of all ranks)
  LU_HD void operator()(int64_t c) const {
    if (c >= c_real) { means[c] = 0.f; means[raw_cpad + c] = 0.f; return; }
    if (write_grads) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[raw_cpad + c];
    }
    means[c] = (float)(sums[c] / (double)npix);
    means[raw_cpad + c] = (float)(sums[raw_cpad + c] / (double)npix);
  }
};

// ---- bilinear x2 backward (transpose of LuUpsample2x): dsrc (=|+=) sum of the up-sampled gradient ---------------
struct LuUpsample2xBwd {   // item = (n, iy, ix, group of 8 channels)
  const uint16_t* gup; uint16_t* gsrc; int h, w, cpad, planes, accumulate;
  // Input pixel j receives from the up-sampled positions 2j-1, 2j, 2j+1, 2j+2 with weights .25, .75, .75, .25; the edge
  // clamp of the forward (in[-1] := in[0], in[n] := in[n-1]) adds .25 to position 0 at j == 0 and to position 2n-1 at
  // j == n-1, and the outer positions do not exist there.  Four fixed taps: absent ones get weight 0 at a clamped index.
  LU_HD static void taps1d(int j, int n, int* idx, float* wt) {
    idx[0] = j >= 1 ? 2 * j - 1 : 0;              wt[0] = j >= 1 ? 0.25f : 0.f;
    idx[1] = 2 * j;                               wt[1] = j == 0 ? 1.0f : 0.75f;
    idx[2] = 2 * j + 1;                           wt[2] = j == n - 1 ? 1.0f : 0.75f;
    idx[3] = j <= n - 2 ? 2 * j + 2 : 2 * n - 1;  wt[3] = j <= n - 2 ? 0.25f : 0.f;
  }
  LU_HD void operator()(int64_t i) const {
    const int cg = cpad / 8;
    const int c = (int)(i % cg) * 8; int64_t p = i / cg;
    const int ix = (int)(p % w); p /= w; const int iy = (int)(p % h); const int64_t n = p / h;
    int yi[4], xi[4]; float yw[4], xw[4];
    taps1d(iy, h, yi, yw); taps1d(ix, w, xi, xw);
    const int ct = cpad * planes;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float t[8];
        lu_ld8planes(gup + (((n * 2 * h + yi[a]) * 2 * w) + xi[b]) * (int64_t)ct + c, cpad, planes, t);
        const float wgt = yw[a] * xw[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += wgt * t[j];
      }
    uint16_t* o = gsrc + ((n * h + iy) * w + ix) * (int64_t)ct + c;
    if (accumulate) {
      float t[8];
      lu_ld8planes(o, cpad, planes, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += t[j];
    }
    lu_st8planes(o, cpad, planes, s);
  }
};

// ---- ConvLSTM cell backward for one time step ----------------------------------------------------------------------
// (A row-loop variant that also accumulated the bias gradient in registers was measured in round 2: 160 registers, one
// resident block per SM, 16.6 instead of 9.3 ms per step -- the bias gradient is a separate row-loop pass over dZ instead.)
// item = (sample pixel, channel < fpad).  gates: (frames,H,W,planes*4*fpad) [i|f|g|o]; dZ same layout (bf16 planes).
struct LuLstmCellBwd {    // item = (sample pixel, group of 8 channels)
  const uint16_t* dH; const uint16_t* gates; const float* c_t; const float* c_prev; float* dC; uint16_t* dZ;
  int64_t pix_per_sample; int T, t, fpad, planes, gate_kind, c_prev_is_init, first;
  LU_HD void operator()(int64_t i) const {
    const int cg = fpad / 8;
    const int ch = (int)(i % cg) * 8; const int64_t sp = i / cg;            // sp = b*HW + pixel
    const int64_t b = sp / pix_per_sample, px = sp % pix_per_sample;
    const int64_t fp = (b * T + t) * pix_per_sample + px;                   // pixel index in frame-major buffers
    const int g4 = 4 * fpad;
    const uint16_t* gp = gates + fp * (int64_t)(g4 * planes) + ch;
    float gi[8], gf[8], gg[8], go[8], ct[8], cp[8], dh[8], dc[8], zi[8], zf[8], zg[8], zo[8];
    lu_ld8planes(gp, g4, planes, gi); lu_ld8planes(gp + fpad, g4, planes, gf);
    lu_ld8planes(gp + 2 * fpad, g4, planes, gg); lu_ld8planes(gp + 3 * fpad, g4, planes, go);
    lu_ld8f(c_t + fp * fpad + ch, ct);
    lu_ld8f(c_prev_is_init ? c_prev + sp * fpad + ch : c_prev + (fp - pix_per_sample) * fpad + ch, cp);
    lu_ld8planes(dH + fp * (int64_t)(fpad * planes) + ch, fpad, planes, dh);
    if (first) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dc[j] = 0.f;
    } else lu_ld8f(dC + sp * fpad + ch, dc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float th = tanhf(ct[j]);
      const float d = dc[j] + dh[j] * go[j] * (1.f - th * th);
      const float d_o = dh[j] * th, d_i = d * gg[j], d_g = d * gi[j], d_f = d * cp[j];
      dc[j] = d * gf[j];
      if (gate_kind == 0) {
        zi[j] = (gi[j] > 0.f && gi[j] < 1.f) ? 0.2f * d_i : 0.f;
        zf[j] = (gf[j] > 0.f && gf[j] < 1.f) ? 0.2f * d_f : 0.f;
        zo[j] = (go[j] > 0.f && go[j] < 1.f) ? 0.2f * d_o : 0.f;
      } else {
        zi[j] = d_i * gi[j] * (1.f - gi[j]); zf[j] = d_f * gf[j] * (1.f - gf[j]); zo[j] = d_o * go[j] * (1.f - go[j]);
      }
      zg[j] = d_g * (1.f - gg[j] * gg[j]);
    }
    lu_st8f(dC + sp * fpad + ch, dc);
    uint16_t* zp = dZ + fp * (int64_t)(g4 * planes) + ch;
    lu_st8planes(zp, g4, planes, zi); lu_st8planes(zp + fpad, g4, planes, zf);
    lu_st8planes(zp + 2 * fpad, g4, planes, zg); lu_st8planes(zp + 3 * fpad, g4, planes, zo);
  }
};

// ---- bias gradient: column sums of a gradient buffer (row-loop kernel: a thread owns 8 channels) ---------------------
struct LuColSumGrad {
  const uint16_t* g; float* dst; int cpad, planes;
  int c_real;        // identity layout: channel c < c_real -> dst[c]
  int gate_F, gate_fpad;   // gate layout (gate_F > 0): channel gate*fpad + ch -> dst[gate*F + ch]
  struct State { float s[8]; };
  static constexpr int NSUM = 8;
  LU_HD void begin(int, State& st) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) st.s[j] = 0.f;
  }
  LU_HD void pixel(int64_t p, int grp, State& st) const {
    float v[8];
    lu_ld8planes(g + p * (int64_t)(cpad * planes) + grp * 8, cpad, planes, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) st.s[j] += v[j];
  }
  LU_HD void partials(const State& st, float* part) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) part[j] = st.s[j];
  }
  LU_HD void flush(int grp, const float* part) const {
    for (int j = 0; j < 8; ++j) {
      const int c = grp * 8 + j;
      int d;
      if (gate_F > 0) { const int gt = c / gate_fpad, ch = c % gate_fpad; if (ch >= gate_F) continue; d = gt * gate_F + ch; }
      else { if (c >= c_real) continue; d = c; }
      lu_atomic_add(&dst[d], part[j]);
    }
  }
};

// ---- weight gradient in PACKED space (scalar engine): dWp[n][k] = sum_pixels A_k[pixel] * dY[pixel][n] --------------
// A_k is exactly the forward operand of K block k/64, channel k%64 (same staging tables), so the mapping back to the
// Keras tensor is the weight-packing descriptor read backwards (LuUnpackWgrad).
struct LuWgradMirror {
  LuConvParams p;                 // forward tables / views (frame mapping set by the launcher)
  const uint16_t* kb_stage;       // K block -> stage index
  const uint16_t* kb_tap;         // K block -> tap offset inside the stage's window
  const uint16_t* dY; int dy_cpad, dy_planes; int64_t dy_frame_mul, dy_frame_add;
  LuColMap cm; int gate_fpad;     // packed column -> dY channel (identity, or lstm gate layout)
  float* dWp; int npad, H, W, frames, chunk;
  int only_src;                   // >= 0: only K blocks of this source view
  int skip_t0_src, T;             // source whose contribution is skipped for frames with t == 0 (h_{t-1} of step 0)
  LU_HD void operator()(int64_t it) const {
    const int64_t npix = (int64_t)frames * H * W;
    const int64_t nchunk = (npix + chunk - 1) / chunk;
    const int64_t pc = it % nchunk; int64_t r = it / nchunk;
    const int n16 = (int)(r % (npad / 16)); r /= (npad / 16);
    const int k = (int)r, kb = k / LU_KBLK, kk = k % LU_KBLK;
    const LuAStage st = p.astages[kb_stage[kb]];
    if (only_src >= 0 && st.src != only_src) return;
    const LuSrcView& v = p.src[st.src];
    if (st.c + kk >= v.dimC) return;
    const int off = kb_tap[kb];
    int dch[16]; bool any = false;
    for (int j = 0; j < 16; ++j) {
      const int n = n16 * 16 + j;
      int d = -1;
      if (cm.kind == LU_COL_IDENTITY) d = n < cm.n_real ? n : -1;
      else { const int bn = 4 * cm.ch_tile, tile = n / bn, rr = n % bn, g = rr / cm.ch_tile, jj = rr % cm.ch_tile;
             const int ch = tile * cm.ch_tile + jj; d = ch < cm.F ? g * gate_fpad + ch : -1; }
      dch[j] = d; any = any || d >= 0;
    }
    if (!any) return;
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    int64_t p0 = pc * chunk, p1 = p0 + chunk; if (p1 > npix) p1 = npix;
    for (int64_t px = p0; px < p1; ++px) {
      const int x = (int)(px % W); int64_t q = px / W; const int y = (int)(q % H); const int64_t f = q / H;
      if (st.src == skip_t0_src && (f % T) == 0) continue;
      const int yy = y + st.dy + off / v.pitch, xx = x + st.dx + off % v.pitch;
      if (yy < 0 || yy >= v.dimH || xx < 0 || xx >= v.dimW) continue;
      const int64_t n = f * v.frame_mul + v.frame_add;
      const float a = lu_bf2f(v.ptr[n * v.sn + (int64_t)yy * v.sh + (int64_t)st.plane * v.sp + (int64_t)xx * v.sw + st.c + kk]);
      if (a == 0.f) continue;
      const uint16_t* dy = dY + (((f * dy_frame_mul + dy_frame_add) * H + y) * W + x) * (int64_t)(dy_cpad * dy_planes);
      for (int j = 0; j < 16; ++j)
        if (dch[j] >= 0) acc[j] += a * lu_ldplanes(dy + dch[j], dy_cpad, dy_planes);
    }
    for (int j = 0; j < 16; ++j)
      if (dch[j] >= 0 && acc[j] != 0.f) lu_atomic_add(&dWp[(int64_t)(n16 * 16 + j) * p.ktot + k], acc[j]);
  }
};
// scatter-add the packed gradient into the Keras-layout gradient tensor; item = (n, k).  Blocks that pair the
// activation with the LO part of the weight (bf16x3) are duplicates of the hi block and are skipped.
struct LuUnpackWgrad {
  const float* dWp; const LuPackDesc* descs; float* grads; LuColMap cm; int ktot;
  LU_HD void operator()(int64_t i) const {
    const int k = (int)(i % ktot); const int n = (int)(i / ktot);
    const float g = dWp[i];
    if (g == 0.f) return;
    const int kb = k / LU_KBLK, kk = k % LU_KBLK;
    const int col = lu_col_of(cm, n);
    if (col < 0) return;
    const LuPackDesc d = descs[kb];
    if (d.wpart != 0) return;
    if (d.kind == 0) {
      if (kk < d.n_valid) lu_atomic_add(&grads[d.w_off + d.tap_off + (int64_t)(d.c_base + kk) * d.cout_total + col], g);
    } else {
      int t = kk;
      if (d.patch_x3 && kk >= 32) t = kk - 32;
      if (t < d.pw * d.pw) {
        const int sh = (d.pw - 1) / 2 - (d.k - 1) / 2;
        const int ky = t / d.pw - sh, kx = t % d.pw - sh;
        if (ky >= 0 && ky < d.k && kx >= 0 && kx < d.k)
          lu_atomic_add(&grads[d.w_off + ((int64_t)(ky * d.k + kx) * d.cin_total + d.c_base) * d.cout_total + col], g);
      }
    }
  }
};
