// Instance labelling of the soft-max maps on the device: the step that follows the model call in the reference's
// Inference2D.inference (Inference2D.py:64-123).  Integer / byte work, HBM- and latency-bound; every
// result is bit-identical to the reference's numpy + SciPy + OpenCV code (oracle/postprocess_oracle.py states the
// library behaviours reproduced here: OpenCV's label numbering, SciPy's hole definition and nearest-feature ties).
//
//   classify           edge = p2 >= thr ; cell = argmax == 1 and not edge                    (:66-69)
//   fill holes         union-find over the 4-connected background, components that do not touch the frame border are
//                      cell                                                                     (:70-71)
//   components         union-find over the 8-connected cell mask; numbered like OpenCV's block-based scan: by the
//                      raster position of the first 2x2 block of the component; areas          (:72-75)
//   edge assignment    every edge pixel closer than edge_dist to a cell takes the nearest cell's label; equidistant
//                      cells: smallest column, then smallest row (SciPy's feature transform)   (:77-78)
//   per-label holes    one CTA per label floods the complement of the label inside its bounding box; enclosed pixels
//                      get += n.  The reference does this sequentially on the running label image; the parallel form
//                      is identical unless an enclosed pixel already carries another label -- detected on the device,
//                      and then a one-CTA-per-frame pass redoes the loop in the reference's order (:80-91)
//   field of view, size filter on the core area, consecutive renumbering, uint16 output        (:94-124)
//
// Pixel kernels are index functors (lu_parallel_for_impl) and the CTA kernels are written as phase-separated strided
// loops, so the TEST-ONLY host build (LU_HOST_EMU) runs the same code with one "thread" per CTA.
#pragma once
#include "lu_elem.cuh"

#define LU_PP_SEG 32             // pixels of a row handled by one item of the run-linking kernels
#define LU_PP_EMPTY 0x7f7f7f7f   // memset(0x7f) pattern: "no pixel yet" in the bounding boxes / block keys
#define LU_PP_SMEM_CROP 24576    // crops up to this many pixels are flooded in shared memory
#define LU_PP_CTA 128

enum { LU_PP_CELL0 = 1, LU_PP_EDGE0 = 2, LU_PP_CELL = 4, LU_PP_EDGE = 8 };

struct LuPost {
  const float* sm;
  int chw;                            // 1: (N,3,H,W), 0: (N,H,W,3)
  int N, H, W, HW, WB, NB, KMAX, G;   // WB x HB 2x2 blocks; KMAX = NB + 1 labels at most; G flood CTAs per frame
  float edge_thresh;
  int d2lim, rad, min_size, max_size, fov;
  uint8_t* cls;                       // [N][HW]
  int32_t *parA, *parB, *key, *area, *cc, *lab, *add;   // [N][HW]
  int32_t* bflag;                     // [N][NB]     flag, then inclusive rank
  int32_t *larea, *present, *newlab;  // [N][KMAX]
  int32_t* bbox;                      // [N][KMAX][4] = rmin, -rmax, cmin, -cmax
  int32_t* info;                      // [N][4] = label count (incl. background), kept, sequential-redo flag, 0
  uint8_t* slab;                      // [N][G][HW] flood scratch for crops larger than LU_PP_SMEM_CROP
  uint16_t* out;                      // [N][HW]
};

LU_HDI int lu_atomic_min_i(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
  return atomicMin(p, v);
#else
  const int32_t o = *p; if (v < o) *p = v; return o;
#endif
}
LU_HDI int lu_atomic_add_i(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  const int32_t o = *p; *p = o + v; return o;
#endif
}
LU_HDI int32_t lu_ld_volatile(const int32_t* p) {
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const volatile int32_t*>(p);
#else
  return *p;
#endif
}

// ---- lock-free union-find (parents only ever decrease; the root of a set is its smallest pixel index) -------------
LU_HDI int lu_uf_find(const int32_t* par, int a) {
  for (;;) { const int b = lu_ld_volatile(par + a); if (b == a) return a; a = b; }
}
LU_HDI void lu_uf_union(int32_t* par, int a, int b) {
  for (;;) {
    a = lu_uf_find(par, a); b = lu_uf_find(par, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = lu_atomic_min_i(par + a, b);       // a was a root: done; otherwise keep merging its old parent
    if (old == a) return;
    a = old;
  }
}

// np.argmax(softmax, 0) == 1 and not edge  (first maximum wins; NaN counts as the maximum, like numpy)
LU_HDI int lu_pp_classify(float s0, float s1, float s2, float thr) {
  const int edge = s2 >= thr;
  int am;
  if (s0 != s0) am = 0; else if (s1 != s1) am = 1; else if (s2 != s2) am = 2;
  else am = (s1 > s0) ? ((s2 > s1) ? 2 : 1) : ((s2 > s0) ? 2 : 0);
  return ((am == 1 && !edge) ? LU_PP_CELL0 : 0) | (edge ? LU_PP_EDGE0 : 0);
}

// item = (frame, row, 32-pixel segment): classify, link background runs (each pixel -> first pixel of its run inside the
// segment; a run continuing from the previous segment links to the pixel on its left)
struct LuPpClassify {
  LuPost q;
  LU_HD int cls_at(int64_t n, int p) const {
    float s0, s1, s2;
    if (q.chw) { const float* b = q.sm + n * 3 * (int64_t)q.HW + p; s0 = b[0]; s1 = b[q.HW]; s2 = b[2 * (int64_t)q.HW]; }
    else { const float* b = q.sm + (n * (int64_t)q.HW + p) * 3; s0 = b[0]; s1 = b[1]; s2 = b[2]; }
    return lu_pp_classify(s0, s1, s2, q.edge_thresh);
  }
  LU_HD void operator()(int64_t i) const {
    const int nseg = (q.W + LU_PP_SEG - 1) / LU_PP_SEG;
    const int seg = (int)(i % nseg); int64_t r = i / nseg;
    const int y = (int)(r % q.H); const int64_t n = r / q.H;
    const int x0 = seg * LU_PP_SEG, x1 = x0 + LU_PP_SEG < q.W ? x0 + LU_PP_SEG : q.W;
    const int64_t fo = n * (int64_t)q.HW;
    bool prev = x0 > 0 && !(cls_at(n, y * q.W + x0 - 1) & LU_PP_CELL0);
    int start = y * q.W + x0 - 1;
    for (int x = x0; x < x1; ++x) {
      const int p = y * q.W + x;
      const int c = cls_at(n, p);
      q.cls[fo + p] = (uint8_t)c;
      const bool bg = !(c & LU_PP_CELL0);
      int link = p;
      if (bg && prev) link = start; else start = p;
      q.parA[fo + p] = link;
      q.cc[fo + p] = 0;                 // "background component touches the frame border" flags, per root
      prev = bg;
    }
  }
};

// item = pixel: vertical links of the 4-connected background.  A pixel whose left and upper-left neighbours are
// background too leaves the link to its left neighbour (same runs).
struct LuPpMergeBg {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    const int y = p / q.W, x = p % q.W;
    const uint8_t* c = q.cls + fo;
    if (y == 0 || (c[p] & LU_PP_CELL0) || (c[p - q.W] & LU_PP_CELL0)) return;
    if (x > 0 && !(c[p - 1] & LU_PP_CELL0) && !(c[p - q.W - 1] & LU_PP_CELL0)) return;
    lu_uf_union(q.parA + fo, p, p - q.W);
  }
};
struct LuPpFlattenBg {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    if (q.cls[fo + p] & LU_PP_CELL0) return;
    const int r = lu_uf_find(q.parA + fo, p);
    q.parA[fo + p] = r;
    const int y = p / q.W, x = p % q.W;
    if (y == 0 || y == q.H - 1 || x == 0 || x == q.W - 1) q.cc[fo + r] = 1;
  }
};

// item = (frame, row, segment): filled cell mask (cell, or background component that never reaches the border), edge
// pixels that are not cell, run links of the cell mask, per-root accumulators reset
struct LuPpFill {
  LuPost q;
  LU_HD bool filled(int64_t fo, int p) const {
    if (q.cls[fo + p] & LU_PP_CELL0) return true;
    return q.cc[fo + q.parA[fo + p]] == 0;
  }
  LU_HD void operator()(int64_t i) const {
    const int nseg = (q.W + LU_PP_SEG - 1) / LU_PP_SEG;
    const int seg = (int)(i % nseg); int64_t r = i / nseg;
    const int y = (int)(r % q.H); const int64_t n = r / q.H;
    const int x0 = seg * LU_PP_SEG, x1 = x0 + LU_PP_SEG < q.W ? x0 + LU_PP_SEG : q.W;
    const int64_t fo = n * (int64_t)q.HW;
    bool prev = x0 > 0 && filled(fo, y * q.W + x0 - 1);
    int start = y * q.W + x0 - 1;
    for (int x = x0; x < x1; ++x) {
      const int p = y * q.W + x;
      const bool f = filled(fo, p);
      int c = q.cls[fo + p] & (LU_PP_CELL0 | LU_PP_EDGE0);
      if (f) c |= LU_PP_CELL; else if (c & LU_PP_EDGE0) c |= LU_PP_EDGE;
      q.cls[fo + p] = (uint8_t)c;
      int link = p;
      if (f && prev) link = start; else start = p;
      q.parB[fo + p] = link;
      q.key[fo + p] = LU_PP_EMPTY;
      q.area[fo + p] = 0;
      prev = f;
    }
  }
};

// item = pixel: links of the 8-connected cell mask to the row above.  With L, UL, U, UR the left / upper neighbours:
// U if not L;  UR if not U;  UL if neither U nor L -- every pair of adjacent runs is joined at the leftmost contact.
struct LuPpMergeFg {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    const int y = p / q.W, x = p % q.W;
    const uint8_t* c = q.cls + fo;
    if (y == 0 || !(c[p] & LU_PP_CELL)) return;
    const bool L = x > 0 && (c[p - 1] & LU_PP_CELL);
    const bool U = (c[p - q.W] & LU_PP_CELL) != 0;
    const bool UL = x > 0 && (c[p - q.W - 1] & LU_PP_CELL);
    const bool UR = x + 1 < q.W && (c[p - q.W + 1] & LU_PP_CELL);
    if (U && !L) lu_uf_union(q.parB + fo, p, p - q.W);
    if (UR && !U) lu_uf_union(q.parB + fo, p, p - q.W + 1);
    if (UL && !U && !L) lu_uf_union(q.parB + fo, p, p - q.W - 1);
  }
};
struct LuPpFlattenFg {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    if (!(q.cls[fo + p] & LU_PP_CELL)) return;
    const int r = lu_uf_find(q.parB + fo, p);
    q.parB[fo + p] = r;
    const int y = p / q.W, x = p % q.W;
    lu_atomic_min_i(q.key + fo + r, (y >> 1) * q.WB + (x >> 1));
    lu_atomic_add_i(q.area + fo + r, 1);
  }
};
struct LuPpMarkBlocks {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    if ((q.cls[i] & LU_PP_CELL) && q.parB[i] == p) q.bflag[n * q.NB + q.key[i]] = 1;
  }
};

// item = pixel: OpenCV label of every cell pixel, core areas per label
struct LuPpAssign {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW; const int64_t fo = i - p;
    int l = 0;
    if (q.cls[i] & LU_PP_CELL) {
      const int r = q.parB[i];
      l = q.bflag[n * q.NB + q.key[fo + r]];
      if (r == p) q.larea[n * q.KMAX + l] = q.area[i];
    }
    q.cc[i] = l;
    q.add[i] = 0;
  }
};

// item = pixel: edge pixels join the nearest cell (squared distance < d2lim; ties: smallest column, then row);
// bounding boxes of the labels
struct LuPpEdges {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW; const int64_t fo = i - p;
    int l = q.cc[i];
    const int y = p / q.W, x = p % q.W;
    if (q.cls[i] & LU_PP_EDGE) {
      int bd = q.d2lim, bx = 0, by = 0; bool found = false;
      for (int dx = -q.rad; dx <= q.rad; ++dx) {            // columns ascending, rows ascending: strict < keeps ties
        const int xx = x + dx;
        if (xx < 0 || xx >= q.W) continue;
        for (int dy = -q.rad; dy <= q.rad; ++dy) {
          const int yy = y + dy, d2 = dx * dx + dy * dy;
          if (yy < 0 || yy >= q.H || d2 >= bd) continue;
          if (q.cls[fo + yy * q.W + xx] & LU_PP_CELL) { bd = d2; bx = xx; by = yy; found = true; }
        }
      }
      if (found) l = q.cc[fo + by * q.W + bx];
    }
    q.lab[i] = l;
    if (l > 0) {
      int32_t* b = q.bbox + (n * q.KMAX + l) * 4;
      lu_atomic_min_i(b + 0, y); lu_atomic_min_i(b + 1, -y);
      lu_atomic_min_i(b + 2, x); lu_atomic_min_i(b + 3, -x);
    }
  }
};

// ---- CTA-level routines ---------------------------------------------------------------------------------------------
struct LuCta { int tid, nthreads; };
#if defined(__CUDA_ARCH__)
#define LU_CTA_SYNC() __syncthreads()
#else
#define LU_CTA_SYNC() do { } while (0)
#endif

// inclusive scan of v[0..n) in place (each thread owns a contiguous chunk; sums[] has nthreads + 1 entries); returns the total
LU_HDI int lu_cta_scan(const LuCta& t, int32_t* v, int n, int32_t* sums) {
  const int chunk = (n + t.nthreads - 1) / t.nthreads;
  const int a = t.tid * chunk < n ? t.tid * chunk : n, b = a + chunk < n ? a + chunk : n;
  int s = 0;
  for (int i = a; i < b; ++i) s += v[i];
  sums[t.tid + 1] = s;
  LU_CTA_SYNC();
  if (t.tid == 0) { sums[0] = 0; for (int i = 0; i < t.nthreads; ++i) sums[i + 1] += sums[i]; }
  LU_CTA_SYNC();
  s = sums[t.tid];
  for (int i = a; i < b; ++i) { s += v[i]; v[i] = s; }
  const int total = sums[t.nthreads];
  LU_CTA_SYNC();
  return total;
}

// Holes of the mask {lab == n} inside its bounding box grown by one pixel: state 0 = mask, 1 = not reached, 2 = reached
// from the crop boundary through 4-connected non-mask pixels.  Alternating row / column sweeps until nothing changes.
// Afterwards every pixel still in state 1 is a hole; `apply(p)` is called for each.  Returns false if there is no mask.
template <class Apply>
LU_HDI void lu_cta_flood_label(const LuCta& t, const LuPost& q, int64_t n_frame, int n, const int32_t* lab,
                               uint8_t* small, uint8_t* big, int32_t* sh_changed, Apply apply) {
  const int32_t* bb = q.bbox + (n_frame * q.KMAX + n) * 4;
  const int rmin = lu_ld_volatile(bb + 0), rmax = -lu_ld_volatile(bb + 1);
  const int cmin = lu_ld_volatile(bb + 2), cmax = -lu_ld_volatile(bb + 3);
  if (rmin == LU_PP_EMPTY) return;
  if (rmax - rmin < 2 || cmax - cmin < 2) return;               // nothing can be enclosed
  const int r0 = rmin > 0 ? rmin - 1 : 0, r1 = rmax + 1 < q.H ? rmax + 1 : q.H - 1;
  const int c0 = cmin > 0 ? cmin - 1 : 0, c1 = cmax + 1 < q.W ? cmax + 1 : q.W - 1;
  const int h = r1 - r0 + 1, w = c1 - c0 + 1;
  uint8_t* s = (h * w <= LU_PP_SMEM_CROP) ? small : big;
  LU_CTA_SYNC();                                                // previous label's use of the buffers is over
  for (int i = t.tid; i < h * w; i += t.nthreads) {
    const int yy = i / w, xx = i % w;
    const bool m = lab[(r0 + yy) * q.W + c0 + xx] == n;
    const bool edge = yy == 0 || yy == h - 1 || xx == 0 || xx == w - 1;
    s[i] = m ? 0 : (edge ? 2 : 1);
  }
  for (;;) {
    LU_CTA_SYNC();
    if (t.tid == 0) *sh_changed = 0;
    LU_CTA_SYNC();
    int ch = 0;
    for (int yy = 1 + t.tid; yy < h - 1; yy += t.nthreads) {
      uint8_t* row = s + yy * w;
      for (int xx = 1; xx < w - 1; ++xx) if (row[xx] == 1 && row[xx - 1] == 2) { row[xx] = 2; ch = 1; }
      for (int xx = w - 2; xx >= 1; --xx) if (row[xx] == 1 && row[xx + 1] == 2) { row[xx] = 2; ch = 1; }
    }
    LU_CTA_SYNC();
    for (int xx = 1 + t.tid; xx < w - 1; xx += t.nthreads) {
      for (int yy = 1; yy < h - 1; ++yy) if (s[yy * w + xx] == 1 && s[(yy - 1) * w + xx] == 2) { s[yy * w + xx] = 2; ch = 1; }
      for (int yy = h - 2; yy >= 1; --yy) if (s[yy * w + xx] == 1 && s[(yy + 1) * w + xx] == 2) { s[yy * w + xx] = 2; ch = 1; }
    }
    if (ch) *sh_changed = 1;
    LU_CTA_SYNC();
    if (!*sh_changed) break;
  }
  for (int i = t.tid; i < h * w; i += t.nthreads)
    if (s[i] == 1) apply((r0 + i / w) * q.W + c0 + i % w);
}

// one CTA per (flood slot g, frame): labels g+1, g+1+G, ... ; enclosed pixels accumulate n in add[]; an enclosed pixel
// that already carries a label raises the frame's sequential-redo flag
LU_HDI void lu_pp_holes_cta(const LuCta& t, const LuPost& q, int g, int64_t n_frame, uint8_t* small, int32_t* sh_changed) {
  const int K = q.info[n_frame * 4 + 0];
  const int64_t fo = n_frame * (int64_t)q.HW;
  const int32_t* lab = q.lab + fo;
  int32_t* add = q.add + fo;
  int32_t* flag = q.info + n_frame * 4 + 2;
  uint8_t* big = q.slab + (n_frame * q.G + g) * (int64_t)q.HW;
  for (int n = 1 + g; n < K; n += q.G)
    lu_cta_flood_label(t, q, n_frame, n, lab, small, big, sh_changed, [&](int p) {
      lu_atomic_add_i(add + p, n);
      if (lab[p] != 0) *flag = 1;
    });
}
// the reference's order, on the running label image (only when the flag is up): one CTA per frame
LU_HDI void lu_pp_holes_seq_cta(const LuCta& t, const LuPost& q, int64_t n_frame, uint8_t* small, int32_t* sh_changed) {
  if (q.info[n_frame * 4 + 2] == 0) return;
  const int K = q.info[n_frame * 4 + 0];
  const int64_t fo = n_frame * (int64_t)q.HW;
  int32_t* lab = q.lab + fo;
  uint8_t* big = q.slab + (n_frame * q.G) * (int64_t)q.HW;
  for (int n = 1; n < K; ++n) {
    lu_cta_flood_label(t, q, n_frame, n, lab, small, big, sh_changed, [&](int p) {
      const int v = lab[p] + n;
      lab[p] = v;
      if (v < K) {                   // the pixel now belongs to label v: its box must contain it when v's turn comes
        int32_t* b = q.bbox + (n_frame * q.KMAX + v) * 4;
        const int y = p / q.W, x = p % q.W;
        lu_atomic_min_i(b + 0, y); lu_atomic_min_i(b + 1, -y);
        lu_atomic_min_i(b + 2, x); lu_atomic_min_i(b + 3, -x);
      }
    });
    LU_CTA_SYNC();
  }
}

// item = pixel: final label value; labels seen inside the field of view (Inference2D.py:94-104, incl. the reference's
// single zeroed column on the left side)
struct LuPpCombine {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    const int K = q.info[n * 4 + 0];
    int v = q.lab[i];
    if (q.info[n * 4 + 2] == 0) v += q.add[i];
    q.lab[i] = v;
    if (q.fov > 0 && v >= 1 && v < K) {
      const int y = p / q.W, x = p % q.W;
      if (y >= q.fov && y < q.H - q.fov && x != q.fov && x < q.W - q.fov) q.present[n * q.KMAX + v] = 1;
    }
  }
};
// one CTA per frame: keep flags -> consecutive new labels
LU_HDI void lu_pp_relabel_cta(const LuCta& t, const LuPost& q, int64_t n_frame, int32_t* sums) {
  const int K = q.info[n_frame * 4 + 0];
  int32_t* nl = q.newlab + n_frame * q.KMAX;
  const int32_t* ar = q.larea + n_frame * q.KMAX;
  const int32_t* pr = q.present + n_frame * q.KMAX;
  for (int n = t.tid; n < K; n += t.nthreads)
    nl[n] = (n >= 1 && ar[n] >= q.min_size && ar[n] <= q.max_size && (q.fov == 0 || pr[n])) ? 1 : 0;
  LU_CTA_SYNC();
  const int kept = lu_cta_scan(t, nl, K, sums);
  for (int n = t.tid; n < K; n += t.nthreads) {
    const bool keep = n >= 1 && ar[n] >= q.min_size && ar[n] <= q.max_size && (q.fov == 0 || pr[n]);
    if (!keep) nl[n] = 0;
  }
  if (t.tid == 0) q.info[n_frame * 4 + 1] = kept;
}
LU_HDI void lu_pp_rank_cta(const LuCta& t, const LuPost& q, int64_t n_frame, int32_t* sums) {
  const int total = lu_cta_scan(t, q.bflag + n_frame * q.NB, q.NB, sums);
  if (t.tid == 0) q.info[n_frame * 4 + 0] = total + 1;
}
struct LuPpOutput {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int64_t n = i / q.HW;
    const int v = q.lab[i];
    q.out[i] = (v >= 1 && v < q.info[n * 4 + 0]) ? (uint16_t)q.newlab[n * q.KMAX + v] : (uint16_t)0;
  }
};

#ifndef LU_HOST_EMU
__global__ void __launch_bounds__(1024) lu_pp_rank_kernel(LuPost q) {
  __shared__ int32_t sums[1025];
  lu_pp_rank_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, sums);
}
__global__ void __launch_bounds__(1024) lu_pp_relabel_kernel(LuPost q) {
  __shared__ int32_t sums[1025];
  lu_pp_relabel_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, sums);
}
__global__ void __launch_bounds__(LU_PP_CTA) lu_pp_holes_kernel(LuPost q) {
  __shared__ uint8_t small[LU_PP_SMEM_CROP];
  __shared__ int32_t changed;
  lu_pp_holes_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, blockIdx.y, small, &changed);
}
__global__ void __launch_bounds__(256) lu_pp_holes_seq_kernel(LuPost q) {
  __shared__ uint8_t small[LU_PP_SMEM_CROP];
  __shared__ int32_t changed;
  lu_pp_holes_seq_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, small, &changed);
}
#endif
