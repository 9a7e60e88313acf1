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
// Pixel kernels are index functors (lu_parallel_for_impl; run links and small reductions use warp ballots / shuffles over
// the 32 consecutive items of a warp, with a serial equivalent in the host build) and the CTA kernels are written as
// phase-separated strided loops, so the TEST-ONLY host build (LU_HOST_EMU) runs the same code with one "thread" per CTA.
#pragma once
#include "lu_elem.cuh"

#define LU_PP_EMPTY 0x7f7f7f7f   // memset(0x7f) pattern: "no pixel yet" in the bounding boxes / block keys
#define LU_PP_SMEM_CROP 24576    // crops up to this many pixels are flooded in shared memory
#define LU_PP_CTA 128
#define LU_PP_SLABS 4            // CTAs per frame that flood the (rare) crops too large for shared memory

enum { LU_PP_CELL0 = 1, LU_PP_EDGE0 = 2, LU_PP_CELL = 4, LU_PP_EDGE = 8 };

struct LuPost {
  const float* sm;
  int chw;                            // 1: (N,3,H,W), 0: (N,H,W,3)
  int N, H, W, HW, WB, HB, NB, KMAX, G;   // WB x HB 2x2 blocks; KMAX = NB + 1 labels at most; G flood CTAs per frame
  float edge_thresh;
  int d2lim, rad, min_size, max_size, fov;
  uint8_t* cls;                       // [N][HW]
  int32_t *parA, *parB, *key, *area, *cc, *lab, *add;   // [N][HW]
  uint8_t* bflag;                     // [N][NB]     1 = a component starts in this 2x2 block
  int32_t* rowcnt;                    // [N][HB]     flagged blocks per block row, then their inclusive scan
  int32_t *larea, *present, *newlab;  // [N][KMAX]
  int32_t* bbox;                      // [N][KMAX][4] = rmin, -rmax, cmin, -cmax
  int32_t* info;                      // [N][4] = label count (incl. background), kept, sequential-redo flag, big crops
  uint8_t* slab;                      // [N][LU_PP_SLABS][HW] flood scratch for crops larger than LU_PP_SMEM_CROP
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


// ---- run links with coalesced accesses ---------------------------------------------------------------------------------
// Items are pixels in linear order; `cont` says "this pixel continues the run of the pixel on its left".  Returns how
// far to the left the link of this pixel points: to the first pixel of its run inside the aligned group of 32 items, or
// to the pixel just before the group when the run started earlier (0 = the pixel heads a run).  On the GPU the group is
// the warp (lu_pf_kernel maps consecutive items to consecutive lanes; all lanes with i < n_items must call this); the
// host build gets the same answer from the link of the previous pixel, which it has already written.
LU_HDI int lu_run_back(int64_t i, int64_t n_items, bool cont, const int32_t* par_left, int p) {
#ifdef __CUDA_ARCH__
  const int lane = (int)(threadIdx.x & 31);
  const int64_t left = n_items - (i - lane);
  const unsigned act = left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u);
  const unsigned m = __ballot_sync(act, cont);
  (void)par_left; (void)p;
  if (!cont) return 0;
  const unsigned z = ~m & ((1u << lane) - 1u);
  return z ? lane - (31 - __clz(z)) : lane + 1;
#else
  (void)n_items;
  if (!cont) return 0;
  return (i & 31) == 0 ? 1 : p - par_left[0];
#endif
}
// value of `v` in the previous item of the group (lane - 1); lane 0 gets `fallback`
LU_HDI int lu_prev_item(int64_t i, int64_t n_items, int v, int fallback) {
#ifdef __CUDA_ARCH__
  const int lane = (int)(threadIdx.x & 31);
  const int64_t left = n_items - (i - lane);
  const unsigned act = left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u);
  const int u = __shfl_up_sync(act, v, 1);
  return lane == 0 ? fallback : u;
#else
  (void)i; (void)n_items; (void)v;
  return fallback;
#endif
}

// np.argmax(softmax, 0) == 1 and not edge  (first maximum wins; NaN counts as the maximum, like numpy)
LU_HDI int lu_pp_classify(float s0, float s1, float s2, float thr) {
  const int edge = s2 >= thr;
  int am;
  if (s0 != s0) am = 0; else if (s1 != s1) am = 1; else if (s2 != s2) am = 2;
  else am = (s1 > s0) ? ((s2 > s1) ? 2 : 1) : ((s2 > s0) ? 2 : 0);
  return ((am == 1 && !edge) ? LU_PP_CELL0 : 0) | (edge ? LU_PP_EDGE0 : 0);
}

// item = pixel: classify, link background runs (each pixel -> first pixel of its run inside the group of 32 items; a
// run continuing from the previous group links to the pixel on its left)
struct LuPpClassify {
  LuPost q; int64_t n_items;
  LU_HD int cls_at(int64_t n, int p) const {
    float s0, s1, s2;
    if (q.chw) { const float* b = q.sm + n * 3 * (int64_t)q.HW + p; s0 = b[0]; s1 = b[q.HW]; s2 = b[2 * (int64_t)q.HW]; }
    else { const float* b = q.sm + (n * (int64_t)q.HW + p) * 3; s0 = b[0]; s1 = b[1]; s2 = b[2]; }
    return lu_pp_classify(s0, s1, s2, q.edge_thresh);
  }
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    const int x = p % q.W;
    const int c = cls_at(n, p);
    const int cl = lu_prev_item(i, n_items, c, (i & 31) == 0 && x > 0 ? cls_at(n, p - 1) : 0);
#ifdef LU_HOST_EMU
    const int cprev = x > 0 ? ((i & 31) == 0 ? cl : q.cls[i - 1]) : 0;
#else
    const int cprev = x > 0 ? cl : 0;
#endif
    const bool bg = !(c & LU_PP_CELL0);
    const bool cont = bg && x > 0 && !(cprev & LU_PP_CELL0);
    const int back = lu_run_back(i, n_items, cont, q.parA + i - 1, p);
    q.cls[i] = (uint8_t)c;
    q.parA[i] = p - back;
    q.cc[i] = 0;                       // "background component touches the frame border" flags, per root
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

// item = pixel: filled cell mask (cell, or background component that never reaches the border), edge pixels that are
// not cell, run links of the cell mask, per-root accumulators reset
struct LuPpFill {
  LuPost q; int64_t n_items;
  LU_HD bool filled(int64_t fo, int p) const {
    if (q.cls[fo + p] & LU_PP_CELL0) return true;
    return q.cc[fo + q.parA[fo + p]] == 0;
  }
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    const int x = p % q.W;
    const bool f = filled(fo, p);
    const int fl = lu_prev_item(i, n_items, f ? 1 : 0, (i & 31) == 0 && x > 0 ? (filled(fo, p - 1) ? 1 : 0) : 0);
#ifdef LU_HOST_EMU
    const bool fprev = x > 0 && ((i & 31) == 0 ? fl != 0 : (q.cls[i - 1] & LU_PP_CELL) != 0);
#else
    const bool fprev = x > 0 && fl != 0;
#endif
    int c = q.cls[i] & (LU_PP_CELL0 | LU_PP_EDGE0);
    if (f) c |= LU_PP_CELL; else if (c & LU_PP_EDGE0) c |= LU_PP_EDGE;
    const int back = lu_run_back(i, n_items, f && fprev, q.parB + i - 1, p);
    q.cls[i] = (uint8_t)c;
    q.parB[i] = p - back;
    q.key[i] = LU_PP_EMPTY;
    q.area[i] = 0;
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
    if ((q.cls[i] & LU_PP_CELL) && q.parB[i] == p) {
      const int k = q.key[i];
      q.bflag[n * q.NB + k] = 1;
      lu_atomic_add_i(q.rowcnt + n * q.HB + k / q.WB, 1);
    }
  }
};
// item = pixel (component roots only): OpenCV's number of the component = 1 + components starting in earlier block rows
// + components starting further left in the same block row; the number replaces the block key at the root
struct LuPpRootLabel {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    if (!(q.cls[i] & LU_PP_CELL) || q.parB[i] != p) return;
    const int k = q.key[i], by = k / q.WB, bx = k % q.WB;
    int l = 1 + (by > 0 ? q.rowcnt[n * q.HB + by - 1] : 0);
    const uint8_t* f = q.bflag + n * q.NB + (int64_t)by * q.WB;
    for (int b = 0; b < bx; ++b) l += f[b];
    q.key[i] = l;
    q.larea[n * q.KMAX + l] = q.area[i];
  }
};

// item = pixel: OpenCV label of every cell pixel, core areas per label
struct LuPpAssign {
  LuPost q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    q.cc[i] = (q.cls[i] & LU_PP_CELL) ? q.key[fo + q.parB[i]] : 0;
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
    if (l > 0) {                       // a (possibly stale) read first: interior pixels skip the atomics
      int32_t* b = q.bbox + (n * q.KMAX + l) * 4;
      if (y < b[0]) lu_atomic_min_i(b + 0, y);
      if (-y < b[1]) lu_atomic_min_i(b + 1, -y);
      if (x < b[2]) lu_atomic_min_i(b + 2, x);
      if (-x < b[3]) lu_atomic_min_i(b + 3, -x);
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
// big == nullptr: a crop that does not fit `small` is appended to the frame's list of big crops (newlab[], count in
// info[3]) for lu_pp_holes_big_cta instead of being flooded here.
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
  if (s == nullptr) {
    if (t.tid == 0) q.newlab[n_frame * q.KMAX + lu_atomic_add_i(q.info + n_frame * 4 + 3, 1)] = n;
    return;
  }
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
  for (int n = 1 + g; n < K; n += q.G)
    lu_cta_flood_label(t, q, n_frame, n, lab, small, (uint8_t*)nullptr, sh_changed, [&](int p) {
      lu_atomic_add_i(add + p, n);
      if (lab[p] != 0) *flag = 1;
    });
}
// crops that did not fit shared memory: LU_PP_SLABS CTAs per frame, each with a frame-sized scratch in global memory
LU_HDI void lu_pp_holes_big_cta(const LuCta& t, const LuPost& q, int g, int64_t n_frame, uint8_t* small, int32_t* sh_changed) {
  const int nbig = q.info[n_frame * 4 + 3];
  const int64_t fo = n_frame * (int64_t)q.HW;
  const int32_t* lab = q.lab + fo;
  int32_t* add = q.add + fo;
  int32_t* flag = q.info + n_frame * 4 + 2;
  uint8_t* big = q.slab + (n_frame * LU_PP_SLABS + g) * (int64_t)q.HW;
  for (int j = g; j < nbig; j += LU_PP_SLABS) {
    const int n = q.newlab[n_frame * q.KMAX + j];
    lu_cta_flood_label(t, q, n_frame, n, lab, small, big, sh_changed, [&](int p) {
      lu_atomic_add_i(add + p, n);
      if (lab[p] != 0) *flag = 1;
    });
  }
}
// the reference's order, on the running label image (only when the flag is up): one CTA per frame
LU_HDI void lu_pp_holes_seq_cta(const LuCta& t, const LuPost& q, int64_t n_frame, uint8_t* small, int32_t* sh_changed) {
  if (q.info[n_frame * 4 + 2] == 0) return;
  const int K = q.info[n_frame * 4 + 0];
  const int64_t fo = n_frame * (int64_t)q.HW;
  int32_t* lab = q.lab + fo;
  uint8_t* big = q.slab + (n_frame * LU_PP_SLABS) * (int64_t)q.HW;
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
  const int total = lu_cta_scan(t, q.rowcnt + n_frame * q.HB, q.HB, sums);
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
__global__ void __launch_bounds__(256) lu_pp_rank_kernel(LuPost q) {
  __shared__ int32_t sums[257];
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
__global__ void __launch_bounds__(1024) lu_pp_holes_big_kernel(LuPost q) {
  __shared__ uint8_t small[16];
  __shared__ int32_t changed;
  lu_pp_holes_big_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, blockIdx.y, small, &changed);
}
__global__ void __launch_bounds__(256) lu_pp_holes_seq_kernel(LuPost q) {
  __shared__ uint8_t small[LU_PP_SMEM_CROP];
  __shared__ int32_t changed;
  lu_pp_holes_seq_cta(LuCta{(int)threadIdx.x, (int)blockDim.x}, q, blockIdx.x, small, &changed);
}
#endif

// =====================================================================================================================
// SEG measure + sparse categorical accuracy of a train / validation step (losses.py:29-88, train2D.py:97-102): the
// reference computes them on the host inside every step (tf.py_function: SciPy labelling + Python loops per object).
// Here: 4-connected union-find of the ground-truth foreground (label == 1) and of the predicted foreground
// (argmax == 1), object areas, intersection counts of every (truth object, predicted object) pair in an open-addressing
// hash table, and the score of every truth object = IoU with the predicted object covering more than half of it.
// =====================================================================================================================
enum { LU_SEG_GT = 1, LU_SEG_OUT = 2 };
#define LU_SEG_NOKEY (-1ll)

struct LuSeg {
  const float* labels; const float* logits;
  int chw, N, H, W, HW, cap;          // cap: hash slots per frame (power of two >= HW)
  uint8_t* cls;                       // [N][HW]
  int32_t *parG, *parS, *areaG, *areaS;   // [N][HW]
  float* score;                       // [N][HW] at truth roots
  long long* keys; int32_t* cnt;      // [N][cap]
  double* result;                     // [4] = sum of scores, truth objects, correct pixels, pixels
};

LU_HDI long long lu_atomic_cas64(long long* p, long long expect, long long v) {
#ifdef __CUDA_ARCH__
  return (long long)atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)expect, (unsigned long long)v);
#else
  const long long o = *p; if (o == expect) *p = v; return o;
#endif
}

// number of items of the 32-group with `pred` set, returned to the first item of the group only (others get 0)
LU_HDI int lu_group_count(int64_t i, int64_t n_items, bool pred) {
#ifdef __CUDA_ARCH__
  const int lane = (int)(threadIdx.x & 31);
  const int64_t left = n_items - (i - lane);
  const unsigned act = left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u);
  const unsigned m = __ballot_sync(act, pred);
  return lane == 0 ? __popc(m) : 0;
#else
  (void)i; (void)n_items;
  return pred ? 1 : 0;
#endif
}

// item = pixel: foregrounds, run links of both masks, correct-pixel count
struct LuSegClassify {
  LuSeg q; int64_t n_items;
  LU_HD int cls_at(int64_t n, int p, int* correct) const {
    float s0, s1, s2;
    if (q.chw) { const float* b = q.logits + n * 3 * (int64_t)q.HW + p; s0 = b[0]; s1 = b[q.HW]; s2 = b[2 * (int64_t)q.HW]; }
    else { const float* b = q.logits + (n * (int64_t)q.HW + p) * 3; s0 = b[0]; s1 = b[1]; s2 = b[2]; }
    int am;
    if (s0 != s0) am = 0; else if (s1 != s1) am = 1; else if (s2 != s2) am = 2;
    else am = (s1 > s0) ? ((s2 > s1) ? 2 : 1) : ((s2 > s0) ? 2 : 0);
    const float l = q.labels[n * (int64_t)q.HW + p];
    if (correct) *correct = ((float)am == l) ? 1 : 0;
    return (l == 1.0f ? LU_SEG_GT : 0) | (am == 1 ? LU_SEG_OUT : 0);
  }
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    const int x = p % q.W;
    int ok;
    const int c = cls_at(n, p, &ok);
    const int cl = lu_prev_item(i, n_items, c, (i & 31) == 0 && x > 0 ? cls_at(n, p - 1, nullptr) : 0);
#ifdef LU_HOST_EMU
    const int cprev = x > 0 ? ((i & 31) == 0 ? cl : q.cls[i - 1]) : 0;
#else
    const int cprev = x > 0 ? cl : 0;
#endif
    const int backG = lu_run_back(i, n_items, (c & cprev & LU_SEG_GT) != 0, q.parG + i - 1, p);
    const int backS = lu_run_back(i, n_items, (c & cprev & LU_SEG_OUT) != 0, q.parS + i - 1, p);
    const int cnt = lu_group_count(i, n_items, ok != 0);
    q.cls[i] = (uint8_t)c;
    q.parG[i] = p - backG; q.parS[i] = p - backS;
    q.areaG[i] = 0; q.areaS[i] = 0; q.score[i] = 0.f;
    if (cnt) lu_atomic_add(q.result + 2, (double)cnt);
  }
};
struct LuSegMerge {       // item = pixel: vertical links of both 4-connected masks
  LuSeg q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    const int y = p / q.W, x = p % q.W;
    if (y == 0) return;
    const uint8_t* c = q.cls + fo;
    const int both = c[p] & c[p - q.W];
    if (!both) return;
    const int skip = x > 0 ? (c[p - 1] & c[p - q.W - 1]) : 0;       // left and upper-left carry the link already
    if ((both & LU_SEG_GT) && !(skip & LU_SEG_GT)) lu_uf_union(q.parG + fo, p, p - q.W);
    if ((both & LU_SEG_OUT) && !(skip & LU_SEG_OUT)) lu_uf_union(q.parS + fo, p, p - q.W);
  }
};
struct LuSegFlatten {     // item = pixel: roots + object areas
  LuSeg q;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t fo = i - p;
    const int c = q.cls[i];
    if (c & LU_SEG_GT) { const int r = lu_uf_find(q.parG + fo, p); q.parG[i] = r; lu_atomic_add_i(q.areaG + fo + r, 1); }
    if (c & LU_SEG_OUT) { const int r = lu_uf_find(q.parS + fo, p); q.parS[i] = r; lu_atomic_add_i(q.areaS + fo + r, 1); }
  }
};
// item = pixel: intersection pixels of every (truth root, predicted root) pair; on the GPU one insertion per run of equal
// pairs inside the 32-group
struct LuSegPairs {
  LuSeg q; int64_t n_items;
  LU_HD void insert(int64_t n, long long key, int k) const {
    long long* keys = q.keys + n * (int64_t)q.cap;
    unsigned long long h = (unsigned long long)key * 0x9E3779B97F4A7C15ull;
    int slot = (int)(h >> 40) & (q.cap - 1);
    for (;;) {
      const long long old = lu_atomic_cas64(keys + slot, LU_SEG_NOKEY, key);
      if (old == LU_SEG_NOKEY || old == key) { lu_atomic_add_i(q.cnt + n * (int64_t)q.cap + slot, k); return; }
      slot = (slot + 1) & (q.cap - 1);
    }
  }
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW); const int64_t n = i / q.HW;
    long long key = LU_SEG_NOKEY;
    if (q.cls[i] == (LU_SEG_GT | LU_SEG_OUT)) key = (long long)q.parG[i] * q.HW + q.parS[i];
    int k = key != LU_SEG_NOKEY ? 1 : 0;
#ifdef __CUDA_ARCH__
    const int lane = (int)(threadIdx.x & 31);
    const int64_t left = n_items - (i - lane);
    const unsigned act = left >= 32 ? 0xffffffffu : ((1u << (int)left) - 1u);
    const long long kprev = __shfl_up_sync(act, key, 1);
    // same pair as the previous item of the same frame (roots are frame-local, so compare the frame too)
    const bool same = lane > 0 && key != LU_SEG_NOKEY && key == kprev && p > 0;
    const unsigned sm = __ballot_sync(act, same);
    if (same) k = 0;
    else if (k && lane < 31) k += __ffs(~(sm >> (lane + 1))) - 1;       // length of the run this item heads
#endif
    if (k) insert(n, key, k);
  }
};
struct LuSegScore {       // item = hash slot: IoU of a pair whose overlap exceeds half of the truth object (float32 like numpy)
  LuSeg q;
  LU_HD void operator()(int64_t i) const {
    const long long key = q.keys[i];
    if (key == LU_SEG_NOKEY) return;
    const int64_t fo = (i / q.cap) * (int64_t)q.HW;
    const int g = (int)(key / q.HW), s = (int)(key % q.HW);
    const float inter = (float)q.cnt[i], l_area = (float)q.areaG[fo + g];
    if (inter / l_area > 0.5f) q.score[fo + g] = inter / (l_area + (float)q.areaS[fo + s] - inter);
  }
};
struct LuSegReduce {      // item = pixel: sum of the truth objects' scores, object and pixel counts
  LuSeg q; int64_t n_items;
  LU_HD void operator()(int64_t i) const {
    const int p = (int)(i % q.HW);
    if (i == 0) lu_atomic_add(q.result + 3, (double)n_items);
    if ((q.cls[i] & LU_SEG_GT) && q.parG[i] == p) { lu_atomic_add(q.result + 0, (double)q.score[i]); lu_atomic_add(q.result + 1, 1.0); }
  }
};
