// Training-reader augmentation on the device (DataHandeling.py:150-395): the per-frame chain of
// CTCRAMReaderSequence2D._load_and_enqueue -- contrast / brightness, random affine (cv2.warpAffine) + elastic
// (scipy map_coordinates) warp of image and segmentation, _fix_transformed_segmentation, flips, rot90 -- and the
// elastic displacement field (_get_indices4elastic_transform: two Gaussian-filtered random fields).
// The reference runs this in Python worker threads per frame; here a sequence chunk is augmented by a handful of
// HBM-bound kernels and written straight into the (B,T,1,H,W) batch tensors.
//
// Arithmetic follows the libraries the reference calls, restated in oracle/augment_oracle.py: OpenCV's fixed-point
// affine coordinates (1/1024 px, 1/32-px bilinear table, float32 weights and sums), SciPy's float64 map_coordinates
// (order 1 'reflect' / order 0 'constant') and correlate1d summation order.  Float products and sums are kept unfused
// (__fmul_rn / __fadd_rn) so the results are those of the CPU libraries; the only value not reproduced bit for bit is
// the frame mean of the contrast step (float64 atomics here, numpy's float32 pairwise sum there).
#pragma once
#include "lu_elem.cuh"

LU_HDI float lu_fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
LU_HDI float lu_fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
LU_HDI double lu_dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
LU_HDI double lu_dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
LU_HDI long long lu_lrint(double v) {
#ifdef __CUDA_ARCH__
  return __double2ll_rn(v);
#else
  return (long long)nearbyint(v);
#endif
}

struct LuAug {
  const float* img; const float* seg;         // (frames, H, W)
  const float* contrast; const float* brightness;   // (frames)
  const double* coords;                       // (2, H, W): y then x sampling coordinates, or NULL (no elastic warp)
  int frames, H, W, HW;
  int randomize, elastic, flip0, flip1, rot90;
  double mi[6];                               // INVERTED affine matrix (cv2.warpAffine's float64 inversion, on the host)
  double* sums;                               // (frames, 2): sum of the image, count of seg pixels != -1
  float *adj, *warp;                          // (frames, H, W) scratch: adjusted image, affine-warped image
  float *wseg, *wnv, *tseg, *tnv;             // segmentation / not-valid mask after the affine warp and after the map
  float* out_img; float* out_seg;             // (frames, Ho, Wo), Ho x Wo = H x W or W x H (odd rot90)
};

// destination of source pixel (y, x) after cv2.flip(.,0), cv2.flip(.,1), np.rot90(., k)
LU_HDI int64_t lu_aug_dest(const LuAug& q, int y, int x) {
  int h = q.H, w = q.W;
  if (q.flip0) y = h - 1 - y;
  if (q.flip1) x = w - 1 - x;
  for (int k = 0; k < q.rot90; ++k) { const int ny = w - 1 - x, nx = y; y = ny; x = nx; const int t = h; h = w; w = t; }
  return (int64_t)y * w + x;
}
LU_HDI int lu_reflect101(long long i, int n) {      // cv::borderInterpolate, BORDER_REFLECT_101
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  long long m = i % p; if (m < 0) m += p;
  return (int)(m >= n ? p - m : m);
}
LU_HDI double lu_ni_reflect_coord(double c, int n) {        // scipy map_coordinate(), NI_EXTEND_REFLECT
  const double s2 = 2.0 * n;
  if (c < 0) {
    if (c < -s2) c = s2 * (double)(long long)(-c / s2) + c;
    c = c < -n ? c + s2 : -c - 1;
  } else if (c > n - 1) {
    c -= s2 * (double)(long long)(c / s2);
    if (c >= n) c = s2 - c - 1;
  }
  return c;
}
LU_HDI int lu_ni_reflect_index(long long i, int n) {
  const long long s2 = 2ll * n;
  if (i < 0) {
    if (i < -s2) i = s2 * ((-i) / s2) + i;
    i = i < -n ? i + s2 : -i - 1;
  } else if (i >= n) {
    i -= s2 * (i / s2);
    if (i >= n) i = s2 - i - 1;
  }
  return (int)i;
}

// sum of `v` over the 32-group of items, returned to its first item (others get 0); all items of the group must call
LU_HDI double lu_group_sum(double v) {
#ifdef __CUDA_ARCH__
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return (threadIdx.x & 31) == 0 ? v : 0.0;
#else
  return v;
#endif
}
// item = (frame, chunk of 1024 pixels, lane): each item strides through its chunk (coalesced across the 32 lanes of a
// chunk); one atomic per chunk: frame sum of the image and number of annotated segmentation pixels
struct LuAugStats {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int chunks = (q.HW + 1023) / 1024;
    const int lane = (int)(i & 31); const int64_t r = i >> 5;
    const int f = (int)(r / chunks), c = (int)(r % chunks);
    const int p1 = (c + 1) * 1024 < q.HW ? (c + 1) * 1024 : q.HW;
    double s = 0, ann = 0;
    for (int p = c * 1024 + lane; p < p1; p += 32) { s += (double)q.img[(int64_t)f * q.HW + p]; ann += q.seg[(int64_t)f * q.HW + p] != -1.0f ? 1.0 : 0.0; }
    s = lu_group_sum(s); ann = lu_group_sum(ann);
    if (s != 0.0) lu_atomic_add(q.sums + 2 * f, s);
    if (ann != 0.0) lu_atomic_add(q.sums + 2 * f + 1, ann);
  }
};
// item = pixel: contrast about the frame mean, brightness (DataHandeling.py:215-240, float32 like numpy)
struct LuAugAdjust {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int f = (int)(i / q.HW);
    float v = q.img[i];
    if (q.randomize) {
      const float m = (float)(q.sums[2 * f] / (double)q.HW);
      v = lu_fadd(lu_fmul(lu_fadd(v, -m), q.contrast[f]), m);
      v = lu_fadd(v, q.brightness[f]);
    }
    q.adj[i] = v;
  }
};
// item = pixel: cv2.warpAffine -- bilinear / BORDER_REFLECT_101 for the image, nearest / constant -1 for the
// segmentation (frame border zeroed first, DataHandeling.py:343-346) and for its not-annotated mask
struct LuAugWarp {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int f = (int)(i / q.HW), p = (int)(i % q.HW);
    const int y = p / q.W, x = p % q.W;
    // unfused float64 products / sums: the rounding to the 1/1024-pixel grid must see OpenCV's values
    const long long ad = lu_lrint(lu_dmul(lu_dmul(q.mi[0], (double)x), 1024.0)), bd = lu_lrint(lu_dmul(lu_dmul(q.mi[3], (double)x), 1024.0));
    const long long xb = lu_lrint(lu_dmul(lu_dadd(lu_dmul(q.mi[1], (double)y), q.mi[2]), 1024.0));
    const long long yb = lu_lrint(lu_dmul(lu_dadd(lu_dmul(q.mi[4], (double)y), q.mi[5]), 1024.0));
    {   // linear: round_delta = 1024 / 32 / 2, coordinates in 1/32 pixel
      const long long X = (xb + 16 + ad) >> 5, Y = (yb + 16 + bd) >> 5;
      long long sx = X >> 5, sy = Y >> 5;
      sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx); sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
      const float fx = (float)(X & 31) / 32.0f, fy = (float)(Y & 31) / 32.0f;
      const int x0 = lu_reflect101(sx, q.W), x1 = lu_reflect101(sx + 1, q.W);
      const int y0 = lu_reflect101(sy, q.H), y1 = lu_reflect101(sy + 1, q.H);
      const float* a = q.adj + (int64_t)f * q.HW;
      const float w00 = lu_fmul(1.0f - fy, 1.0f - fx), w01 = lu_fmul(1.0f - fy, fx), w10 = lu_fmul(fy, 1.0f - fx), w11 = lu_fmul(fy, fx);
      float v = lu_fmul(a[y0 * q.W + x0], w00);
      v = lu_fadd(v, lu_fmul(a[y0 * q.W + x1], w01));
      v = lu_fadd(v, lu_fmul(a[y1 * q.W + x0], w10));
      v = lu_fadd(v, lu_fmul(a[y1 * q.W + x1], w11));
      q.warp[i] = v;
    }
    {   // nearest: round_delta = 512, integer pixel
      const long long X = (xb + 512 + ad) >> 10, Y = (yb + 512 + bd) >> 10;
      float s = -1.0f, nv = -1.0f;
      if (X >= 0 && X < q.W && Y >= 0 && Y < q.H) {
        const float o = q.seg[(int64_t)f * q.HW + Y * q.W + X];
        const bool border = X == 0 || X == q.W - 1 || Y == 0 || Y == q.H - 1;
        s = border ? 0.0f : o;
        nv = o == -1.0f ? 1.0f : 0.0f;
      }
      q.wseg[i] = s; q.wnv[i] = nv;
    }
  }
};
// item = pixel: scipy map_coordinates at the elastic coordinates -- order 1 'reflect' for the image (float64 sums in
// SciPy's order), order 0 'constant' -1 for the segmentation and the mask; the image goes to its final position
struct LuAugMap {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int f = (int)(i / q.HW), p = (int)(i % q.HW);
    const int y = p / q.W, x = p % q.W;
    const double cy0 = q.coords[p], cx0 = q.coords[q.HW + p];
    {
      const double cy = lu_ni_reflect_coord(cy0, q.H), cx = lu_ni_reflect_coord(cx0, q.W);
      const double fy = floor(cy), fx = floor(cx), ty = cy - fy, tx = cx - fx;
      const int y0 = lu_ni_reflect_index((long long)fy, q.H), y1 = lu_ni_reflect_index((long long)fy + 1, q.H);
      const int x0 = lu_ni_reflect_index((long long)fx, q.W), x1 = lu_ni_reflect_index((long long)fx + 1, q.W);
      const float* w = q.warp + (int64_t)f * q.HW;
      double t = lu_dmul(lu_dmul((double)w[y0 * q.W + x0], 1.0 - ty), 1.0 - tx);
      t = lu_dadd(t, lu_dmul(lu_dmul((double)w[y0 * q.W + x1], 1.0 - ty), tx));
      t = lu_dadd(t, lu_dmul(lu_dmul((double)w[y1 * q.W + x0], ty), 1.0 - tx));
      t = lu_dadd(t, lu_dmul(lu_dmul((double)w[y1 * q.W + x1], ty), tx));
      q.out_img[(int64_t)f * q.HW + lu_aug_dest(q, y, x)] = (float)t;
    }
    {
      float s = -1.0f, nv = -1.0f;
      if (cy0 >= 0 && cy0 <= q.H - 1 && cx0 >= 0 && cx0 <= q.W - 1) {
        const int iy = (int)floor(cy0 + 0.5), ix = (int)floor(cx0 + 0.5);
        s = q.wseg[(int64_t)f * q.HW + iy * q.W + ix];
        nv = q.wnv[(int64_t)f * q.HW + iy * q.W + ix];
      }
      q.tseg[i] = s; q.tnv[i] = nv;
    }
  }
};
// item = pixel, no elastic warp: the adjusted image goes straight to its final position
struct LuAugPlace {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int f = (int)(i / q.HW), p = (int)(i % q.HW);
    q.out_img[(int64_t)f * q.HW + lu_aug_dest(q, p / q.W, p % q.W)] = q.adj[i];
  }
};
// item = pixel: _fix_transformed_segmentation (round, 3x3 grey dilation with SciPy's 'reflect' border, touching
// objects -> 2), pixels without annotation -> -1; frames without any annotation pass through (all -1)
struct LuAugSegFix {
  LuAug q;
  LU_HD void operator()(int64_t i) const {
    const int f = (int)(i / q.HW), p = (int)(i % q.HW);
    const int y = p / q.W, x = p % q.W;
    const float* src = (q.elastic ? q.tseg : q.seg) + (int64_t)f * q.HW;
    float out;
    if (q.elastic && q.sums[2 * f + 1] == 0.0) out = q.seg[i];        // np.equal(seg_crop, -1).all()
    else {
      const float r = rintf(src[p]);
      int dil = (int)r;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          yy = yy < 0 ? 0 : (yy >= q.H ? q.H - 1 : yy);               // half-sample symmetric == clamp for radius 1
          xx = xx < 0 ? 0 : (xx >= q.W ? q.W - 1 : xx);
          const int v = (int)rintf(src[yy * q.W + xx]);
          dil = v > dil ? v : dil;
        }
      out = r < 1.0f ? r : 1.0f;
      if ((float)dil != r && dil > 0) out = 2.0f;
      if (q.elastic && (q.tnv[i] > 0.5f || src[p] == -1.0f)) out = -1.0f;
    }
    q.out_seg[(int64_t)f * q.HW + lu_aug_dest(q, y, x)] = out;
  }
};

// ---- elastic displacement field: scipy.ndimage.gaussian_filter (two correlate1d passes, 'reflect') ------------------
struct LuGauss {
  const double* in; double* out; const double* w;     // w[2*lw+1]
  int H, W, lw, axis, fields;
  double scale, offset_in;                            // first pass reads rand*2-1: in*scale + offset_in
  double alpha; int add_grid;                         // second pass: out = grid + alpha * filtered
};
struct LuGaussPass {
  LuGauss g;
  LU_HD double at(const double* f, int y, int x) const { return f[(int64_t)y * g.W + x] * g.scale + g.offset_in; }
  LU_HD void operator()(int64_t i) const {
    const int HW = g.H * g.W;
    const int fi = (int)(i / HW), p = (int)(i % HW);
    const int y = p / g.W, x = p % g.W;
    const double* f = g.in + (int64_t)fi * HW;
    const int n = g.axis == 0 ? g.H : g.W, c = g.axis == 0 ? y : x;
    double t;
    if (g.axis == 0) {
      t = at(f, y, x) * g.w[g.lw];
      for (int j = -g.lw; j < 0; ++j)
        t = lu_dadd(t, lu_dmul(lu_dadd(at(f, lu_ni_reflect_index(c + j, n), x), at(f, lu_ni_reflect_index(c - j, n), x)), g.w[j + g.lw]));
    } else {
      t = at(f, y, x) * g.w[g.lw];
      for (int j = -g.lw; j < 0; ++j)
        t = lu_dadd(t, lu_dmul(lu_dadd(at(f, y, lu_ni_reflect_index(c + j, n)), at(f, y, lu_ni_reflect_index(c - j, n))), g.w[j + g.lw]));
    }
    if (g.add_grid) {
      // fields arrive as (x field, y field) and leave as coords (y + dy, x + dx): swap on output
      const double base = fi == 0 ? (double)x : (double)y;
      g.out[(int64_t)(1 - fi) * HW + p] = lu_dadd(base, lu_dmul(t, g.alpha));
    } else g.out[i] = t;
  }
};
