// Host side of the device augmentation (lu_aug.cuh): workspace layout, launch sequence, C-ABI.  Included by lu_api.cu.

struct AugLayout { size_t sums, adj, warp, wseg, wnv, tseg, tnv, total; };
static int aug_layout(int frames, int H, int W, AugLayout* L) {
  LU_REQUIRE(frames > 0 && H > 0 && W > 0 && (int64_t)frames * H * W < (1ll << 31), "augmentation needs 0 < frames*H*W < 2^31");
  const size_t n = (size_t)frames * H * W * 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L->sums = take((size_t)frames * 16);
  L->adj = take(n); L->warp = take(n); L->wseg = take(n); L->wnv = take(n); L->tseg = take(n); L->tnv = take(n);
  L->total = off;
  return 0;
}
extern "C" int lu_aug_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes) {
  AugLayout L;
  if (aug_layout(frames, H, W, &L)) return 1;
  LU_REQUIRE(bytes, "null argument");
  *bytes = L.total;
  return 0;
}

extern "C" int lu_augment_sequence(const float* dev_img, const float* dev_seg, const float* dev_contrast,
                                   const float* dev_brightness, const double* dev_coords, const lu_aug_params* ap,
                                   float* dev_img_out, float* dev_seg_out, void* dev_ws, size_t ws_bytes, void* stream) {
  LU_REQUIRE(dev_img && dev_seg && ap && dev_img_out && dev_seg_out && dev_ws, "null argument");
  AugLayout L;
  if (aug_layout(ap->frames, ap->H, ap->W, &L)) return 1;
  LU_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
  LU_REQUIRE(((uintptr_t)dev_ws & 255) == 0, "workspace must be 256-byte aligned");
  LU_REQUIRE(!ap->randomize || (dev_contrast && dev_brightness), "randomize needs the contrast / brightness arrays");
  LU_REQUIRE(!ap->elastic || dev_coords, "elastic augmentation needs the sampling coordinates (lu_elastic_coords)");
  LU_REQUIRE(ap->rot90 >= 0 && ap->rot90 <= 3, "rot90 must be 0..3");
  uint8_t* ws = (uint8_t*)dev_ws;
  LuAug q;
  memset(&q, 0, sizeof q);
  q.img = dev_img; q.seg = dev_seg; q.contrast = dev_contrast; q.brightness = dev_brightness; q.coords = dev_coords;
  q.frames = ap->frames; q.H = ap->H; q.W = ap->W; q.HW = ap->H * ap->W;
  q.randomize = ap->randomize; q.elastic = ap->elastic; q.flip0 = ap->flip0; q.flip1 = ap->flip1; q.rot90 = ap->rot90;
  {   // cv2.warpAffine's inversion of the 2x3 matrix (float64)
    double M[6];
    for (int i = 0; i < 6; ++i) M[i] = ap->affine[i];
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1.0 / D : 0.0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    for (int i = 0; i < 6; ++i) q.mi[i] = M[i];
  }
  q.sums = (double*)(ws + L.sums);
  q.adj = (float*)(ws + L.adj); q.warp = (float*)(ws + L.warp); q.wseg = (float*)(ws + L.wseg); q.wnv = (float*)(ws + L.wnv);
  q.tseg = (float*)(ws + L.tseg); q.tnv = (float*)(ws + L.tnv);
  q.out_img = dev_img_out; q.out_seg = dev_seg_out;
  const int64_t npix = (int64_t)q.frames * q.HW;
  LU_MEMSET(q.sums, 0, (size_t)q.frames * 16, stream);
  post_pf((int64_t)q.frames * ((q.HW + 1023) / 1024) * 32, stream, LuAugStats{q});
  post_pf(npix, stream, LuAugAdjust{q});
  if (q.elastic) {
    post_pf(npix, stream, LuAugWarp{q});
    post_pf(npix, stream, LuAugMap{q});
  } else {
    post_pf(npix, stream, LuAugPlace{q});
  }
  post_pf(npix, stream, LuAugSegFix{q});
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "augmentation launch failed: %s", cudaGetErrorString(e));
#endif
  return 0;
}

// _get_indices4elastic_transform (DataHandeling.py:183-193).  dev_rand: (2,H,W) float64 uniform [0,1) in the order the
// reference draws them (x field, then y field); dev_weights: the 2*lw+1 normalised Gaussian taps (host: lu_gauss_taps);
// dev_coords: (2,H,W) = (y + dy, x + dx); dev_tmp: (2,H,W) float64 scratch.
extern "C" int lu_elastic_coords(const double* dev_rand, const double* dev_weights, int32_t lw, int32_t H, int32_t W,
                                 double alpha, double* dev_tmp, double* dev_coords, void* stream) {
  LU_REQUIRE(dev_rand && dev_weights && dev_tmp && dev_coords && H > 0 && W > 0 && lw >= 0, "bad argument");
  LuGauss g;
  memset(&g, 0, sizeof g);
  g.H = H; g.W = W; g.lw = lw; g.fields = 2; g.w = dev_weights;
  g.in = dev_rand; g.out = dev_tmp; g.axis = 0; g.scale = 2.0; g.offset_in = -1.0; g.add_grid = 0; g.alpha = alpha;
  post_pf((int64_t)2 * H * W, stream, LuGaussPass{g});
  g.in = dev_tmp; g.out = dev_coords; g.axis = 1; g.scale = 1.0; g.offset_in = 0.0; g.add_grid = 1;
  post_pf((int64_t)2 * H * W, stream, LuGaussPass{g});
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "elastic field launch failed: %s", cudaGetErrorString(e));
#endif
  return 0;
}
