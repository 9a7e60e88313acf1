// Host side of the instance-labelling post-processing (lu_post.cuh): workspace layout, launch sequence, C-ABI.
// Included by lu_api.cu.

static int64_t g_post_launches = 0;

struct PostLayout {
  size_t cls, parA, parB, key, area, cc, lab, add, bflag, rowcnt, larea, present, newlab, bbox, info, slab, total;
  size_t zero_begin, zero_end;     // bflag .. info: cleared at the start of every call (bbox with the 0x7f pattern)
  int WB, HB, NB, KMAX, G;
};

static int post_layout(int frames, int H, int W, PostLayout* L) {
  LU_REQUIRE(frames > 0 && H > 0 && W > 0, "post-processing needs frames, H, W > 0");
  LU_REQUIRE((int64_t)H * W < (1ll << 30) && (int64_t)frames * H * W < (1ll << 40), "frame too large");
  const size_t HW = (size_t)H * W, N = (size_t)frames;
  L->WB = (W + 1) / 2;
  L->HB = (H + 1) / 2;
  L->NB = L->WB * L->HB;
  L->KMAX = L->NB + 1;
  int g = 1184 / frames;          // flood CTAs per frame: about 8 resident 128-thread CTAs per SM over the batch
  L->G = g < 4 ? 4 : (g > 148 ? 148 : g);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L->cls = take(N * HW);
  L->parA = take(N * HW * 4); L->parB = take(N * HW * 4); L->key = take(N * HW * 4); L->area = take(N * HW * 4);
  L->cc = take(N * HW * 4); L->lab = take(N * HW * 4); L->add = take(N * HW * 4);
  L->larea = take(N * L->KMAX * 4); L->newlab = take(N * L->KMAX * 4);
  L->zero_begin = off;
  L->bflag = take(N * L->NB); L->rowcnt = take(N * L->HB * 4); L->present = take(N * L->KMAX * 4); L->info = take(N * 4 * 4);
  L->zero_end = off;
  L->bbox = take(N * L->KMAX * 16);
  L->slab = take(N * LU_PP_SLABS * HW);
  L->total = off;
  return 0;
}

extern "C" int lu_post_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes) {
  PostLayout L;
  if (post_layout(frames, H, W, &L)) return 1;
  LU_REQUIRE(bytes, "null argument");
  *bytes = L.total;
  return 0;
}

extern "C" int lu_post_launch_count(int64_t* launches, int32_t reset) {
  if (launches) *launches = g_post_launches;
  if (reset) g_post_launches = 0;
  return 0;
}

template <class F>
static void post_pf(int64_t n, void* stream, F f) {
  if (n <= 0) return;
  g_post_launches++;
  lu_parallel_for_impl(n, stream, f);
}

extern "C" int lu_postprocess(const float* dev_softmax, int32_t frames, int32_t H, int32_t W, const lu_post_params* pp,
                   uint16_t* dev_labels, int32_t* dev_info, void* dev_ws, size_t ws_bytes, void* stream) {
  LU_REQUIRE(dev_softmax && pp && dev_labels && dev_ws, "null argument");
  PostLayout L;
  if (post_layout(frames, H, W, &L)) return 1;
  LU_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
  LU_REQUIRE(((uintptr_t)dev_ws & 255) == 0, "workspace must be 256-byte aligned");
  LU_REQUIRE(pp->fov >= 0 && (pp->fov == 0 || pp->fov < W), "FOV must be smaller than the frame width (the reference "
             "indexes column FOV, Inference2D.py:98)");
  LU_REQUIRE(pp->edge_d2_limit >= 0 && pp->edge_d2_limit <= 64 * 64, "edge distance out of range");
#ifndef LU_HOST_EMU
  {
    int dev = 0; cudaDeviceProp prop;
    cudaError_t e = cudaGetDevice(&dev);
    LU_REQUIRE(e == cudaSuccess, "cudaGetDevice: %s (no CUDA device: there is no CPU fallback)", cudaGetErrorString(e));
    static int checked_major = -1;
    if (checked_major < 0) {
      e = cudaGetDeviceProperties(&prop, dev);
      LU_REQUIRE(e == cudaSuccess, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
      checked_major = prop.major;
    }
    LU_REQUIRE(checked_major == 10, "this library is built for sm_100a (Blackwell B200)");
  }
#endif
  uint8_t* ws = (uint8_t*)dev_ws;
  LuPost q;
  memset(&q, 0, sizeof q);
  q.sm = dev_softmax; q.chw = pp->channels_first ? 1 : 0;
  q.N = frames; q.H = H; q.W = W; q.HW = H * W; q.WB = L.WB; q.HB = L.HB; q.NB = L.NB; q.KMAX = L.KMAX; q.G = L.G;
  q.edge_thresh = pp->edge_thresh; q.d2lim = pp->edge_d2_limit;
  q.rad = 0;
  while ((q.rad + 1) * (q.rad + 1) < q.d2lim) q.rad++;
  q.min_size = pp->min_cell_size; q.max_size = pp->max_cell_size; q.fov = pp->fov;
  q.cls = ws + L.cls;
  q.parA = (int32_t*)(ws + L.parA); q.parB = (int32_t*)(ws + L.parB); q.key = (int32_t*)(ws + L.key);
  q.area = (int32_t*)(ws + L.area); q.cc = (int32_t*)(ws + L.cc); q.lab = (int32_t*)(ws + L.lab); q.add = (int32_t*)(ws + L.add);
  q.bflag = ws + L.bflag; q.rowcnt = (int32_t*)(ws + L.rowcnt); q.larea = (int32_t*)(ws + L.larea); q.present = (int32_t*)(ws + L.present);
  q.newlab = (int32_t*)(ws + L.newlab); q.bbox = (int32_t*)(ws + L.bbox); q.info = (int32_t*)(ws + L.info);
  q.slab = ws + L.slab; q.out = dev_labels;

  LU_MEMSET(ws + L.zero_begin, 0, L.zero_end - L.zero_begin, stream);
  LU_MEMSET(ws + L.bbox, 0x7f, (size_t)frames * L.KMAX * 16, stream);
  const int64_t npix = (int64_t)frames * q.HW;
  post_pf(npix, stream, LuPpClassify{q, npix});
  post_pf(npix, stream, LuPpMergeBg{q});
  post_pf(npix, stream, LuPpFlattenBg{q});
  post_pf(npix, stream, LuPpFill{q, npix});
  post_pf(npix, stream, LuPpMergeFg{q});
  post_pf(npix, stream, LuPpFlattenFg{q});
  post_pf(npix, stream, LuPpMarkBlocks{q});
#ifdef LU_HOST_EMU
  int32_t sums[2], changed = 0;
  std::vector<uint8_t> small(LU_PP_SMEM_CROP);
  const LuCta one{0, 1};
  for (int n = 0; n < frames; ++n) lu_pp_rank_cta(one, q, n, sums);
#else
  lu_pp_rank_kernel<<<frames, 256, 0, (cudaStream_t)stream>>>(q);
#endif
  g_post_launches++;
  post_pf(npix, stream, LuPpRootLabel{q});
  post_pf(npix, stream, LuPpAssign{q});
  post_pf(npix, stream, LuPpEdges{q});
#ifdef LU_HOST_EMU
  for (int n = 0; n < frames; ++n)
    for (int g = 0; g < q.G; ++g) lu_pp_holes_cta(one, q, g, n, small.data(), &changed);
  for (int n = 0; n < frames; ++n)
    for (int g = 0; g < LU_PP_SLABS; ++g) lu_pp_holes_big_cta(one, q, g, n, small.data(), &changed);
  for (int n = 0; n < frames; ++n) lu_pp_holes_seq_cta(one, q, n, small.data(), &changed);
#else
  lu_pp_holes_kernel<<<dim3(q.G, frames), LU_PP_CTA, 0, (cudaStream_t)stream>>>(q);
  lu_pp_holes_big_kernel<<<dim3(LU_PP_SLABS, frames), 1024, 0, (cudaStream_t)stream>>>(q);
  lu_pp_holes_seq_kernel<<<frames, 256, 0, (cudaStream_t)stream>>>(q);
#endif
  g_post_launches += 3;
  post_pf(npix, stream, LuPpCombine{q});
#ifdef LU_HOST_EMU
  for (int n = 0; n < frames; ++n) lu_pp_relabel_cta(one, q, n, sums);
#else
  lu_pp_relabel_kernel<<<frames, 1024, 0, (cudaStream_t)stream>>>(q);
#endif
  g_post_launches++;
  post_pf(npix, stream, LuPpOutput{q});
  if (dev_info) LU_D2D(dev_info, q.info, (size_t)frames * 16, stream);
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "post-processing launch failed: %s", cudaGetErrorString(e));
#endif
  return 0;
}

// ---- SEG measure + accuracy (losses.py:29-88, train2D.py:97-102) ---------------------------------------------------------
struct SegLayout { size_t cls, parG, parS, areaG, areaS, score, keys, cnt, total; int cap; };
static int seg_layout(int frames, int H, int W, SegLayout* L) {
  LU_REQUIRE(frames > 0 && H > 0 && W > 0, "SEG measure needs frames, H, W > 0");
  LU_REQUIRE((int64_t)H * W < (1ll << 30), "frame too large");
  const size_t HW = (size_t)H * W, N = (size_t)frames;
  int cap = 64;
  while ((size_t)cap < HW) cap <<= 1;
  L->cap = cap;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L->cls = take(N * HW);
  L->parG = take(N * HW * 4); L->parS = take(N * HW * 4); L->areaG = take(N * HW * 4); L->areaS = take(N * HW * 4);
  L->score = take(N * HW * 4);
  L->keys = take(N * (size_t)cap * 8); L->cnt = take(N * (size_t)cap * 4);
  L->total = off;
  return 0;
}
extern "C" int lu_seg_workspace_bytes(int32_t frames, int32_t H, int32_t W, size_t* bytes) {
  SegLayout L;
  if (seg_layout(frames, H, W, &L)) return 1;
  LU_REQUIRE(bytes, "null argument");
  *bytes = L.total;
  return 0;
}
extern "C" int lu_seg_measure(const float* dev_labels, const float* dev_logits, int32_t frames, int32_t H, int32_t W,
                              int32_t channels_first, double* dev_result4, void* dev_ws, size_t ws_bytes, void* stream) {
  LU_REQUIRE(dev_labels && dev_logits && dev_result4 && dev_ws, "null argument");
  SegLayout L;
  if (seg_layout(frames, H, W, &L)) return 1;
  LU_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
  LU_REQUIRE(((uintptr_t)dev_ws & 255) == 0, "workspace must be 256-byte aligned");
  uint8_t* ws = (uint8_t*)dev_ws;
  LuSeg q;
  memset(&q, 0, sizeof q);
  q.labels = dev_labels; q.logits = dev_logits; q.chw = channels_first ? 1 : 0;
  q.N = frames; q.H = H; q.W = W; q.HW = H * W; q.cap = L.cap;
  q.cls = ws + L.cls;
  q.parG = (int32_t*)(ws + L.parG); q.parS = (int32_t*)(ws + L.parS);
  q.areaG = (int32_t*)(ws + L.areaG); q.areaS = (int32_t*)(ws + L.areaS);
  q.score = (float*)(ws + L.score); q.keys = (long long*)(ws + L.keys); q.cnt = (int32_t*)(ws + L.cnt);
  q.result = dev_result4;
  LU_MEMSET(ws + L.keys, 0xff, (size_t)frames * L.cap * 8, stream);
  LU_MEMSET(ws + L.cnt, 0, (size_t)frames * L.cap * 4, stream);
  LU_MEMSET(dev_result4, 0, 4 * sizeof(double), stream);
  const int64_t npix = (int64_t)frames * q.HW;
  post_pf(npix, stream, LuSegClassify{q, npix});
  post_pf(npix, stream, LuSegMerge{q});
  post_pf(npix, stream, LuSegFlatten{q});
  post_pf(npix, stream, LuSegPairs{q, npix});
  post_pf((int64_t)frames * L.cap, stream, LuSegScore{q});
  post_pf(npix, stream, LuSegReduce{q, npix});
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "SEG measure launch failed: %s", cudaGetErrorString(e));
#endif
  return 0;
}
