// Host side of liblstm_unet_b200.so: plan construction (layer graph, activation-staging tables, weight-packing
// descriptors, workspace layout), TMA tensor maps, kernel launches and the C-ABI declared in
// include/lstm_unet_b200.h.  Mirrors Networks.py:35-291 of the reference (structure only; all arithmetic is in
// lu_conv.cuh / lu_elem.cuh).
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/lstm_unet_b200.h"
#include "lu_conv.cuh"
#include "lu_elem.cuh"
#include "lu_train.cuh"
#include "lu_post.cuh"
#include "lu_aug.cuh"

#ifdef LU_HOST_EMU
#define LU_MEMSET(p, v, n, s) memset((p), (v), (n))
#define LU_H2D(d, s_, n, st) memcpy((d), (s_), (n))
#define LU_D2D(d, s_, n, st) memcpy((d), (s_), (n))
#else
#define LU_D2D(d, s_, n, st) cudaMemcpyAsync((d), (s_), (n), cudaMemcpyDeviceToDevice, (cudaStream_t)(st))
#define LU_MEMSET(p, v, n, s) cudaMemsetAsync((p), (v), (n), (cudaStream_t)(s))
#define LU_H2D(d, s_, n, st) cudaMemcpyAsync((d), (s_), (n), cudaMemcpyHostToDevice, (cudaStream_t)(st))
#endif

static thread_local std::string g_err;
#define LU_FAIL(...)                                   \
  do {                                                 \
    char buf_[512];                                    \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);         \
    g_err = buf_;                                      \
    return 1;                                          \
  } while (0)
#define LU_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) LU_FAIL(__VA_ARGS__); \
  } while (0)

#define LU_WG_MAX_TASKS 16384
static inline int ceil_to(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------------------------
struct ParamT {
  std::string name;
  int64_t shape[4];
  int rank;
  int64_t offset;   // in the flat fp32 buffer
  int64_t count;
  bool trainable;
};

struct ActBuf {      // NHWC bf16 activation buffer, channel layout [hi: cpad][lo: cpad] when planes == 2
  size_t off = 0;
  int frames = 0, H = 0, W = 0, creal = 0, cpad = 0, planes = 1;
  size_t bytes() const { return (size_t)frames * H * W * cpad * planes * 2; }
};

struct ConvIn {
  int buf = -1;        // index into acts, or -1 = image patches
  int creal = 0;
  int w_param = -1;    // kernel tensor holding this input's weights
  int c_base = 0;      // first input channel inside that tensor (concat offset)
};

struct ConvPlan {
  std::string name;
  int kind = LU_EPI_CONV;
  int k = 3, stride = 1;
  int n_in = 0;
  ConvIn in[2];
  int bias_param = -1, gamma = -1, beta = -1, mov_mean = -1, mov_var = -1;
  int cout = 0, BN = 0, n_tiles_n = 0, npad = 0;
  int Hin = 0, Win = 0, Hout = 0, Wout = 0;
  int out_buf = -1;            // ActBuf index of the activation output (-1: raw only)
  int raw_cpad = 0;
  size_t off_raw = 0;          // fp32 raw output (training BN input / logits), 0 = shared scratch
  bool has_bn = false;
  LuColMap cm{};
  std::vector<LuAStage> astages;
  std::vector<uint16_t> taps;
  std::vector<LuPackDesc> packs;
  int n_views = 0;
  LuSrcView views[LU_MAX_SRC];
  int view_buf[LU_MAX_SRC];    // ActBuf index (or -1 patches) behind each view
  size_t off_astages = 0, off_taps = 0, off_packs = 0, off_w = 0, off_bias = 0, off_scale = 0, off_shift = 0;
  size_t off_bscale = 0, off_bshift = 0, off_sums = 0, off_mom = 0, off_save_mean = 0, off_save_invstd = 0;
  int ktot = 0;
  int nA = 2, nB = 4, a_bytes = 0, b_bytes = 0, smem = 0, b_group = 1;
  bool b_resident = false;              // the whole weight panel of the (single) N tile stays in shared memory
  bool ptab_ok = false;                 // staging tables fit the kernel-parameter copies
  std::vector<LuAStage> pstages;        // stages with tap_begin remapped into the de-duplicated tap list
  std::vector<uint16_t> ptaps;
  double macs_per_frame = 0;
  // lstm only
  int F = 0, fpad = 0, level = 0, layer = 0;
  size_t off_hstate[2] = {0, 0}, off_cstate = 0, off_save_gates = 0, off_save_c = 0;
  int hseq_buf = -1;
  // training (cfg.train): data-gradient plans per input, packed-space weight-gradient maps, BPTT buffers
  int fwd = -1, fwd_in = 0;             // LU_EPI_GRAD plans: the forward conv / input they differentiate
  int oy_mul = 1, oy_add = 0, ox_mul = 1, ox_add = 0, OH = 0, OW = 0;
  int dz_buf = -1;                      // lstm: gradient wrt the gate pre-activations (frames,H,W,4*fpad)
  std::vector<int> dgrads[2];
  std::vector<uint16_t> kb_stage, kb_tap;
  int wg_cached_T[2] = {-1, -1}, wg_n_tasks[2] = {0, 0}, wg_n_ptasks[2] = {0, 0};   // weight-gradient task lists resident on the device
  size_t off_kb_stage = 0, off_kb_tap = 0, off_bwd_sums = 0, off_c_init = 0, off_dc = 0, off_wg_tasks = 0, off_wg_ptasks = 0, off_bwd_means = 0;
#ifndef LU_HOST_EMU
  CUtensorMap tmA[LU_MAX_SRC];
  CUtensorMap tmB, tmBh;
  CUtensorMap tmHstate[2];
#endif
};

struct UpStage {
  int src_buf = -1, dst_buf = -1;   // bilinear x2 from src to dst
};

struct lu_handle_s {
  lu_config cfg;
  int L = 0, pw = 0, planes = 1;
  int fmt = 0;                    // 16-bit operand format: 0 bf16, 1 fp16 (LU_PREC_FP16, inference handles only)
  int Hp = 0, Wp = 0, pad_y0 = 0, pad_x0 = 0;
  int lvlH[LU_MAX_LEVELS + 1], lvlW[LU_MAX_LEVELS + 1];
  std::vector<ParamT> params;
  int64_t n_elems = 0, n_train = 0;
  std::vector<ActBuf> acts;
  std::vector<ConvPlan> convs;
  // execution order
  std::vector<std::vector<int>> lstm_of_level, conv_of_level, conv_of_up;
  std::vector<UpStage> ups;       // per up block (src_buf = -1: no resize)
  int logits_conv = -1;
  int img_buf = -1;               // in_channels > 1: the reflect-padded image as an ordinary NHWC activation buffer
  int skip_in_buf = -1;           // LU_BLOCK_UP: the skip input
  int block_out_conv = -1;        // stand-alone blocks: the convolution whose output the block returns
  size_t off_patches = 0, off_raw_scratch = 0, raw_scratch_bytes = 0, off_logits_raw = 0;
  size_t ws_bytes = 0;
  uint8_t* ws = nullptr;
  float* dparams = nullptr;
  int hcur = 0;                   // which h-state ping-pong buffer is current (all layers flip together)
  int last_T = 0, last_training = 0;
  int64_t launches = 0;
  bool bound = false, packed = false;
  bool bn_fold_stale = false;     // a training forward moved the moving statistics: re-fold them before the next inference forward
  int num_sms = 148;
  TrainState tr;
#ifndef LU_HOST_EMU
  std::vector<CUtensorMap> acts_tm;   // cfg.train: dense 16x8-tile map of every activation buffer (dY operand of wgrad)
#endif
  std::vector<int> gidx;          // activation buffer -> its gradient twin (cfg.train)
  int g_logits_buf = -1;
  size_t tr_off_dwp = 0, tr_dwp_bytes = 0;
  // CUDA-graph replay of the inference forward for launch-bound shapes (Inference2D's B=1, T=1 frame loop): one
  // instantiated graph per (T, h ping-pong parity); inputs / outputs go through fixed staging buffers in the workspace
  int graph_mode = 0;
  bool graph_ok = false;
  size_t off_gx = 0, off_glogits = 0, off_gsoftmax = 0, g_x_bytes = 0, g_out_bytes = 0;
  struct FwdGraph { int T, hcur; int64_t launches;
#ifndef LU_HOST_EMU
    cudaGraphExec_t exec;
#endif
  };
  std::vector<FwdGraph> graphs;
#ifndef LU_HOST_EMU
  cudaStream_t cap_stream = nullptr;
#endif
  // data-parallel training: called on the host as soon as the launches that finalise the gradients of one block of
  // parameters (one Up / Down block = one contiguous range of the flat gradient buffer) have been enqueued
  lu_grad_bucket_fn bucket_fn = nullptr;
  void* bucket_user = nullptr;
  // synchronised BatchNorm: in-place sum over the ranks of a small fp64 device vector, enqueued on the compute stream
  lu_bn_sync_fn bn_sync_fn = nullptr;
  void* bn_sync_user = nullptr;
  int bn_sync_world = 1;
  // optional CUDA-event timing of every tensor-core launch, by kernel class (bench.py rooflines)
  bool time_on = false;
  size_t ev_used = 0;
  std::vector<int> ev_class;      // class of event pair i (events[2i], events[2i+1])
#ifndef LU_HOST_EMU
  std::vector<cudaEvent_t> events;
#endif
  // cudaFuncSetAttribute is per device and a handle is bound to one device: remembered per handle, not per process
  bool attr_set[12] = {false, false, false, false, false, false, false, false, false, false, false, false};
  bool wg_attr_set = false;
};

#ifndef LU_HOST_EMU
static void time_begin(lu_handle_s* h, int cls, void* stream) {
  if (!h->time_on) return;
  while (h->events.size() < h->ev_used + 2) { cudaEvent_t ev; cudaEventCreate(&ev); h->events.push_back(ev); }
  if (h->ev_class.size() < h->ev_used / 2 + 1) h->ev_class.resize(h->ev_used / 2 + 1);
  h->ev_class[h->ev_used / 2] = cls;
  cudaEventRecord(h->events[h->ev_used], (cudaStream_t)stream);
}
static void time_end(lu_handle_s* h, void* stream) {
  if (!h->time_on) return;
  cudaEventRecord(h->events[h->ev_used + 1], (cudaStream_t)stream);
  h->ev_used += 2;
}
#endif

static void train_layout(lu_handle_s* h, size_t& off);
static int build_train_plan(lu_handle_s* h);
static void train_upload(lu_handle_s* h, void* stream);
static void train_destroy(lu_handle_s* h);
extern "C" int lu_lstm_flops(lu_handle h, int32_t T, double* flops);

static int find_param(lu_handle_s* h, const std::string& name) {
  for (size_t i = 0; i < h->params.size(); ++i)
    if (h->params[i].name == name) return (int)i;
  return -1;
}

// ------------------------------------------------------------------------------------------------------------------
// parameter layout: Keras variables, trainable first (same names as oracle/lstm_unet_oracle.py build_param_specs)
// ------------------------------------------------------------------------------------------------------------------
static void add_param(std::vector<ParamT>& v, const std::string& name, std::vector<int64_t> shape, bool trainable) {
  ParamT p;
  p.name = name;
  p.rank = (int)shape.size();
  p.count = 1;
  for (int i = 0; i < 4; ++i) p.shape[i] = i < p.rank ? shape[i] : 1;
  for (int i = 0; i < p.rank; ++i) p.count *= shape[i];
  p.trainable = trainable;
  p.offset = 0;
  v.push_back(p);
}

static void build_params(lu_handle_s* h) {
  const lu_config& c = h->cfg;
  std::vector<ParamT> all;
  char nm[128];
  int cin = c.in_channels;
  std::vector<int> skip_ch;
  const bool only_up = c.block_kind == LU_BLOCK_UP, only_down = c.block_kind == LU_BLOCK_DOWN;
  if (only_up) skip_ch.push_back(c.skip_channels);
  for (int l = 0; l < h->L && !only_up; ++l) {
    skip_ch.push_back(cin);
    for (int j = 0; j < c.n_lstm[l]; ++j) {
      const int k = c.lstm_k[l][j], f = c.lstm_f[l][j];
      snprintf(nm, sizeof nm, "DownLayers/%d/ConvLSTM/%d/", l, j);
      add_param(all, std::string(nm) + "kernel", {k, k, cin, 4 * f}, true);
      add_param(all, std::string(nm) + "recurrent_kernel", {k, k, f, 4 * f}, true);
      add_param(all, std::string(nm) + "bias", {4 * f}, true);
      cin = f;
    }
    for (int j = 0; j < c.n_down[l]; ++j) {
      const int k = c.down_k[l][j], f = c.down_f[l][j];
      snprintf(nm, sizeof nm, "DownLayers/%d/Conv/%d/", l, j);
      add_param(all, std::string(nm) + "kernel", {k, k, cin, f}, true);
      add_param(all, std::string(nm) + "bias", {f}, true);
      snprintf(nm, sizeof nm, "DownLayers/%d/BN/%d/", l, j);
      add_param(all, std::string(nm) + "gamma", {f}, true);
      add_param(all, std::string(nm) + "beta", {f}, true);
      add_param(all, std::string(nm) + "moving_mean", {f}, false);
      add_param(all, std::string(nm) + "moving_variance", {f}, false);
      cin = f;
    }
  }
  for (int u = 0; u < h->L && !only_down; ++u) {
    cin += skip_ch[h->L - 1 - u];
    for (int j = 0; j < c.n_up[u]; ++j) {
      const int k = c.up_k[u][j], f = c.up_f[u][j];
      snprintf(nm, sizeof nm, "UpLayers/%d/Conv/%d/", u, j);
      add_param(all, std::string(nm) + "kernel", {k, k, cin, f}, true);
      add_param(all, std::string(nm) + "bias", {f}, true);
      // the BN / LeakyReLU objects of the logits convolution are never called and own no variables (Networks.py:148-149)
      const bool is_logits = (u == h->L - 1) && (j == c.n_up[u] - 1) && (!only_up || c.return_logits);
      if (!is_logits) {
        snprintf(nm, sizeof nm, "UpLayers/%d/BN/%d/", u, j);
        add_param(all, std::string(nm) + "gamma", {f}, true);
        add_param(all, std::string(nm) + "beta", {f}, true);
        add_param(all, std::string(nm) + "moving_mean", {f}, false);
        add_param(all, std::string(nm) + "moving_variance", {f}, false);
      }
      cin = f;
    }
  }
  int64_t off = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (auto& p : all)
      if (p.trainable == (pass == 0)) {
        p.offset = off;
        off += p.count;
        h->params.push_back(p);
        if (pass == 0) h->n_train = off;
      }
  h->n_elems = off;
}

// ------------------------------------------------------------------------------------------------------------------
// activation-staging tables + weight packing descriptors for one convolution
// ------------------------------------------------------------------------------------------------------------------
static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

static void tf_same(int in, int k, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int total = (*out - 1) * s + k - in;
  if (total < 0) total = 0;
  *before = total / 2;
}

static int finish_tables(lu_handle_s* h, ConvPlan& cv);

static int build_tables(lu_handle_s* h, ConvPlan& cv) {
  const bool x3 = h->cfg.precision == LU_PREC_BF16X3;
  const bool halo = h->cfg.a_mode == LU_AMODE_HALO;
  const int k = cv.k;
  cv.n_views = 0;
  for (int i = 0; i < cv.n_in; ++i) {
    const ConvIn& in = cv.in[i];
    const ParamT& wp = h->params[in.w_param];
    const int cin_total = (int)wp.shape[2], cout_total = (int)wp.shape[3];
    LuSrcView v;
    memset(&v, 0, sizeof v);
    cv.macs_per_frame += (double)k * k * in.creal * cv.cout;      // x Ho*Wo applied by the caller
    const int vi = cv.n_views++;
    LU_REQUIRE(vi < LU_MAX_SRC, "too many source views");
    cv.view_buf[vi] = in.buf;
    if (in.buf < 0) {
      // ---- 1-channel image as pw x pw patches: a 1x1 tap over 64 "channels"
      LU_REQUIRE(cv.stride == 1, "image patches feed stride-1 convolutions only");
      v.dimC = 64; v.dimW = h->Wp; v.dimP = 1; v.dimH = h->Hp; v.dimN = h->cfg.batch * h->cfg.max_t;
      v.sw = 64; v.sh = (int64_t)h->Wp * 64; v.sp = v.sh; v.sn = (int64_t)h->Hp * h->Wp * 64;
      v.rows = LU_TILE_H; v.pitch = LU_TILE_W; v.frame_mul = 1; v.frame_add = 0;
      LuAStage st; memset(&st, 0, sizeof st);
      st.src = (uint8_t)vi; st.tap_begin = (uint32_t)cv.taps.size(); st.ntaps = x3 ? 2 : 1;
      cv.astages.push_back(st);
      for (int part = 0; part < (x3 ? 2 : 1); ++part) {
        cv.taps.push_back(0);
        LuPackDesc d; memset(&d, 0, sizeof d);
        d.w_off = wp.offset; d.c_base = in.c_base; d.cin_total = cin_total; d.cout_total = cout_total;
        d.kind = 1; d.k = (int8_t)k; d.pw = (int8_t)h->pw; d.patch_x3 = x3; d.wpart = (int8_t)part;
        d.patch_hi_only = (int8_t)(part == 1);
        cv.packs.push_back(d);
      }
      cv.views[vi] = v;
      continue;
    }
    const ActBuf& ab = h->acts[in.buf];
    const int ctot = ab.cpad * ab.planes;
    const int nchunks = ab.cpad / LU_KBLK;
    v.dimN = ab.frames; v.frame_mul = 1; v.frame_add = 0;
    // tap geometry: source row = s*oy + r with r = ky - pad_before;  r = s*a + par
    int ho, wo, pt, pl;
    tf_same(ab.H, k, cv.stride, &ho, &pt);
    tf_same(ab.W, k, cv.stride, &wo, &pl);
    const int s = cv.stride;
    int a_min = 1 << 20, a_max = -(1 << 20), b_min = 1 << 20, b_max = -(1 << 20);
    for (int ky = 0; ky < k; ++ky) { int a = floordiv(ky - pt, s); a_min = a < a_min ? a : a_min; a_max = a > a_max ? a : a_max; }
    for (int kx = 0; kx < k; ++kx) { int b = floordiv(kx - pl, s); b_min = b < b_min ? b : b_min; b_max = b > b_max ? b : b_max; }
    if (s == 1) {
      v.dimC = ctot; v.dimW = ab.W; v.dimP = 1; v.dimH = ab.H;
      v.sw = ctot; v.sh = (int64_t)ab.W * ctot; v.sp = v.sh; v.sn = (int64_t)ab.H * ab.W * ctot;
    } else {
      LU_REQUIRE(s == 2 && ab.H % 2 == 0 && ab.W % 2 == 0, "stride-2 convolution needs even input size");
      v.dimC = 2 * ctot; v.dimW = ab.W / 2; v.dimP = 2; v.dimH = ab.H / 2;
      v.sw = 2 * ctot; v.sp = (int64_t)ab.W * ctot; v.sh = 2 * (int64_t)ab.W * ctot; v.sn = (int64_t)ab.H * ab.W * ctot;
    }
    if (halo) {
      // exact halo window: (16 + kh - 1) rows of (8 + kw - 1) pixels.  The UMMA descriptor's 8-row group stride is
      // pitch*128 bytes; it need not be a multiple of the 1024-byte swizzle atom because the swizzle is a function
      // of the shared-memory address bits (see make_desc in lu_conv.cuh).
      v.rows = LU_TILE_H + (a_max - a_min);
      v.pitch = LU_TILE_W + (b_max - b_min);
    } else {
      v.rows = LU_TILE_H; v.pitch = LU_TILE_W;
    }
    cv.views[vi] = v;
    for (int par_y = 0; par_y < s; ++par_y)
      for (int par_x = 0; par_x < s; ++par_x) {
        std::vector<int> kys, kxs;
        for (int ky = 0; ky < k; ++ky) if (((ky - pt) % s + s) % s == par_y) kys.push_back(ky);
        for (int kx = 0; kx < k; ++kx) if (((kx - pl) % s + s) % s == par_x) kxs.push_back(kx);
        if (kys.empty() || kxs.empty()) continue;
        for (int ch = 0; ch < nchunks; ++ch) {
          int nvalid = in.creal - ch * LU_KBLK;
          nvalid = nvalid < 0 ? 0 : (nvalid > LU_KBLK ? LU_KBLK : nvalid);
          if (nvalid == 0) continue;                      // all-padding chunk: contributes nothing
          for (int aplane = 0; aplane < (x3 ? 2 : 1); ++aplane) {
            const int nw = (x3 && aplane == 0) ? 2 : 1;   // a_hi pairs with w_hi and w_lo; a_lo with w_hi only
            const int ccoord = par_x * ctot + aplane * ab.cpad + ch * LU_KBLK;
            LuAStage st; memset(&st, 0, sizeof st);
            st.src = (uint8_t)vi; st.plane = (uint8_t)par_y; st.c = ccoord;
            if (halo) {
              st.dy = (int16_t)a_min; st.dx = (int16_t)b_min;
              st.tap_begin = (uint32_t)cv.taps.size(); st.ntaps = 0;
            }
            for (int wpart = 0; wpart < nw; ++wpart)
              for (int ky : kys)
                for (int kx : kxs) {
                  const int a = floordiv(ky - pt, s), b = floordiv(kx - pl, s);
                  LuPackDesc d; memset(&d, 0, sizeof d);
                  d.w_off = wp.offset; d.tap_off = (ky * k + kx) * cin_total * cout_total;
                  d.c_base = in.c_base + ch * LU_KBLK; d.n_valid = nvalid;
                  d.cin_total = cin_total; d.cout_total = cout_total; d.wpart = (int8_t)wpart; d.kind = 0;
                  cv.packs.push_back(d);
                  if (halo) {
                    cv.taps.push_back((uint16_t)((a - a_min) * v.pitch + (b - b_min)));
                    st.ntaps++;
                  } else {
                    LuAStage s1 = st;
                    s1.dy = (int16_t)a; s1.dx = (int16_t)b; s1.ntaps = 1; s1.tap_begin = (uint32_t)cv.taps.size();
                    cv.taps.push_back(0);
                    cv.astages.push_back(s1);
                  }
                }
            if (halo) cv.astages.push_back(st);
          }
        }
      }
  }
  return finish_tables(h, cv);
}

// K extent, kernel-parameter table copies and shared-memory pipeline shape of a conv whose tables are complete
static int finish_tables(lu_handle_s* h, ConvPlan& cv) {
  (void)h;
  cv.ktot = (int)cv.packs.size() * LU_KBLK;
  // kernel-parameter copies of the tables: identical tap lists are stored once
  cv.pstages = cv.astages; cv.ptaps.clear();
  for (auto& st : cv.pstages) {
    const uint16_t* lst = cv.taps.data() + st.tap_begin;
    int found = -1;
    for (int b = 0; b + (int)st.ntaps <= (int)cv.ptaps.size() && found < 0; ++b) {
      bool same = true;
      for (int t = 0; t < st.ntaps && same; ++t) same = cv.ptaps[b + t] == lst[t];
      if (same) found = b;
    }
    if (found < 0) { found = (int)cv.ptaps.size(); cv.ptaps.insert(cv.ptaps.end(), lst, lst + st.ntaps); }
    st.tap_begin = (uint32_t)found;
  }
  cv.ptab_ok = (int)cv.pstages.size() <= LU_PT_STAGES && (int)cv.ptaps.size() <= LU_PT_TAPS;
  // shared-memory pipeline shape
  cv.a_bytes = 0;
  for (int i = 0; i < cv.n_views; ++i) {
    const int b = (int)align_up((size_t)cv.views[i].rows * cv.views[i].pitch * 128, 1024);
    cv.a_bytes = b > cv.a_bytes ? b : cv.a_bytes;
  }
  // weight stage = b_group K blocks (about 16-32 KB): one barrier round trip per group instead of per tap
  cv.b_group = 256 / cv.BN;
  if (cv.b_group > 8) cv.b_group = 8;
  if (cv.b_group > (int)cv.packs.size()) cv.b_group = (int)cv.packs.size();
  cv.b_bytes = cv.b_group * cv.BN * 128;
  const int budget = 232448 - 1024 - 512 - 6144;   // alignment slack, barriers, epilogue constants
  // weight stages first (one is consumed per tap: >= 4 in flight), then as many activation windows as fit
  const int ngrp = ((int)cv.packs.size() + cv.b_group - 1) / cv.b_group;
  int nb_min = ngrp < 4 ? (ngrp < 2 ? 2 : ngrp) : 4;
  cv.nA = (budget - nb_min * cv.b_bytes) / cv.a_bytes;
  if (cv.nA > 6) cv.nA = 6;
  LU_REQUIRE(cv.nA >= 2, "shared memory budget exceeded for %s", cv.name.c_str());
  cv.nB = (budget - cv.nA * cv.a_bytes) / cv.b_bytes;
  if (cv.nB > 8) cv.nB = 8;
  LU_REQUIRE(cv.nB >= 2, "shared memory budget exceeded for %s", cv.name.c_str());
  // Narrow convolutions (one N tile, small K -- the 32/64-channel decoder tail, the logits conv): every tile would
  // re-stream the same few tens of KB of weights from L2; instead the panel is loaded once per CTA and stays resident,
  // and the shared memory it does not need goes to deeper activation prefetch.
  cv.b_resident = false;
  if (cv.kind != LU_EPI_LSTM && cv.n_tiles_n == 1 && ngrp <= 8 && (size_t)ngrp * cv.b_bytes <= 96 * 1024) {
    const int na = (budget - ngrp * cv.b_bytes) / cv.a_bytes;
    if (na >= 3) { cv.b_resident = true; cv.nB = ngrp; cv.nA = na > 8 ? 8 : na; }
  }
  cv.smem = cv.nA * cv.a_bytes + cv.nB * cv.b_bytes + 1024 + 512 + 6144;
  return 0;
}

static int pick_bn(int cout) {
  if (cout > 128) return 256;
  if (cout > 64) return 128;
  if (cout > 32) return 64;
  if (cout > 16) return 32;
  return 16;
}

static int new_act(lu_handle_s* h, int frames, int H, int W, int creal) {
  ActBuf a;
  a.frames = frames; a.H = H; a.W = W; a.creal = creal; a.cpad = ceil_to(creal, LU_KBLK); a.planes = h->planes;
  h->acts.push_back(a);
  return (int)h->acts.size() - 1;
}

// ------------------------------------------------------------------------------------------------------------------
// layer graph (Networks.py:179-254)
// ------------------------------------------------------------------------------------------------------------------
static int build_plan(lu_handle_s* h) {
  const lu_config& c = h->cfg;
  const int L = h->L, B = c.batch, N = c.batch * c.max_t;
  const bool net = c.block_kind == LU_BLOCK_NET, only_down = c.block_kind == LU_BLOCK_DOWN, only_up = c.block_kind == LU_BLOCK_UP;
  const int ts = 1 << (L - 1);                                  // total_stride (Networks.py:197-199)
  const int minpad = (net && c.pad_image) ? ts : 0;             // Networks.py:210; a block on its own never pads
  h->pad_y0 = minpad; h->pad_x0 = minpad;
  const int pad_y1 = net ? minpad + (ts - c.height % ts) % ts : 0, pad_x1 = net ? minpad + (ts - c.width % ts) % ts : 0;
  LU_REQUIRE(pad_y1 < c.height && pad_x1 < c.width && minpad < c.height && minpad < c.width,
             "REFLECT padding needs pad < image size (H=%d W=%d pad=%d/%d)", c.height, c.width, pad_y1, pad_x1);
  h->Hp = c.height + minpad + pad_y1; h->Wp = c.width + minpad + pad_x1;
  // stride of the first convolution of each encoder block: 2 except in the last one (Networks.py:195-196), or what the
  // caller of a stand-alone DownBlock2D passed
  int lvl_stride[LU_MAX_LEVELS];
  for (int l = 0; l < L; ++l) lvl_stride[l] = only_down ? c.block_stride : (l < L - 1 ? 2 : 1);
  h->lvlH[0] = h->Hp; h->lvlW[0] = h->Wp;
  for (int l = 0; l < L; ++l) {
    LU_REQUIRE(h->lvlH[l] % lvl_stride[l] == 0 && h->lvlW[l] % lvl_stride[l] == 0,
               "a stride-%d block needs even input sizes, got %d x %d", lvl_stride[l], h->lvlH[l], h->lvlW[l]);
    h->lvlH[l + 1] = h->lvlH[l] / lvl_stride[l]; h->lvlW[l + 1] = h->lvlW[l] / lvl_stride[l];
  }
  LU_REQUIRE(c.in_channels >= 1 && c.in_channels <= 4096, "bad in_channels=%d", c.in_channels);
  if (c.in_channels == 1 && net) {
    // the CTC path: the 1-channel image is expanded into pw x pw patches (64 "channels", one 1x1 tap);
    // patch window = the largest kernel that reads the image directly
    h->pw = c.lstm_k[0][0];
    if (c.up_k[L - 1][0] > h->pw) h->pw = c.up_k[L - 1][0];
    LU_REQUIRE(h->pw * h->pw <= (h->planes == 2 ? 32 : 64), "kernel size %d too large for the image patch path", h->pw);
  }

  h->lstm_of_level.assign(L, {}); h->conv_of_level.assign(L, {}); h->conv_of_up.assign(L, {});
  h->ups.assign(L, UpStage());
  char nm[128];
  int cur_buf = -1, cur_c = c.in_channels;      // -1 = image patches
  if (c.in_channels > 1 || !net) {              // multi-channel images (the reference unit_test feeds 3, Networks.py:266):
    h->img_buf = new_act(h, N, h->Hp, h->Wp, c.in_channels);      // generic path, the image is just another NHWC source
    cur_buf = h->img_buf;
  }
  std::vector<int> skip_buf, skip_c;
  if (only_up) {                                // UpBlock2D alone: the skip tensor is the second input (Networks.py:142)
    LU_REQUIRE(c.skip_channels >= 1 && c.skip_channels <= 4096, "bad skip_channels=%d", c.skip_channels);
    h->lvlH[0] = h->Hp * c.block_stride; h->lvlW[0] = h->Wp * c.block_stride;
    h->skip_in_buf = new_act(h, N, h->lvlH[0], h->lvlW[0], c.skip_channels);
    skip_buf.push_back(h->skip_in_buf); skip_c.push_back(c.skip_channels);
  }
  for (int l = 0; l < L && !only_up; ++l) {
    const int H = h->lvlH[l], W = h->lvlW[l];
    skip_buf.push_back(cur_buf); skip_c.push_back(cur_c);
    for (int j = 0; j < c.n_lstm[l]; ++j) {
      ConvPlan cv;
      snprintf(nm, sizeof nm, "DownLayers/%d/ConvLSTM/%d", l, j);
      cv.name = nm; cv.kind = LU_EPI_LSTM; cv.k = c.lstm_k[l][j]; cv.stride = 1; cv.level = l; cv.layer = j;
      cv.F = c.lstm_f[l][j]; cv.fpad = ceil_to(cv.F, LU_KBLK); cv.cout = 4 * cv.F;
      LU_REQUIRE(cv.k % 2 == 1, "ConvLSTM kernel size must be odd");
      cv.BN = 256; cv.n_tiles_n = cv.fpad / 64; cv.npad = cv.n_tiles_n * 256;
      cv.cm.kind = LU_COL_LSTM; cv.cm.F = cv.F; cv.cm.ch_tile = 64; cv.cm.n_real = 4 * cv.F;
      cv.Hin = cv.Hout = H; cv.Win = cv.Wout = W;
      cv.hseq_buf = new_act(h, N, H, W, cv.F);
      cv.out_buf = cv.hseq_buf;
      cv.n_in = 2;
      cv.in[0].buf = cur_buf; cv.in[0].creal = cur_c; cv.in[0].c_base = 0;
      cv.in[0].w_param = find_param(h, std::string(nm) + "/kernel");
      cv.in[1].buf = cv.hseq_buf; cv.in[1].creal = cv.F; cv.in[1].c_base = 0;
      cv.in[1].w_param = find_param(h, std::string(nm) + "/recurrent_kernel");
      cv.bias_param = find_param(h, std::string(nm) + "/bias");
      if (build_tables(h, cv)) return 1;
      cv.macs_per_frame *= (double)H * W;
      h->lstm_of_level[l].push_back((int)h->convs.size());
      h->convs.push_back(cv);
      cur_buf = cv.hseq_buf; cur_c = cv.F;
    }
    for (int j = 0; j < c.n_down[l]; ++j) {
      ConvPlan cv;
      snprintf(nm, sizeof nm, "DownLayers/%d/Conv/%d", l, j);
      cv.name = nm; cv.k = c.down_k[l][j]; cv.stride = (j == 0) ? lvl_stride[l] : 1;
      const int Hi = (j == 0) ? H : h->lvlH[l + 1], Wi = (j == 0) ? W : h->lvlW[l + 1];
      cv.Hin = Hi; cv.Win = Wi; cv.Hout = h->lvlH[l + 1]; cv.Wout = h->lvlW[l + 1];
      cv.cout = c.down_f[l][j]; cv.BN = pick_bn(cv.cout); cv.npad = ceil_to(cv.cout, cv.BN); cv.n_tiles_n = cv.npad / cv.BN;
      cv.cm.kind = LU_COL_IDENTITY; cv.cm.n_real = cv.cout;
      cv.out_buf = new_act(h, N, cv.Hout, cv.Wout, cv.cout);
      cv.n_in = 1; cv.in[0].buf = cur_buf; cv.in[0].creal = cur_c; cv.in[0].c_base = 0;
      cv.in[0].w_param = find_param(h, std::string(nm) + "/kernel");
      cv.bias_param = find_param(h, std::string(nm) + "/bias");
      snprintf(nm, sizeof nm, "DownLayers/%d/BN/%d", l, j);
      cv.has_bn = true;
      cv.gamma = find_param(h, std::string(nm) + "/gamma"); cv.beta = find_param(h, std::string(nm) + "/beta");
      cv.mov_mean = find_param(h, std::string(nm) + "/moving_mean"); cv.mov_var = find_param(h, std::string(nm) + "/moving_variance");
      LU_REQUIRE(cur_buf >= 0, "a ConvLSTM must precede the first convolution of level 0");
      if (build_tables(h, cv)) return 1;
      cv.macs_per_frame *= (double)cv.Hout * cv.Wout;
      h->conv_of_level[l].push_back((int)h->convs.size());
      if (only_down) h->block_out_conv = (int)h->convs.size();
      h->convs.push_back(cv);
      cur_buf = cv.out_buf; cur_c = cv.cout;
    }
  }
  for (int u = 0; u < L && !only_down; ++u) {
    const int sl = L - 1 - u;                        // skip index (skip list reversed, Networks.py:242)
    const int H = h->lvlH[sl], W = h->lvlW[sl];
    int up_buf = cur_buf;
    if (only_up ? c.block_stride == 2 : u > 0) {     // up_factor 2 except for the first block (Networks.py:202)
      up_buf = new_act(h, N, H, W, cur_c);
      h->ups[u].src_buf = cur_buf; h->ups[u].dst_buf = up_buf;
    }
    LU_REQUIRE(h->acts[up_buf].H == H && h->acts[up_buf].W == W, "up path shape mismatch at block %d", u);
    for (int j = 0; j < c.n_up[u]; ++j) {
      ConvPlan cv;
      snprintf(nm, sizeof nm, "UpLayers/%d/Conv/%d", u, j);
      cv.name = nm; cv.k = c.up_k[u][j]; cv.stride = 1;
      cv.Hin = cv.Hout = H; cv.Win = cv.Wout = W;
      cv.cout = c.up_f[u][j];
      const bool is_logits = (u == L - 1) && (j == c.n_up[u] - 1) && (!only_up || c.return_logits);
      cv.BN = pick_bn(cv.cout); cv.npad = ceil_to(cv.cout, cv.BN); cv.n_tiles_n = cv.npad / cv.BN;
      cv.cm.kind = LU_COL_IDENTITY; cv.cm.n_real = cv.cout;
      const int wparam = find_param(h, std::string(nm) + "/kernel");
      if (j == 0) {                                  // concat([upsampled, skip]) (Networks.py:145)
        cv.n_in = 2;
        cv.in[0].buf = up_buf; cv.in[0].creal = cur_c; cv.in[0].c_base = 0; cv.in[0].w_param = wparam;
        cv.in[1].buf = skip_buf[sl]; cv.in[1].creal = skip_c[sl]; cv.in[1].c_base = cur_c; cv.in[1].w_param = wparam;
      } else {
        cv.n_in = 1; cv.in[0].buf = cur_buf; cv.in[0].creal = cur_c; cv.in[0].c_base = 0; cv.in[0].w_param = wparam;
      }
      cv.bias_param = find_param(h, std::string(nm) + "/bias");
      if (!is_logits) {
        cv.out_buf = new_act(h, N, H, W, cv.cout);
        snprintf(nm, sizeof nm, "UpLayers/%d/BN/%d", u, j);
        cv.has_bn = true;
        cv.gamma = find_param(h, std::string(nm) + "/gamma"); cv.beta = find_param(h, std::string(nm) + "/beta");
        cv.mov_mean = find_param(h, std::string(nm) + "/moving_mean"); cv.mov_var = find_param(h, std::string(nm) + "/moving_variance");
      } else {
        cv.out_buf = -1;
        h->logits_conv = (int)h->convs.size();
      }
      if (build_tables(h, cv)) return 1;
      cv.macs_per_frame *= (double)H * W;
      h->conv_of_up[u].push_back((int)h->convs.size());
      if (only_up) h->block_out_conv = (int)h->convs.size();
      h->convs.push_back(cv);
      cur_buf = cv.out_buf; cur_c = cv.cout;
    }
  }
  LU_REQUIRE(!net || h->logits_conv >= 0, "network has no output convolution");
  LU_REQUIRE(net || h->block_out_conv >= 0, "the block has no convolution");
  (void)B;
  if (c.train && build_train_plan(h)) return 1;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------------------------
static void layout_workspace(lu_handle_s* h) {
  const lu_config& c = h->cfg;
  const int B = c.batch, N = c.batch * c.max_t;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  h->off_patches = take(h->img_buf < 0 ? (size_t)N * h->Hp * h->Wp * 64 * 2 : 1024);
  for (auto& a : h->acts) a.off = take(a.bytes());
  h->raw_scratch_bytes = 0;
  for (auto& cv : h->convs) {
    cv.off_astages = take(cv.astages.size() * sizeof(LuAStage));
    cv.off_taps = take(cv.taps.size() * sizeof(uint16_t));
    cv.off_packs = take(cv.packs.size() * sizeof(LuPackDesc));
    cv.off_w = take((size_t)cv.npad * cv.ktot * 2);
    cv.off_bias = take((size_t)cv.npad * 4);
    cv.off_scale = take((size_t)cv.npad * 4);
    cv.off_shift = take((size_t)cv.npad * 4);
    if (cv.kind == LU_EPI_GRAD) continue;
    if (cv.kind == LU_EPI_LSTM) {
      const size_t px = (size_t)B * cv.Hout * cv.Wout;
      cv.off_hstate[0] = take(px * cv.fpad * h->planes * 2);
      cv.off_hstate[1] = take(px * cv.fpad * h->planes * 2);
      cv.off_cstate = take(px * cv.fpad * 4);
      if (c.train) {
        const size_t pn = (size_t)N * cv.Hout * cv.Wout;
        cv.off_save_gates = take(pn * 4 * cv.fpad * h->planes * 2);
        cv.off_save_c = take(pn * cv.fpad * 4);
      }
    } else {
      // fp32 pre-activation output: all packed columns for a BN conv, the 16-column chunks that hold real channels for
      // the logits conv (3 classes: 64 B per pixel instead of 256 B)
      cv.raw_cpad = cv.has_bn ? cv.npad : ceil_to(cv.cout, 16);
      const size_t rb = (size_t)N * cv.Hout * cv.Wout * cv.raw_cpad * 4;
      cv.off_bscale = take((size_t)cv.npad * 4);
      cv.off_bshift = take((size_t)cv.npad * 4);
      cv.off_sums = take((size_t)cv.npad * 2 * 8);
      cv.off_mom = take((size_t)cv.npad * 3 * 8);
      cv.off_save_mean = take((size_t)cv.npad * 4);
      cv.off_save_invstd = take((size_t)cv.npad * 4);
      if (!cv.has_bn) cv.off_raw = take(rb);                     // logits
      else if (c.train) cv.off_raw = take(rb);                   // kept for backward
      else { cv.off_raw = (size_t)-1; if (rb > h->raw_scratch_bytes) h->raw_scratch_bytes = rb; }
    }
  }
  h->off_raw_scratch = take(h->raw_scratch_bytes ? h->raw_scratch_bytes : 16);
  for (auto& cv : h->convs)
    if (cv.off_raw == (size_t)-1) cv.off_raw = h->off_raw_scratch;
  train_layout(h, off);
  h->graph_ok = !c.train && N <= 8 && c.block_kind == LU_BLOCK_NET;
  if (h->graph_ok) {
    h->g_x_bytes = (size_t)N * c.in_channels * c.height * c.width * 4;
    h->g_out_bytes = (size_t)N * h->convs[h->logits_conv].cout * c.height * c.width * 4;
    h->off_gx = take(h->g_x_bytes);
    h->off_glogits = take(h->g_out_bytes);
    h->off_gsoftmax = take(h->g_out_bytes);
  }
  h->ws_bytes = off;
}

// ------------------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------------------
template <class F>
static void pf(lu_handle_s* h, int64_t n, void* stream, F f) {
  if (n <= 0) return;
  h->launches++;
  lu_parallel_for_impl(n, stream, f);
}
// row-loop launch: `groups` groups of 8 channels x npix pixels (lu_elem.cuh)
template <class F>
static void rows(lu_handle_s* h, int64_t npix, int groups, void* stream, F f) {
  if (npix <= 0 || groups <= 0) return;
  h->launches++;
  lu_rows_impl(npix, groups, stream, f);
}

#ifndef LU_HOST_EMU
typedef CUresult (*PFN_tmEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmEncodeTiled g_encode = nullptr;

static int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  LU_REQUIRE(e == cudaSuccess && q == cudaDriverEntryPointSuccess && fn, "cuTensorMapEncodeTiled unavailable: %s",
             cudaGetErrorString(e));
  g_encode = (PFN_tmEncodeTiled)fn;
  return 0;
}

static int encode_view(CUtensorMap* tm, const LuSrcView& v, const void* ptr, int fmt = 0) {
  cuuint64_t dims[5] = {(cuuint64_t)v.dimC, (cuuint64_t)v.dimW, (cuuint64_t)v.dimP, (cuuint64_t)v.dimH, (cuuint64_t)v.dimN};
  cuuint64_t strides[4] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sp * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)v.pitch, 1, (cuuint32_t)v.rows, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, fmt ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (dims %d %d %d %d %d)", (int)r, v.dimC, v.dimW, v.dimP,
             v.dimH, v.dimN);
  return 0;
}

static int encode_weights(CUtensorMap* tm, const void* ptr, int npad, int ktot, int BN, int fmt = 0) {
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)npad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(tm, fmt ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LU_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  return 0;
}
#endif

static const uint16_t* view_ptr(lu_handle_s* h, int buf) {
  return reinterpret_cast<const uint16_t*>(h->ws + (buf < 0 ? h->off_patches : h->acts[buf].off));
}

// Fill the runtime part of the conv parameters (pointers) and launch on the selected engine.
//   frames      : number of tile frames in this launch (B for a ConvLSTM step, B*T otherwise)
//   mul/add     : per-view frame coordinate mapping; hstate_sel >= 0 swaps view 1 for the h-state buffer
static int launch_conv(lu_handle_s* h, ConvPlan& cv, int frames, const int* mul, const int* add, int hstate_sel,
                       const LuEpi& epi, void* stream) {
  LuConvParams p;
  memset(&p, 0, sizeof p);
  for (int i = 0; i < cv.n_views; ++i) {
    p.src[i] = cv.views[i];
    p.src[i].ptr = view_ptr(h, cv.view_buf[i]);
    p.src[i].frame_mul = mul[i]; p.src[i].frame_add = add[i];
  }
  if (hstate_sel >= 0) {
    p.src[1].ptr = reinterpret_cast<const uint16_t*>(h->ws + cv.off_hstate[hstate_sel]);
    p.src[1].dimN = h->cfg.batch;
  }
  p.astages = reinterpret_cast<const LuAStage*>(h->ws + cv.off_astages);
  p.taps = reinterpret_cast<const uint16_t*>(h->ws + cv.off_taps);
  p.wpacked = reinterpret_cast<const uint16_t*>(h->ws + cv.off_w);
  p.n_astages = (int)cv.astages.size(); p.ktot = cv.ktot;
  p.tiles_x = (cv.Wout + LU_TILE_W - 1) / LU_TILE_W; p.tiles_y = (cv.Hout + LU_TILE_H - 1) / LU_TILE_H;
  p.frames = frames; p.n_tiles_n = cv.n_tiles_n; p.BN = cv.BN;
  p.epi = epi;
  const int64_t m_tiles = (int64_t)p.tiles_x * p.tiles_y * frames;
  if (h->cfg.engine == LU_ENGINE_SIMT) {
    const int chunks = (epi.kind == LU_EPI_LSTM) ? epi.ch_tile / 16 : cv.BN / 16;
    LuMirrorItem it; it.p = p;
    pf(h, m_tiles * cv.n_tiles_n * 128 * chunks, stream, it);
    return 0;
  }
#ifdef LU_HOST_EMU
  LU_FAIL("the tcgen05 engine does not exist in the host test build");
#else
  bool* attr_set = h->attr_set;
  LuTcParams tp;
  memset(&tp, 0, sizeof tp);
  for (int i = 0; i < cv.n_views; ++i) tp.tmA[i] = cv.tmA[i];
  if (hstate_sel >= 0) tp.tmA[1] = cv.tmHstate[hstate_sel];
  tp.tmB = cv.tmB;
  tp.cp = p;
  tp.n_a_stages = cv.nA; tp.n_b_stages = cv.nB; tp.a_stage_bytes = cv.a_bytes; tp.b_stage_bytes = cv.b_bytes;
  tp.b_group = cv.b_group;
  // instruction descriptor: D = fp32 (bit 4), A / B format at bits [7,10) / [10,13): 1 = bf16, 0 = fp16; N >> 3, M >> 4
  const uint32_t ab_fmt = h->fmt ? 0u : ((1u << 7) | (1u << 10));
  tp.idesc = (1u << 4) | ab_fmt | ((uint32_t)(cv.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // thread-block clusters of 2 (weight multicast) for the ConvLSTM launches; LU_CLUSTER=1 disables
  static int cluster_env = -1;
  if (cluster_env < 0) { const char* ce = getenv("LU_CLUSTER"); cluster_env = ce ? atoi(ce) : 2; }
  // ... and for the conv / data-gradient launches whose weight stream is large (K >= 2048, N tile >= 128): without
  // the multicast every CTA pulls the whole K x N weight panel from L2 for every tile (measured on the level-1
  // data gradient: 70 % tensor-pipe activity at ~14 TB/s of L2->SM weight traffic).  LU_CLUSTER_WIDE=0 disables.
  static int wide_env = -1;
  if (wide_env < 0) { const char* ce = getenv("LU_CLUSTER_WIDE"); wide_env = ce ? atoi(ce) : 1; }
  // thresholds of "large weight stream" (experiment switches LU_WIDE_MIN_K / LU_WIDE_MIN_BN)
  static int wide_min_k = -1, wide_min_bn = -1;
  if (wide_min_k < 0) { const char* ce = getenv("LU_WIDE_MIN_K"); wide_min_k = ce ? atoi(ce) : 1024; }      // round 2: 2048 -> 1024, +1 % on the C2 step
  if (wide_min_bn < 0) { const char* ce = getenv("LU_WIDE_MIN_BN"); wide_min_bn = ce ? atoi(ce) : 128; }
  const bool wide = wide_env == 1 && epi.kind != LU_EPI_LSTM && cv.ktot >= wide_min_k && cv.BN >= wide_min_bn && m_tiles >= 2 * h->num_sms;
  const bool cl2 = cluster_env == 2 && (epi.kind == LU_EPI_LSTM || wide) && cv.ptab_ok && (h->num_sms % 2 == 0);
  // The cluster launches run ONE M = 256 MMA per CTA pair (tcgen05.mma.cta_group::2, kernel cluster mode 3) instead of two
  // M = 128 MMAs fed by a multicast weight stage: each CTA stages only half of every weight K block, so the weight stages
  // shrink to half, the shared memory they free goes to deeper activation prefetch, and every MMA reads a third less
  // operand data per SM.  Measured on B200 (round 2, profiles/README.md): C2 inference 358.7 -> 376.9 frames/s, ConvLSTM
  // launches 0.92 -> 0.975 of the sustained bf16 peak.  LU_PAIR=0 selects the multicast form (cluster mode 2) for A/B runs.
  static int pair_env = -1;
  if (pair_env < 0) { const char* ce = getenv("LU_PAIR"); pair_env = ce ? atoi(ce) : 1; }
  const bool pair = cl2 && pair_env == 1 && cv.BN >= 32;
  int smem_bytes = cv.smem;
  if (pair) {
    const int budget = 232448 - 1024 - 512 - 6144;               // as in the stage sizing of the plan
    tp.b_stage_bytes = cv.b_bytes / 2;
    int na = (budget - cv.nB * tp.b_stage_bytes) / cv.a_bytes;
    if (na > 8) na = 8;
    if (na > cv.nA) tp.n_a_stages = na;
    smem_bytes = tp.n_a_stages * cv.a_bytes + tp.n_b_stages * tp.b_stage_bytes + 1024 + 512 + 6144;
  }
  tp.num_mt = (int)m_tiles;
  tp.total_tiles = cl2 ? (int)(((m_tiles + 1) / 2) * cv.n_tiles_n) : (int)(m_tiles * cv.n_tiles_n);
  tp.tmBh = cv.tmBh;
  static int resident_env = -1;
  if (resident_env < 0) { const char* ce = getenv("LU_B_RESIDENT"); resident_env = ce ? atoi(ce) : 1; }
  tp.b_resident = (cv.b_resident && !cl2 && resident_env == 1) ? 1 : 0;
  static int cst_env = -1;
  if (cst_env < 0) { const char* ce = getenv("LU_CST_PER_TILE"); cst_env = ce ? atoi(ce) : 0; }
  tp.cst_per_tile = cst_env;
  // experiment switch LU_ACC_SPLIT (default off; measured no effect, see LuTcParams::acc_split): R partial accumulators for
  // narrow N tiles
  static int split_env = -1;
  if (split_env < 0) { const char* ce = getenv("LU_ACC_SPLIT"); split_env = ce ? atoi(ce) : 1; }
  tp.acc_split = 1;
  if (epi.kind != LU_EPI_LSTM && !cl2 && (split_env == 2 || split_env == 4) && cv.BN * split_env * 2 <= 512 && cv.BN <= 64)
    tp.acc_split = split_env;
  static int two_env = -1;
  if (two_env < 0) { const char* ce = getenv("LU_TWO_ISSUERS"); two_env = ce ? atoi(ce) : 0; }
  // EXPERIMENTAL (LU_TWO_ISSUERS=1, default off): a second issuing thread for the resident-weight (narrow) convolutions.
  // Measured on the C2 bench (1-4 activation stages per tile, 8 ring slots): the non-ConvLSTM part of the step 12.5 -> 11.8 ms.
  // The two issuers walk ONE in-order ring of activation stages, each skipping the other's tiles without looking at their
  // barriers.  That is sound only while the ring is DEEPER than one tile's stages: with n_a_stages <= n_astages an issuer can
  // reach a slot two laps after the producer last filled it, and the parity wait aliases (it passes before the data of that
  // lap has landed) -- the first version hung on hardware in the per-tap `direct` staging mode (27 stages per tile, 8 slots).
  // tests/test_ring_protocol.py models the protocol and pins the rule; not yet re-run on hardware with this guard.
  tp.two_issuers = (tp.b_resident && !cl2 && epi.kind != LU_EPI_LSTM && two_env == 1 &&
                    tp.n_a_stages > (int)cv.astages.size()) ? 1 : 0;
  tp.tables_in_params = cv.ptab_ok ? 1 : 0;
  if (cv.ptab_ok) {
    memcpy(tp.st_tab, cv.pstages.data(), cv.pstages.size() * sizeof(LuAStage));
    memcpy(tp.tap_tab, cv.ptaps.data(), cv.ptaps.size() * sizeof(uint16_t));
  }
  int grid = tp.total_tiles * (cl2 ? 2 : 1) < h->num_sms ? tp.total_tiles * (cl2 ? 2 : 1) : h->num_sms;
  const int ei = cl2 ? (epi.kind == LU_EPI_LSTM ? 6 : (epi.kind == LU_EPI_GRAD ? 7 : 8)) + (pair ? 3 : 0) : epi.kind * 2 + (cv.ptab_ok ? 1 : 0);
  typedef void (*KernelFn)(const LuTcParams);
  static const KernelFn kfn[12] = {lu_conv_tc_kernel<LU_EPI_CONV, false, 1>, lu_conv_tc_kernel<LU_EPI_CONV, true, 1>,
                                  lu_conv_tc_kernel<LU_EPI_LSTM, false, 1>, lu_conv_tc_kernel<LU_EPI_LSTM, true, 1>,
                                  lu_conv_tc_kernel<LU_EPI_GRAD, false, 1>, lu_conv_tc_kernel<LU_EPI_GRAD, true, 1>,
                                  lu_conv_tc_kernel<LU_EPI_LSTM, true, 2>, lu_conv_tc_kernel<LU_EPI_GRAD, true, 2>,
                                  lu_conv_tc_kernel<LU_EPI_CONV, true, 2>,
                                  lu_conv_tc_kernel<LU_EPI_LSTM, true, 3>, lu_conv_tc_kernel<LU_EPI_GRAD, true, 3>,
                                  lu_conv_tc_kernel<LU_EPI_CONV, true, 3>};
  if (!attr_set[ei]) {
    cudaError_t e = cudaFuncSetAttribute(kfn[ei], cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    LU_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[ei] = true;
  }
  h->launches++;
  time_begin(h, epi.kind == LU_EPI_LSTM ? LU_KC_LSTM_FWD : (epi.kind == LU_EPI_GRAD ? LU_KC_DGRAD : LU_KC_CONV_FWD), stream);
  if (cl2) {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(lutc::kThreads); lc.dynamicSmemBytes = smem_bytes;
    lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, kfn[ei], tp);
    LU_REQUIRE(e == cudaSuccess, "cluster launch (%s) failed: %s", cv.name.c_str(), cudaGetErrorString(e));
  } else {
    kfn[ei]<<<grid, lutc::kThreads, cv.smem, (cudaStream_t)stream>>>(tp);
  }
  time_end(h, stream);
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "conv launch (%s) failed: %s", cv.name.c_str(), cudaGetErrorString(e));
  return 0;
#endif
}

#include "lu_train_host.inl"

// ------------------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------------------
template <int PW>
static void prep_patches(lu_handle_s* h, const float* dev_x, int N, void* stream) {
  const lu_config& c = h->cfg;
  LuPrepPatchesT<PW> pp;
  pp.x = dev_x; pp.out = reinterpret_cast<uint16_t*>(h->ws + h->off_patches);
  pp.H = c.height; pp.W = c.width; pp.Hp = h->Hp; pp.Wp = h->Wp; pp.pad_y0 = h->pad_y0; pp.pad_x0 = h->pad_x0;
  pp.pw = h->pw; pp.x3 = h->planes == 2; pp.fmt = h->fmt;
  pf(h, (int64_t)N * h->Hp * h->Wp, stream, pp);
}

extern "C" {

const char* lu_last_error(void) { return g_err.c_str(); }
int lu_version(void) { return 1; }
int lu_is_cuda_build(void) {
#ifdef LU_HOST_EMU
  return 0;
#else
  return 1;
#endif
}

int lu_create(const lu_config* cfg, lu_handle* out) {
  LU_REQUIRE(cfg && out, "null argument");
  LU_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= LU_MAX_LEVELS, "n_levels must be in [1,%d]", LU_MAX_LEVELS);
  LU_REQUIRE(cfg->batch >= 1 && cfg->max_t >= 1 && cfg->height >= 1 && cfg->width >= 1, "bad shape");
  LU_REQUIRE(cfg->lrelu_alpha >= 0.f && cfg->lrelu_alpha <= 1.f, "lrelu_alpha must be in [0, 1] (Keras LeakyReLU default: 0.3), got %g", (double)cfg->lrelu_alpha);
  LU_REQUIRE(cfg->block_kind == LU_BLOCK_NET || cfg->block_kind == LU_BLOCK_DOWN || cfg->block_kind == LU_BLOCK_UP,
             "unknown block_kind %d", cfg->block_kind);
  if (cfg->block_kind != LU_BLOCK_NET) {
    LU_REQUIRE(cfg->n_levels == 1, "a stand-alone block is described by level 0 of the lists (n_levels = 1)");
    LU_REQUIRE(cfg->block_stride == 1 || cfg->block_stride == 2, "stride / up_factor must be 1 or 2, got %d", cfg->block_stride);
    LU_REQUIRE(!cfg->train, "stand-alone blocks are forward-only handles");
    LU_REQUIRE(cfg->block_kind == LU_BLOCK_DOWN || cfg->max_t == 1, "UpBlock2D takes 4-D inputs: max_t must be 1");
  }
  for (int l = 0; l < cfg->n_levels; ++l) {
    if (cfg->block_kind != LU_BLOCK_UP) {
      LU_REQUIRE(cfg->n_lstm[l] >= (l == 0 ? 1 : 0) && cfg->n_lstm[l] <= LU_MAX_PER_LEVEL, "bad ConvLSTM count at level %d", l);
      LU_REQUIRE(cfg->n_down[l] >= 1 && cfg->n_down[l] <= LU_MAX_PER_LEVEL, "bad conv count at level %d", l);
    }
    if (cfg->block_kind != LU_BLOCK_DOWN)
      LU_REQUIRE(cfg->n_up[l] >= 1 && cfg->n_up[l] <= LU_MAX_PER_LEVEL, "bad up-conv count at level %d", l);
  }
#ifdef LU_HOST_EMU
  LU_REQUIRE(cfg->engine == LU_ENGINE_SIMT, "host test build only has the scalar mirror engine");
#endif
  lu_handle_s* h = new lu_handle_s();
  h->cfg = *cfg;
  h->L = cfg->n_levels;
  LU_REQUIRE(cfg->precision == LU_PREC_BF16 || cfg->precision == LU_PREC_BF16X3 || cfg->precision == LU_PREC_FP16, "unknown precision %d", cfg->precision);
  LU_REQUIRE(!(cfg->precision == LU_PREC_FP16 && cfg->train),
             "fp16 operands are an inference mode (gradients of ~1e-7 underflow fp16); train with bf16 or bf16x3");
  h->planes = cfg->precision == LU_PREC_BF16X3 ? 2 : 1;
  h->fmt = cfg->precision == LU_PREC_FP16 ? 1 : 0;
  build_params(h);
  if (build_plan(h)) { delete h; return 1; }
  layout_workspace(h);
  *out = h;
  return 0;
}

int lu_destroy(lu_handle h) {
  if (h) {
    for (size_t i = 0; i < h->graphs.size(); ++i) {
#ifndef LU_HOST_EMU
      cudaGraphExecDestroy(h->graphs[i].exec);
#endif
    }
    h->graphs.clear();
#ifndef LU_HOST_EMU
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
#endif
    train_destroy(h); delete h;
  }
  return 0;
}

int lu_workspace_bytes(lu_handle h, size_t* bytes) {
  LU_REQUIRE(h && bytes, "null argument");
  *bytes = h->ws_bytes;
  return 0;
}

int lu_bind_workspace(lu_handle h, void* dev_ws, size_t bytes, void* stream) {
  LU_REQUIRE(h && dev_ws, "null argument");
  LU_REQUIRE(bytes >= h->ws_bytes, "workspace too small: %zu < %zu", bytes, h->ws_bytes);
  LU_REQUIRE(((uintptr_t)dev_ws & 1023) == 0, "workspace must be 1024-byte aligned");
  h->ws = (uint8_t*)dev_ws;
  for (size_t i = 0; i < h->graphs.size(); ++i) {      // instantiated graphs point into the previous workspace
#ifndef LU_HOST_EMU
    cudaGraphExecDestroy(h->graphs[i].exec);
#endif
  }
  h->graphs.clear();
#ifndef LU_HOST_EMU
  {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    LU_REQUIRE(e == cudaSuccess, "no CUDA device: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    LU_REQUIRE(e == cudaSuccess, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    LU_REQUIRE(prop.major == 10, "this library is built for sm_100a (Blackwell B200); found sm_%d%d", prop.major, prop.minor);
    h->num_sms = prop.multiProcessorCount;
  }
#endif
  LU_MEMSET(h->ws, 0, h->ws_bytes, stream);
  for (auto& cv : h->convs) {
    cv.wg_cached_T[0] = cv.wg_cached_T[1] = -1;       // device-resident task lists were just wiped
    LU_H2D(h->ws + cv.off_astages, cv.astages.data(), cv.astages.size() * sizeof(LuAStage), stream);
    LU_H2D(h->ws + cv.off_taps, cv.taps.data(), cv.taps.size() * sizeof(uint16_t), stream);
    LU_H2D(h->ws + cv.off_packs, cv.packs.data(), cv.packs.size() * sizeof(LuPackDesc), stream);
  }
  train_upload(h, stream);
#ifndef LU_HOST_EMU
  if (h->cfg.engine == LU_ENGINE_TCGEN05) {
    if (get_encode()) return 1;
    for (auto& cv : h->convs) {
      for (int i = 0; i < cv.n_views; ++i)
        if (encode_view(&cv.tmA[i], cv.views[i], view_ptr(h, cv.view_buf[i]), h->fmt)) return 1;
      if (cv.kind == LU_EPI_LSTM)
        for (int s = 0; s < 2; ++s) {
          LuSrcView v = cv.views[1];
          v.dimN = h->cfg.batch;
          if (encode_view(&cv.tmHstate[s], v, h->ws + cv.off_hstate[s], h->fmt)) return 1;
        }
      if (encode_weights(&cv.tmB, h->ws + cv.off_w, cv.npad, cv.ktot, cv.BN, h->fmt)) return 1;
      if (encode_weights(&cv.tmBh, h->ws + cv.off_w, cv.npad, cv.ktot, cv.BN / 2, h->fmt)) return 1;
    }
    if (h->cfg.train) {
      h->acts_tm.resize(h->acts.size());
      for (size_t i = 0; i < h->acts.size(); ++i) {
        const ActBuf& a = h->acts[i];
        LuSrcView v; memset(&v, 0, sizeof v);
        const int ctot = a.cpad * a.planes;
        v.dimC = ctot; v.dimW = a.W; v.dimP = 1; v.dimH = a.H; v.dimN = a.frames;
        v.sw = ctot; v.sh = (int64_t)a.W * ctot; v.sp = v.sh; v.sn = (int64_t)a.H * a.W * ctot;
        v.rows = LU_TILE_H; v.pitch = LU_TILE_W;
        if (encode_view(&h->acts_tm[i], v, h->ws + a.off)) return 1;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "bind_workspace: %s", cudaGetErrorString(e));
#endif
  h->bound = true;
  h->hcur = 0;
  return 0;
}

int lu_param_count(lu_handle h, int32_t* n_tensors, int64_t* n_elements, int64_t* n_trainable_elements) {
  LU_REQUIRE(h, "null handle");
  if (n_tensors) *n_tensors = (int32_t)h->params.size();
  if (n_elements) *n_elements = h->n_elems;
  if (n_trainable_elements) *n_trainable_elements = h->n_train;
  return 0;
}

int lu_param_info(lu_handle h, int32_t idx, char* name, int32_t name_cap, int64_t* shape4, int32_t* rank, int64_t* offset,
                  int32_t* trainable) {
  LU_REQUIRE(h && idx >= 0 && idx < (int)h->params.size(), "bad parameter index %d", idx);
  const ParamT& p = h->params[idx];
  if (name && name_cap > 0) { strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = p.shape[i];
  if (rank) *rank = p.rank;
  if (offset) *offset = p.offset;
  if (trainable) *trainable = p.trainable ? 1 : 0;
  return 0;
}

int lu_bind_params(lu_handle h, float* dev_params) {
  LU_REQUIRE(h && dev_params, "null argument");
  h->dparams = dev_params;
  h->packed = false;
  return 0;
}

// inference BatchNorm: moving statistics folded into the per-channel scale / shift of the conv epilogue
static void fold_bn(lu_handle_s* h, void* stream) {
  for (auto& cv : h->convs) {
    if (cv.kind == LU_EPI_GRAD || !cv.has_bn) continue;
    LuBnFold f;
    f.gamma = h->dparams + h->params[cv.gamma].offset; f.beta = h->dparams + h->params[cv.beta].offset;
    f.mov_mean = h->dparams + h->params[cv.mov_mean].offset; f.mov_var = h->dparams + h->params[cv.mov_var].offset;
    f.scale = reinterpret_cast<float*>(h->ws + cv.off_scale); f.shift = reinterpret_cast<float*>(h->ws + cv.off_shift);
    f.c_real = cv.cout; f.eps = 1e-3f;
    pf(h, cv.npad, stream, f);
  }
  h->bn_fold_stale = false;
}

int lu_params_changed(lu_handle h, void* stream) {
  LU_REQUIRE(h && h->bound && h->dparams, "bind workspace and parameters first");
  for (auto& cv : h->convs) {
    LuPackWeights pw;
    pw.params = h->dparams; pw.descs = reinterpret_cast<const LuPackDesc*>(h->ws + cv.off_packs);
    pw.out = reinterpret_cast<uint16_t*>(h->ws + cv.off_w); pw.cm = cv.cm; pw.ktot = cv.ktot; pw.fmt = h->fmt;
    pf(h, (int64_t)cv.npad * cv.ktot, stream, pw);
    if (cv.kind == LU_EPI_GRAD) continue;
    LuPackVec pb;
    pb.src = h->dparams + h->params[cv.bias_param].offset; pb.out = reinterpret_cast<float*>(h->ws + cv.off_bias); pb.cm = cv.cm;
    pf(h, cv.npad, stream, pb);
  }
  fold_bn(h, stream);
  h->packed = true;
  return 0;
}

// ---- forward -----------------------------------------------------------------------------------------------------
static int run_conv_layer(lu_handle_s* h, ConvPlan& cv, int T, int training, void* stream) {
  const int N = h->cfg.batch * T;
  int mul[LU_MAX_SRC] = {1, 1, 1, 1}, add[LU_MAX_SRC] = {0, 0, 0, 0};
  LuEpi e;
  memset(&e, 0, sizeof e);
  e.kind = LU_EPI_CONV; e.H = cv.Hout; e.W = cv.Wout; e.fmt = h->fmt;
  e.oy_mul = 1; e.ox_mul = 1; e.OH = cv.Hout; e.OW = cv.Wout;
  e.bias = reinterpret_cast<const float*>(h->ws + cv.off_bias);
  e.out_frame_mul = 1; e.out_frame_add = 0; e.alpha = h->cfg.lrelu_alpha;
  e.raw_cpad = cv.raw_cpad;
  float* raw = reinterpret_cast<float*>(h->ws + cv.off_raw);
  const bool bn_batch = cv.has_bn && training;
  if (!cv.has_bn) {
    e.out_raw = raw;                                            // logits: raw fp32 only
  } else if (bn_batch) {
    e.out_raw = raw;                                            // batch statistics need the whole output first
    // ... and are accumulated by the conv epilogue itself (per-channel sum / sum of squares of output - bias).  The epilogue
    // combines its warps' partial sums with fp32 shared-memory atomics, so the statistics -- and with them the training
    // forward -- are reproducible to the last bits only, not bitwise (like cuDNN's atomics-based reductions);
    // LU_BN_STATS=separate selects the fixed-order separate pass (1.5 ms per C3 step) for bitwise-repeatable runs.
    static int bn_separate = -1;
    if (bn_separate < 0) { const char* ce = getenv("LU_BN_STATS"); bn_separate = (ce && strcmp(ce, "separate") == 0) ? 1 : 0; }
    if (cv.npad <= 512 && !bn_separate) e.bn_sums = reinterpret_cast<double*>(h->ws + cv.off_sums);
  } else {
    const ActBuf& ob = h->acts[cv.out_buf];
    e.out_act = reinterpret_cast<uint16_t*>(h->ws + ob.off); e.out_cpad = ob.cpad; e.out_planes = ob.planes;
    e.scale = reinterpret_cast<const float*>(h->ws + cv.off_scale);
    e.shift = reinterpret_cast<const float*>(h->ws + cv.off_shift);
  }
  if (launch_conv(h, cv, N, mul, add, -1, e, stream)) return 1;
  if (bn_batch) {
    const ActBuf& ob = h->acts[cv.out_buf];
    const int64_t npix = (int64_t)N * cv.Hout * cv.Wout;
    LuBnStats st;
    st.raw = raw; st.sums = reinterpret_cast<double*>(h->ws + cv.off_sums);
    st.shift_src = e.bn_sums ? e.bias : raw;                   // the shift the sums are taken about: the bias, or (separate pass) the first pixel
    st.cpad = cv.raw_cpad; st.c_real = cv.cout;
    if (!e.bn_sums) rows(h, npix, cv.raw_cpad / 8, stream, st);        // layers wider than the epilogue's 512-column accumulators
    LuBnFinalize fin;
    fin.sums = st.sums; fin.shift_src = st.shift_src;
    fin.gamma = h->dparams + h->params[cv.gamma].offset; fin.beta = h->dparams + h->params[cv.beta].offset;
    fin.mov_mean = h->dparams + h->params[cv.mov_mean].offset; fin.mov_var = h->dparams + h->params[cv.mov_var].offset;
    fin.scale = reinterpret_cast<float*>(h->ws + cv.off_bscale); fin.shift = reinterpret_cast<float*>(h->ws + cv.off_bshift);
    fin.save_mean = reinterpret_cast<float*>(h->ws + cv.off_save_mean);
    fin.save_invstd = reinterpret_cast<float*>(h->ws + cv.off_save_invstd);
    fin.npix = npix; fin.cpad = cv.raw_cpad; fin.c_real = cv.cout; fin.eps = 1e-3f; fin.momentum = 0.99f;
    if (h->bn_sync_fn) {
      // statistics over the GLOBAL batch (the reference's single-device semantics): local moments ->
      // in-place sum over the ranks by the caller (one small fp64 vector per BN layer) -> combined mean / variance
      LuBnMoments mo;
      mo.sums = st.sums; mo.shift_src = st.shift_src; mo.mom = reinterpret_cast<double*>(h->ws + cv.off_mom);
      mo.npix = npix; mo.cpad = cv.raw_cpad; mo.c_real = cv.cout;
      pf(h, cv.raw_cpad, stream, mo);
      h->bn_sync_fn(mo.mom, (int64_t)3 * cv.raw_cpad, h->bn_sync_user);
      LuBnFinalizeSync fs;
      fs.mom = mo.mom; fs.gamma = fin.gamma; fs.beta = fin.beta; fs.mov_mean = fin.mov_mean; fs.mov_var = fin.mov_var;
      fs.scale = fin.scale; fs.shift = fin.shift; fs.save_mean = fin.save_mean; fs.save_invstd = fin.save_invstd;
      fs.npix = npix; fs.cpad = cv.raw_cpad; fs.c_real = cv.cout; fs.world = h->bn_sync_world; fs.eps = fin.eps; fs.momentum = fin.momentum;
      pf(h, cv.raw_cpad, stream, fs);
    } else {
      pf(h, cv.raw_cpad, stream, fin);
    }
    LuBnApply ap;
    ap.raw = raw; ap.scale = fin.scale; ap.shift = fin.shift;
    ap.out = reinterpret_cast<uint16_t*>(h->ws + ob.off); ap.raw_cpad = cv.raw_cpad; ap.out_cpad = ob.cpad; ap.planes = ob.planes;
    ap.alpha = h->cfg.lrelu_alpha; ap.fmt = h->fmt;
    rows(h, npix, ob.cpad / 8, stream, ap);
  }
  return 0;
}

static int run_lstm_layer(lu_handle_s* h, ConvPlan& cv, int T, int training, void* stream) {
  const ActBuf& hs = h->acts[cv.hseq_buf];
  if (training && h->cfg.train)      // BPTT needs c before the first step of this call (c is updated in place)
    LU_D2D(h->ws + cv.off_c_init, h->ws + cv.off_cstate, (size_t)h->cfg.batch * cv.Hout * cv.Wout * cv.fpad * 4, stream);
  for (int t = 0; t < T; ++t) {
    int mul[LU_MAX_SRC] = {T, T, 1, 1}, add[LU_MAX_SRC] = {t, t - 1, 0, 0};
    int sel = -1;
    if (t == 0) { sel = h->hcur; mul[1] = 1; add[1] = 0; }
    LuEpi e;
    memset(&e, 0, sizeof e);
    e.kind = LU_EPI_LSTM; e.H = cv.Hout; e.W = cv.Wout; e.fmt = h->fmt;
    e.oy_mul = 1; e.ox_mul = 1; e.OH = cv.Hout; e.OW = cv.Wout;
    e.bias = reinterpret_cast<const float*>(h->ws + cv.off_bias);
    e.out_frame_mul = T; e.out_frame_add = t;
    e.out_act = reinterpret_cast<uint16_t*>(h->ws + hs.off); e.out_cpad = hs.cpad; e.out_planes = hs.planes;
    e.c_state = reinterpret_cast<float*>(h->ws + cv.off_cstate);
    e.h_state_out = (t == T - 1) ? reinterpret_cast<uint16_t*>(h->ws + cv.off_hstate[h->hcur ^ 1]) : nullptr;
    e.f_pad = cv.fpad; e.ch_tile = 64; e.gate_kind = h->cfg.gate;
    if (training && h->cfg.train) {
      e.save_gates = reinterpret_cast<uint16_t*>(h->ws + cv.off_save_gates);
      e.save_c = reinterpret_cast<float*>(h->ws + cv.off_save_c);
    }
    if (launch_conv(h, cv, h->cfg.batch, mul, add, sel, e, stream)) return 1;
  }
  return 0;
}

static int forward_body(lu_handle h, const float* dev_x, int32_t T, int32_t training, float* dev_logits,
                        float* dev_softmax, void* stream, const float* dev_skip = nullptr);

int lu_set_graph_mode(lu_handle h, int32_t enable, int32_t* effective) {
  LU_REQUIRE(h, "null handle");
  h->graph_mode = (enable && h->graph_ok) ? 1 : 0;
#ifdef LU_HOST_EMU
  h->graph_mode = 0;
#endif
  if (effective) *effective = h->graph_mode;
  return 0;
}

int lu_forward(lu_handle h, const float* dev_x, int32_t T, int32_t training, float* dev_logits, float* dev_softmax,
               void* stream) {
  LU_REQUIRE(h && h->bound && h->dparams, "bind workspace and parameters first");
  LU_REQUIRE(h->cfg.block_kind == LU_BLOCK_NET, "this handle is a stand-alone block: call lu_block_forward");
  LU_REQUIRE(dev_x && dev_logits && dev_softmax, "null tensor");
  LU_REQUIRE(T >= 1 && T <= h->cfg.max_t, "T=%d outside [1,%d]", T, h->cfg.max_t);
  if (!h->packed && lu_params_changed(h, stream)) return 1;
  // model(x, training=True) updates the BatchNorm moving statistics (Keras does so without any optimizer step, as in the
  // reference's unit_test loops): the next inference forward must normalise with the new ones
  if (!training && h->bn_fold_stale) fold_bn(h, stream);
#ifndef LU_HOST_EMU
  if (h->graph_mode && !training && !h->time_on) {
    // replay path: stage the input, launch the instantiated graph of this (T, state parity), copy the outputs out
    const lu_config& c = h->cfg;
    float* gx = reinterpret_cast<float*>(h->ws + h->off_gx);
    float* gl = reinterpret_cast<float*>(h->ws + h->off_glogits);
    float* gs = reinterpret_cast<float*>(h->ws + h->off_gsoftmax);
    lu_handle_s::FwdGraph* g = nullptr;
    for (auto& e : h->graphs) if (e.T == T && e.hcur == h->hcur) g = &e;
    if (!g) {
      if (!h->cap_stream) {
        cudaError_t e = cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
        LU_REQUIRE(e == cudaSuccess, "cudaStreamCreate: %s", cudaGetErrorString(e));
      }
      const int hcur0 = h->hcur; const int64_t l0 = h->launches;
      cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
      LU_REQUIRE(e == cudaSuccess, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
      const int rc = forward_body(h, gx, T, 0, gl, gs, h->cap_stream);
      cudaGraph_t graph = nullptr;
      e = cudaStreamEndCapture(h->cap_stream, &graph);
      h->hcur = hcur0;
      lu_handle_s::FwdGraph ng; ng.T = T; ng.hcur = hcur0; ng.launches = h->launches - l0; ng.exec = nullptr;
      h->launches = l0;
      if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
      LU_REQUIRE(e == cudaSuccess && graph, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
      e = cudaGraphInstantiate(&ng.exec, graph, 0);
      cudaGraphDestroy(graph);
      LU_REQUIRE(e == cudaSuccess, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
      h->graphs.push_back(ng);
      g = &h->graphs.back();
    }
    const size_t frac_x = (size_t)c.batch * T * c.in_channels * c.height * c.width * 4;
    const size_t frac_o = (size_t)c.batch * T * h->convs[h->logits_conv].cout * c.height * c.width * 4;
    cudaMemcpyAsync(gx, dev_x, frac_x, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
    LU_REQUIRE(e == cudaSuccess, "cudaGraphLaunch: %s", cudaGetErrorString(e));
    cudaMemcpyAsync(dev_logits, gl, frac_o, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    cudaMemcpyAsync(dev_softmax, gs, frac_o, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    h->launches += g->launches;
    h->hcur ^= 1;
    h->last_T = T; h->last_training = 0;
    e = cudaGetLastError();
    LU_REQUIRE(e == cudaSuccess, "forward (graph): %s", cudaGetErrorString(e));
    return 0;
  }
#endif
  return forward_body(h, dev_x, T, training, dev_logits, dev_softmax, stream);
}

static int forward_body(lu_handle h, const float* dev_x, int32_t T, int32_t training, float* dev_logits,
                        float* dev_softmax, void* stream, const float* dev_skip) {
  const lu_config& c = h->cfg;
  const int N = c.batch * T;
  if (h->skip_in_buf >= 0) {                      // UpBlock2D alone: its second input
    const ActBuf& sb = h->acts[h->skip_in_buf];
    LuPrepImage pi;
    pi.x = dev_skip; pi.out = reinterpret_cast<uint16_t*>(h->ws + sb.off);
    pi.C = c.skip_channels; pi.H = sb.H; pi.W = sb.W; pi.Hp = sb.H; pi.Wp = sb.W; pi.pad_y0 = 0; pi.pad_x0 = 0;
    pi.cpad = sb.cpad; pi.planes = sb.planes; pi.fmt = h->fmt; pi.channels_first = c.channels_first;
    pf(h, (int64_t)N * sb.H * sb.W * (sb.cpad / 8), stream, pi);
  }
  if (h->img_buf >= 0) {
    const ActBuf& ib = h->acts[h->img_buf];
    LuPrepImage pi;
    pi.x = dev_x; pi.out = reinterpret_cast<uint16_t*>(h->ws + ib.off);
    pi.C = c.in_channels; pi.H = c.height; pi.W = c.width; pi.Hp = h->Hp; pi.Wp = h->Wp; pi.pad_y0 = h->pad_y0; pi.pad_x0 = h->pad_x0;
    pi.cpad = ib.cpad; pi.planes = ib.planes; pi.fmt = h->fmt; pi.channels_first = c.channels_first;
    pf(h, (int64_t)N * h->Hp * h->Wp * (ib.cpad / 8), stream, pi);
  } else {
    if (h->pw == 5) prep_patches<5>(h, dev_x, N, stream);
    else if (h->pw == 3) prep_patches<3>(h, dev_x, N, stream);
    else prep_patches<0>(h, dev_x, N, stream);
  }
  for (int l = 0; l < h->L; ++l) {
    for (int ci : h->lstm_of_level[l])
      if (run_lstm_layer(h, h->convs[ci], T, training, stream)) return 1;
    for (int ci : h->conv_of_level[l])
      if (run_conv_layer(h, h->convs[ci], T, training, stream)) return 1;
  }
  for (int u = 0; u < h->L; ++u) {
    if (h->ups[u].src_buf >= 0) {
      const ActBuf& s = h->acts[h->ups[u].src_buf];
      const ActBuf& d = h->acts[h->ups[u].dst_buf];
      LuUpsample2x up;
      up.in = reinterpret_cast<const uint16_t*>(h->ws + s.off); up.out = reinterpret_cast<uint16_t*>(h->ws + d.off);
      up.h = s.H; up.w = s.W; up.cpad = s.cpad; up.planes = s.planes; up.fmt = h->fmt;
      pf(h, (int64_t)N * s.H * s.W * (s.cpad / 8), stream, up);     // one item per INPUT pixel and 8 channels
    }
    for (int ci : h->conv_of_up[u])
      if (run_conv_layer(h, h->convs[ci], T, training, stream)) return 1;
  }
  if (c.block_kind != LU_BLOCK_NET) {
    // what the block returns (Networks.py:75,153): the last activation, or the last convolution's raw output
    // (return_logits), as fp32 in the caller's layout
    ConvPlan& oc = h->convs[h->block_out_conv];
    LuBlockOut bo;
    memset(&bo, 0, sizeof bo);
    bo.out = dev_logits; bo.C = oc.cout; bo.H = oc.Hout; bo.W = oc.Wout; bo.channels_first = c.channels_first; bo.fmt = h->fmt;
    if (oc.out_buf >= 0) {
      const ActBuf& ob = h->acts[oc.out_buf];
      bo.act = reinterpret_cast<const uint16_t*>(h->ws + ob.off); bo.cpad = ob.cpad; bo.planes = ob.planes;
    } else {
      bo.raw = reinterpret_cast<const float*>(h->ws + oc.off_raw); bo.cpad = oc.raw_cpad;
    }
    pf(h, (int64_t)N * oc.Hout * oc.Wout * oc.cout, stream, bo);
  } else {
    ConvPlan& lc = h->convs[h->logits_conv];
    LuSoftmaxCrop sm;
    sm.raw = reinterpret_cast<const float*>(h->ws + lc.off_raw); sm.logits = dev_logits; sm.softmax = dev_softmax;
    sm.B = c.batch; sm.T = T; sm.H = c.height; sm.W = c.width; sm.Hp = h->Hp; sm.Wp = h->Wp;
    sm.py0 = h->pad_y0; sm.px0 = h->pad_x0; sm.raw_cpad = lc.raw_cpad; sm.D = lc.cout; sm.channels_first = c.channels_first;
    const int64_t items = c.channels_first ? (int64_t)N * c.height * c.width : (int64_t)T * c.height * c.width * lc.cout;
    pf(h, items, stream, sm);
  }
  h->hcur ^= 1;                  // the last step of every ConvLSTM wrote h_T into the other buffer
  h->last_T = T; h->last_training = training;
  if (training) h->bn_fold_stale = true;
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "forward: %s", cudaGetErrorString(e));
#endif
  return 0;
}

int lu_block_forward(lu_handle h, const float* dev_x, const float* dev_skip, int32_t T, int32_t training, float* dev_out,
                     void* stream) {
  LU_REQUIRE(h && h->bound && h->dparams, "bind workspace and parameters first");
  LU_REQUIRE(h->cfg.block_kind != LU_BLOCK_NET, "this handle is the whole network: call lu_forward");
  LU_REQUIRE(dev_x && dev_out, "null tensor");
  LU_REQUIRE((h->cfg.block_kind == LU_BLOCK_UP) == (dev_skip != nullptr), "the skip input belongs to UpBlock2D (and only to it)");
  LU_REQUIRE(T >= 1 && T <= h->cfg.max_t, "T=%d outside [1,%d]", T, h->cfg.max_t);
  if (!h->packed && lu_params_changed(h, stream)) return 1;
  if (!training && h->bn_fold_stale) fold_bn(h, stream);
  return forward_body(h, dev_x, T, training, dev_out, nullptr, stream, dev_skip);
}

int lu_block_out_shape(lu_handle h, int64_t* shape4) {
  LU_REQUIRE(h && shape4, "null argument");
  LU_REQUIRE(h->cfg.block_kind != LU_BLOCK_NET, "not a stand-alone block");
  const ConvPlan& oc = h->convs[h->block_out_conv];
  shape4[0] = h->cfg.batch; shape4[1] = oc.cout; shape4[2] = oc.Hout; shape4[3] = oc.Wout;
  return 0;
}

// ---- recurrent state API -----------------------------------------------------------------------------------------
static ConvPlan* find_lstm(lu_handle_s* h, int level, int layer) {
  if (level < 0 || level >= h->L) return nullptr;
  if (layer < 0 || layer >= (int)h->lstm_of_level[level].size()) return nullptr;
  return &h->convs[h->lstm_of_level[level][layer]];
}

static void reset_level(lu_handle_s* h, int l, const float* dev_mask, void* stream) {
  for (int ci : h->lstm_of_level[l]) {
    ConvPlan& cv = h->convs[ci];
    const int64_t px = (int64_t)cv.Hout * cv.Wout;
    LuStateMask mh;
    mh.h = reinterpret_cast<uint16_t*>(h->ws + cv.off_hstate[h->hcur]); mh.c = nullptr; mh.mask = dev_mask;
    mh.per_sample_h = px * cv.fpad * h->planes; mh.per_sample_c = 0; mh.fmt = h->fmt;
    pf(h, (int64_t)h->cfg.batch * mh.per_sample_h, stream, mh);
    LuStateMaskC mc;
    mc.c = reinterpret_cast<float*>(h->ws + cv.off_cstate); mc.mask = dev_mask; mc.per_sample = px * cv.fpad;
    pf(h, (int64_t)h->cfg.batch * mc.per_sample, stream, mc);
  }
}

int lu_reset_states(lu_handle h, const float* dev_mask, void* stream) {
  LU_REQUIRE(h && h->bound && dev_mask, "null argument / unbound handle");
  for (int l = 0; l < h->L; ++l) reset_level(h, l, dev_mask, stream);
  return 0;
}

int lu_reset_level_states(lu_handle h, int32_t level, const float* dev_mask, void* stream) {
  LU_REQUIRE(h && h->bound && dev_mask, "null argument / unbound handle");
  LU_REQUIRE(level >= 0 && level < h->L, "no level %d", level);
  reset_level(h, level, dev_mask, stream);
  return 0;
}

int lu_state_shape(lu_handle h, int32_t level, int32_t layer, int64_t* shape4) {
  LU_REQUIRE(h && shape4, "null argument");
  ConvPlan* cv = find_lstm(h, level, layer);
  LU_REQUIRE(cv, "no ConvLSTM at level %d layer %d", level, layer);
  if (h->cfg.channels_first) { shape4[0] = h->cfg.batch; shape4[1] = cv->F; shape4[2] = cv->Hout; shape4[3] = cv->Wout; }
  else { shape4[0] = h->cfg.batch; shape4[1] = cv->Hout; shape4[2] = cv->Wout; shape4[3] = cv->F; }
  return 0;
}

int lu_get_state(lu_handle h, int32_t level, int32_t layer, int32_t which, float* dev_out, void* stream) {
  LU_REQUIRE(h && h->bound && dev_out, "null argument / unbound handle");
  ConvPlan* cv = find_lstm(h, level, layer);
  LU_REQUIRE(cv, "no ConvLSTM at level %d layer %d", level, layer);
  LuStateGet g;
  g.h = reinterpret_cast<const uint16_t*>(h->ws + cv->off_hstate[h->hcur]); g.c = reinterpret_cast<const float*>(h->ws + cv->off_cstate);
  g.out = dev_out; g.which = which; g.H = cv->Hout; g.W = cv->Wout; g.F = cv->F; g.fpad = cv->fpad; g.planes = h->planes;
  g.channels_first = h->cfg.channels_first; g.fmt = h->fmt;
  pf(h, (int64_t)h->cfg.batch * cv->F * cv->Hout * cv->Wout, stream, g);
  return 0;
}

int lu_set_state(lu_handle h, int32_t level, int32_t layer, int32_t which, const float* dev_in, void* stream) {
  LU_REQUIRE(h && h->bound, "unbound handle");
  ConvPlan* cv = find_lstm(h, level, layer);
  LU_REQUIRE(cv, "no ConvLSTM at level %d layer %d", level, layer);
  LuStateSet s;
  s.h = reinterpret_cast<uint16_t*>(h->ws + cv->off_hstate[h->hcur]); s.c = reinterpret_cast<float*>(h->ws + cv->off_cstate);
  s.in = dev_in; s.which = which; s.H = cv->Hout; s.W = cv->Wout; s.F = cv->F; s.fpad = cv->fpad; s.planes = h->planes;
  s.channels_first = h->cfg.channels_first; s.fmt = h->fmt;
  pf(h, (int64_t)h->cfg.batch * cv->F * cv->Hout * cv->Wout, stream, s);
  return 0;
}

int lu_launch_count(lu_handle h, int64_t* launches, int32_t reset) {
  LU_REQUIRE(h, "null handle");
  if (launches) *launches = h->launches;
  if (reset) h->launches = 0;
  return 0;
}

int lu_forward_flops(lu_handle h, int32_t T, double* flops) {
  LU_REQUIRE(h && flops, "null argument");
  double macs = 0;
  for (auto& cv : h->convs) if (cv.kind != LU_EPI_GRAD) macs += cv.macs_per_frame;
  *flops = 2.0 * macs * T;
  return 0;
}

int lu_lstm_flops(lu_handle h, int32_t T, double* flops) {
  LU_REQUIRE(h && flops, "null argument");
  double macs = 0;
  for (auto& cv : h->convs) if (cv.kind == LU_EPI_LSTM) macs += cv.macs_per_frame;
  *flops = 2.0 * macs * T;
  return 0;
}

int lu_kernel_times(lu_handle h, int32_t enable, float* ms4, int32_t* launches4) {
  LU_REQUIRE(h, "null handle");
  float total[LU_KC_COUNT] = {0.f, 0.f, 0.f, 0.f}; int n[LU_KC_COUNT] = {0, 0, 0, 0};
#ifndef LU_HOST_EMU
  if (h->ev_used) {
    cudaError_t e = cudaEventSynchronize(h->events[h->ev_used - 1]);
    LU_REQUIRE(e == cudaSuccess, "event sync: %s", cudaGetErrorString(e));
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->events[i], h->events[i + 1]);
      const int c = h->ev_class[i / 2];
      total[c] += ms; ++n[c];
    }
  }
#endif
  h->ev_used = 0;
  h->time_on = enable != 0;
  for (int c = 0; c < LU_KC_COUNT; ++c) { if (ms4) ms4[c] = total[c]; if (launches4) launches4[c] = n[c]; }
  return 0;
}

int lu_lstm_kernel_time(lu_handle h, int32_t enable, float* ms_total, int32_t* launches) {
  float ms[LU_KC_COUNT]; int32_t n[LU_KC_COUNT];
  if (lu_kernel_times(h, enable, ms, n)) return 1;
  if (ms_total) *ms_total = ms[LU_KC_LSTM_FWD];
  if (launches) *launches = n[LU_KC_LSTM_FWD];
  return 0;
}

// algorithmic FLOPs (2 * MAC) per kernel class of one training step (forward over T frames per sample x batch, and its
// backward): ConvLSTM forward, other forward convolutions, data gradients (no gradient flows to the image; the
// recurrent term exists for t > 0 only), weight gradients (every forward MAC once)
int lu_class_flops(lu_handle h, int32_t T, double* flops4) {
  LU_REQUIRE(h && flops4, "null argument");
  double lstm = 0, conv = 0, dgrad = 0;
  for (auto& cv : h->convs) {
    if (cv.kind == LU_EPI_GRAD) continue;
    (cv.kind == LU_EPI_LSTM ? lstm : conv) += cv.macs_per_frame;
    for (int i = 0; i < cv.n_in; ++i) {
      if (cv.in[i].buf < 0) continue;
      const double m = (double)cv.k * cv.k * cv.in[i].creal * cv.cout * cv.Hout * cv.Wout;
      dgrad += (cv.kind == LU_EPI_LSTM && i == 1) ? m * (T - 1) / (double)T : m;
    }
  }
  const double f = 2.0 * T * h->cfg.batch;
  flops4[LU_KC_LSTM_FWD] = f * lstm; flops4[LU_KC_CONV_FWD] = f * conv; flops4[LU_KC_DGRAD] = f * dgrad;
  flops4[LU_KC_WGRAD] = f * (lstm + conv);
  return 0;
}

// debug / test hook: copy one internal buffer of the conv named `name` to `out` as fp32 NHWC with the real channel
// count.  kind 0: activation output, 1: its gradient twin (after backward), 2: gate pre-activation gradient (lstm)
int lu_debug_buffer(lu_handle h, const char* name, int32_t kind, float* out, int64_t* shape4, void* stream) {
  LU_REQUIRE(h && h->bound && name && shape4, "null argument / unbound handle");
  for (auto& cv : h->convs) {
    if (cv.name != name) continue;
    int buf = cv.out_buf;
    if (kind == 1) { LU_REQUIRE(h->cfg.train, "no gradient buffers"); buf = buf >= 0 ? h->gidx[buf] : h->g_logits_buf; }
    if (kind == 2) buf = cv.dz_buf;
    LU_REQUIRE(buf >= 0, "no such buffer for %s", name);
    const ActBuf& a = h->acts[buf];
    shape4[0] = a.frames; shape4[1] = a.H; shape4[2] = a.W; shape4[3] = a.creal;
    if (out) {
      LuDebugRead r;
      r.src = reinterpret_cast<const uint16_t*>(h->ws + a.off); r.out = out; r.creal = a.creal; r.cpad = a.cpad; r.planes = a.planes; r.fmt = h->fmt;
      pf(h, (int64_t)a.frames * a.H * a.W * a.creal, stream, r);
    }
    return 0;
  }
  LU_FAIL("no conv named %s", name);
}

int lu_set_bn_sync_callback(lu_handle h, lu_bn_sync_fn fn, void* user, int32_t world_size) {
  LU_REQUIRE(h, "null handle");
  LU_REQUIRE(fn == nullptr || world_size >= 1, "world size must be >= 1");
  h->bn_sync_fn = fn; h->bn_sync_user = user; h->bn_sync_world = fn ? world_size : 1;
  return 0;
}

int lu_set_grad_bucket_callback(lu_handle h, lu_grad_bucket_fn fn, void* user) {
  LU_REQUIRE(h, "null handle");
  h->bucket_fn = fn; h->bucket_user = user;
  return 0;
}

int lu_loss_backward(lu_handle h, const float* dev_labels, const float* class_weights3, float* dev_loss, float* dev_grads,
                     void* stream) {
  return train_loss_backward(h, dev_labels, class_weights3, dev_loss, dev_grads, stream);
}

int lu_adam_step(lu_handle h, const float* dev_grads, float* dev_m, float* dev_v, float lr, float beta1, float beta2,
                 float eps, int64_t step, void* stream) {
  return train_adam(h, dev_grads, dev_m, dev_v, lr, beta1, beta2, eps, step, stream);
}

}  // extern "C"

#include "lu_post_host.inl"
#include "lu_aug_host.inl"

