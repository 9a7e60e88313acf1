// Host side of the training step (included by lu_api.cu): backward plan (data-gradient convolutions, buffers) and the
// reverse traversal of the layer graph.  Mirrors what tf.GradientTape + tape.gradient do for train2D.py:89-92.
//
// Gradient buffers are bf16 hi[/lo] planes with the same NHWC layout as the activations ("twin" ActBufs), so the data
// gradient of every convolution is just another table-driven implicit-GEMM convolution over the upstream gradient with
// transposed weight packing (LU_EPI_GRAD epilogue; first writer stores, later writers accumulate).

static int add_dgrad(lu_handle_s* h, int fi, int in_idx, int ry, int rx) {
  const bool x3 = h->planes == 2;
  ConvPlan d;
  {
    const ConvPlan& f = h->convs[fi];
    const ConvIn& fin = f.in[in_idx];
    const ParamT& wp = h->params[fin.w_param];
    const int k = f.k, s = f.stride;
    const int cin_total = (int)wp.shape[2], cout_total = (int)wp.shape[3];
    char nm[160];
    snprintf(nm, sizeof nm, "dgrad[%s in%d p%d%d]", f.name.c_str(), in_idx, ry, rx);
    d.name = nm; d.kind = LU_EPI_GRAD; d.k = k; d.stride = 1; d.fwd = fi; d.fwd_in = in_idx;
    const int gsrc = f.kind == LU_EPI_LSTM ? f.dz_buf : (f.out_buf >= 0 ? h->gidx[f.out_buf] : h->g_logits_buf);
    const ActBuf& gb = h->acts[gsrc];
    d.n_in = 1; d.in[0].buf = gsrc; d.in[0].creal = gb.creal; d.in[0].w_param = fin.w_param; d.in[0].c_base = 0;
    d.out_buf = h->gidx[fin.buf];
    d.cout = fin.creal; d.BN = pick_bn(d.cout); d.npad = ceil_to(d.cout, d.BN); d.n_tiles_n = d.npad / d.BN;
    d.cm.kind = LU_COL_IDENTITY; d.cm.n_real = d.cout;
    d.Hin = d.Hout = gb.H; d.Win = d.Wout = gb.W;
    d.oy_mul = s; d.oy_add = ry; d.ox_mul = s; d.ox_add = rx; d.OH = f.Hin; d.OW = f.Win;
    int ho, wo, pt, pl;
    tf_same(f.Hin, k, s, &ho, &pt);
    tf_same(f.Win, k, s, &wo, &pl);
    std::vector<int> kys, kxs, as, bs;
    for (int ky = 0; ky < k; ++ky) if (((ry + pt - ky) % s + s) % s == 0) { kys.push_back(ky); as.push_back(floordiv(ry + pt - ky, s)); }
    for (int kx = 0; kx < k; ++kx) if (((rx + pl - kx) % s + s) % s == 0) { kxs.push_back(kx); bs.push_back(floordiv(rx + pl - kx, s)); }
    if (kys.empty() || kxs.empty()) return -1;
    int a_min = as[0], a_max = as[0], b_min = bs[0], b_max = bs[0];
    for (int a : as) { a_min = a < a_min ? a : a_min; a_max = a > a_max ? a : a_max; }
    for (int b : bs) { b_min = b < b_min ? b : b_min; b_max = b > b_max ? b : b_max; }
    LuSrcView v; memset(&v, 0, sizeof v);
    const int ctot = gb.cpad * gb.planes;
    v.dimC = ctot; v.dimW = gb.W; v.dimP = 1; v.dimH = gb.H; v.dimN = gb.frames;
    v.sw = ctot; v.sh = (int64_t)gb.W * ctot; v.sp = v.sh; v.sn = (int64_t)gb.H * gb.W * ctot;
    v.frame_mul = 1; v.frame_add = 0;
    v.rows = LU_TILE_H + (a_max - a_min); v.pitch = LU_TILE_W + (b_max - b_min);
    d.n_views = 1; d.views[0] = v; d.view_buf[0] = gsrc;
    const int nchunks = gb.cpad / LU_KBLK;
    for (int ch = 0; ch < nchunks; ++ch) {
      int c_base, nvalid;
      if (f.kind == LU_EPI_LSTM) {
        const int per = f.fpad / LU_KBLK, gate = ch / per, cc = ch % per;
        c_base = gate * f.F + cc * LU_KBLK; nvalid = f.F - cc * LU_KBLK;
      } else { c_base = ch * LU_KBLK; nvalid = f.cout - ch * LU_KBLK; }
      nvalid = nvalid < 0 ? 0 : (nvalid > LU_KBLK ? LU_KBLK : nvalid);
      if (nvalid == 0) continue;
      for (int aplane = 0; aplane < (x3 ? 2 : 1); ++aplane) {
        const int nw = (x3 && aplane == 0) ? 2 : 1;
        LuAStage st; memset(&st, 0, sizeof st);
        st.src = 0; st.plane = 0; st.c = aplane * gb.cpad + ch * LU_KBLK;
        st.dy = (int16_t)a_min; st.dx = (int16_t)b_min; st.tap_begin = (uint32_t)d.taps.size(); st.ntaps = 0;
        for (int wpart = 0; wpart < nw; ++wpart)
          for (size_t iy = 0; iy < kys.size(); ++iy)
            for (size_t ix = 0; ix < kxs.size(); ++ix) {
              LuPackDesc pd; memset(&pd, 0, sizeof pd);
              pd.w_off = wp.offset; pd.tap_off = (kys[iy] * k + kxs[ix]) * cin_total * cout_total;
              pd.c_base = c_base; pd.n_valid = nvalid; pd.cin_total = cin_total; pd.cout_total = cout_total;
              pd.wpart = (int8_t)wpart; pd.kind = 0; pd.transposed = 1; pd.col_base = fin.c_base;
              d.packs.push_back(pd);
              d.taps.push_back((uint16_t)((as[iy] - a_min) * v.pitch + (bs[ix] - b_min)));
              st.ntaps++;
            }
        d.astages.push_back(st);
      }
    }
  }
  if (finish_tables(h, d)) return -2;
  d.macs_per_frame = 0;
  h->convs.push_back(d);
  return (int)h->convs.size() - 1;
}

static int new_act_raw(lu_handle_s* h, int frames, int H, int W, int creal) { return new_act(h, frames, H, W, creal); }

// called at the end of build_plan when cfg.train
static int build_train_plan(lu_handle_s* h) {
  const int N = h->cfg.batch * h->cfg.max_t;
  const int n_fwd_acts = (int)h->acts.size();
  h->gidx.assign(n_fwd_acts, -1);
  for (int i = 0; i < n_fwd_acts; ++i) {
    const ActBuf a = h->acts[i];
    h->gidx[i] = new_act_raw(h, a.frames, a.H, a.W, a.creal);
  }
  {
    const ConvPlan& lc = h->convs[h->logits_conv];
    h->g_logits_buf = new_act_raw(h, N, lc.Hout, lc.Wout, lc.cout);
  }
  const int n_fwd = (int)h->convs.size();
  for (int fi = 0; fi < n_fwd; ++fi)
    if (h->convs[fi].kind == LU_EPI_LSTM) {
      const int H = h->convs[fi].Hout, W = h->convs[fi].Wout, fp = h->convs[fi].fpad;
      const int b = new_act_raw(h, N, H, W, 4 * fp);
      h->convs[fi].dz_buf = b;
    }
  for (int fi = 0; fi < n_fwd; ++fi) {
    const int n_in = h->convs[fi].n_in, s = h->convs[fi].stride;
    for (int i = 0; i < n_in; ++i) {
      if (h->convs[fi].in[i].buf < 0 || h->convs[fi].in[i].buf == h->img_buf) continue;   // the image needs no gradient
      for (int ry = 0; ry < s; ++ry)
        for (int rx = 0; rx < s; ++rx) {
          const int di = add_dgrad(h, fi, i, ry, rx);
          if (di == -2) return 1;
          if (di >= 0) h->convs[fi].dgrads[i].push_back(di);
        }
    }
  }
  // K block -> (stage, tap) maps for the packed-space weight gradient
  for (int fi = 0; fi < n_fwd; ++fi) {
    ConvPlan& f = h->convs[fi];
    for (size_t s = 0; s < f.astages.size(); ++s)
      for (int t = 0; t < f.astages[s].ntaps; ++t) {
        f.kb_stage.push_back((uint16_t)s);
        f.kb_tap.push_back(f.taps[f.astages[s].tap_begin + t]);
      }
  }
  return 0;
}

static void train_layout(lu_handle_s* h, size_t& off) {
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  h->tr.off_loss_acc = take(64);
  if (!h->cfg.train) return;
  size_t dwp = 0;
  for (auto& cv : h->convs) {
    if (cv.kind == LU_EPI_GRAD) continue;
    cv.off_kb_stage = take(cv.kb_stage.size() * 2);
    cv.off_kb_tap = take(cv.kb_tap.size() * 2);
    cv.off_bwd_sums = take((size_t)cv.npad * 2 * 8);
    cv.off_bwd_means = take((size_t)cv.npad * 2 * 4);
    cv.off_wg_tasks = take((size_t)2 * LU_WG_MAX_TASKS * sizeof(LuWgTask));
    cv.off_wg_ptasks = take((size_t)2 * LU_WG_MAX_TASKS * sizeof(LuWgPairTask));
    const size_t b = (size_t)cv.npad * cv.ktot * 4;
    if (b > dwp) dwp = b;
    if (cv.kind == LU_EPI_LSTM) {
      const size_t px = (size_t)h->cfg.batch * cv.Hout * cv.Wout;
      cv.off_c_init = take(px * cv.fpad * 4);
      cv.off_dc = take(px * cv.fpad * 4);
    }
  }
  h->tr_off_dwp = take(dwp);
  h->tr_dwp_bytes = dwp;
}
static void train_destroy(lu_handle_s*) {}

static void train_upload(lu_handle_s* h, void* stream) {
  if (!h->cfg.train) return;
  for (auto& cv : h->convs) {
    if (cv.kind == LU_EPI_GRAD) continue;
    LU_H2D(h->ws + cv.off_kb_stage, cv.kb_stage.data(), cv.kb_stage.size() * 2, stream);
    LU_H2D(h->ws + cv.off_kb_tap, cv.kb_tap.data(), cv.kb_tap.size() * 2, stream);
  }
}

static uint16_t* act_ptr(lu_handle_s* h, int buf) { return reinterpret_cast<uint16_t*>(h->ws + h->acts[buf].off); }

// ---- pieces of the reverse traversal -------------------------------------------------------------------------------
static int run_dgrads(lu_handle_s* h, ConvPlan& f, int in_idx, int frames, int src_mul, int src_add, int out_mul,
                      int out_add, int force_acc, std::vector<char>& gwritten, void* stream) {
  if (f.in[in_idx].buf < 0) return 0;
  const int gdst = h->gidx[f.in[in_idx].buf];
  const int acc = force_acc >= 0 ? force_acc : (gwritten[gdst] ? 1 : 0);
  for (int di : f.dgrads[in_idx]) {
    ConvPlan& d = h->convs[di];
    const ActBuf& ob = h->acts[gdst];
    int mul[LU_MAX_SRC] = {src_mul, 1, 1, 1}, add[LU_MAX_SRC] = {src_add, 0, 0, 0};
    LuEpi e; memset(&e, 0, sizeof e);
    e.kind = LU_EPI_GRAD; e.H = d.Hout; e.W = d.Wout;
    e.oy_mul = d.oy_mul; e.oy_add = d.oy_add; e.ox_mul = d.ox_mul; e.ox_add = d.ox_add; e.OH = d.OH; e.OW = d.OW;
    e.accumulate = acc; e.out_frame_mul = out_mul; e.out_frame_add = out_add;
    e.out_act = act_ptr(h, gdst); e.out_cpad = ob.cpad; e.out_planes = ob.planes;
    e.bias = reinterpret_cast<const float*>(h->ws + d.off_bias);
    if (launch_conv(h, d, frames, mul, add, -1, e, stream)) return 1;
  }
  gwritten[gdst] = 1;
  return 0;
}

static void run_colsum(lu_handle_s* h, int gbuf, int frames_used, float* dst, int c_real, int gate_F, int gate_fpad,
                       void* stream) {
  const ActBuf& g = h->acts[gbuf];
  LuColSumGrad cs;
  cs.g = act_ptr(h, gbuf); cs.dst = dst; cs.cpad = g.cpad; cs.planes = g.planes;
  cs.c_real = c_real; cs.gate_F = gate_F; cs.gate_fpad = gate_fpad;
  rows(h, (int64_t)frames_used * g.H * g.W, g.cpad / 8, stream, cs);
}

// ---- tcgen05 weight gradient: task lists for one forward conv (see lu_wgrad_tc_kernel / lu_wgrad_pair_kernel) ----------
// nb_want: 0 = independent CTAs only; 1 / 2 = CTA-pair tasks (one or, where a source has >= 4 chunks, two 64-channel
// chunks per CTA) for every group of chunks and every full 256-column slab, independent tasks for what is left;
// 3 = CTA-pair tasks with one chunk per CTA whose accumulator entries are TAP PAIRS (N = 256: each CTA's B half is its
// window at two tap offsets, so the dY operand is read once per two taps).
static void build_wg_tasks(lu_handle_s* h, const ConvPlan& f, int frames, int only_src, int nb_want,
                           std::vector<LuWgTask>& out, std::vector<LuWgPairTask>& pout) {
  const bool x3 = h->planes == 2;
  if (x3) nb_want = 0;
  const int tiles = frames * ((f.Hout + LU_TILE_H - 1) / LU_TILE_H) * ((f.Wout + LU_TILE_W - 1) / LU_TILE_W);
  // K block index of the first tap of every stage
  std::vector<int> kb_begin(f.astages.size());
  { int kb = 0; for (size_t s = 0; s < f.astages.size(); ++s) { kb_begin[s] = kb; kb += f.astages[s].ntaps; } }
  auto valid_taps = [&](int s) {
    std::vector<int> t;
    for (int i = 0; i < f.astages[s].ntaps; ++i) if (f.packs[kb_begin[s] + i].wpart == 0) t.push_back(i);
    return t;
  };
  auto a_is_lo = [&](int s) {
    const LuAStage& st = f.astages[s];
    const int buf = f.view_buf[st.src];
    if (!x3 || buf < 0) return 0;
    const ActBuf& ab = h->acts[buf];
    const int ctot = ab.cpad * ab.planes;
    return ((st.c % ctot) >= ab.cpad) ? 1 : 0;
  };
  // groups of stages = the 64-channel chunks of one source that share window geometry and tap list
  struct Group { std::vector<int> stages; std::vector<int> taps; int a_is_lo; };
  std::vector<Group> groups;
  std::vector<char> used(f.astages.size(), 0);
  for (size_t s = 0; s < f.astages.size(); ++s) {
    if (used[s]) continue;
    const LuAStage& a = f.astages[s];
    if (only_src >= 0 && a.src != only_src) continue;
    Group g; g.stages.push_back((int)s); g.taps = valid_taps((int)s); g.a_is_lo = a_is_lo((int)s);
    used[s] = 1;
    for (size_t s2 = s + 1; s2 < f.astages.size(); ++s2) {
      if (used[s2]) continue;
      const LuAStage& c = f.astages[s2];
      if (c.src != a.src || c.plane != a.plane || c.dy != a.dy || c.dx != a.dx || a_is_lo((int)s2) != g.a_is_lo) continue;
      std::vector<int> t2 = valid_taps((int)s2);
      if (t2.size() != g.taps.size()) continue;
      bool same = true;
      for (size_t i = 0; i < t2.size() && same; ++i)
        same = f.taps[c.tap_begin + t2[i]] == f.taps[a.tap_begin + g.taps[i]];
      if (!same) continue;
      g.stages.push_back((int)s2); used[s2] = 1;
    }
    if (!g.taps.empty()) groups.push_back(g);
  }
  const int n_chunks = f.kind == LU_EPI_LSTM ? f.npad / 64 : ceil_to(f.cout, 64) / 64;
  auto ychan_of = [&](int ci) { return f.kind == LU_EPI_LSTM ? (ci % 4) * f.fpad + (ci / 4) * 64 : ci * 64; };
  const int full_slabs = nb_want > 0 ? n_chunks / 4 : 0;            // 256-column slabs the pair kernel takes
  // units: what one task (family) covers in the channel-row direction
  struct Unit { int st[4]; int n, nb; const Group* g; };             // pair: n = 2 * nb stages; single: n = 1 or 2, nb = 0
  std::vector<Unit> punits, sunits_all, sunits_rest;                  // pair units; single units over ALL columns / over the remainder columns
  for (auto& g : groups) {
    size_t i = 0;
    const size_t ns = g.stages.size();
    if (full_slabs > 0) {
      while (ns - i >= 2) {
        const int nb = (nb_want == 2 && ns - i >= 4) ? 2 : 1;
        Unit u; u.n = 2 * nb; u.nb = nb; u.g = &g;
        for (int k = 0; k < 4; ++k) u.st[k] = k < u.n ? g.stages[i + k] : -1;
        punits.push_back(u);
        for (int k = 0; k < u.n; k += 2) {                            // the same stages for the remainder columns
          Unit r; r.n = 2; r.nb = 0; r.g = &g; r.st[0] = g.stages[i + k]; r.st[1] = g.stages[i + k + 1]; r.st[2] = r.st[3] = -1;
          sunits_rest.push_back(r);
        }
        i += u.n;
      }
    }
    for (; i < ns; i += 2) {
      Unit r; r.nb = 0; r.g = &g; r.st[0] = g.stages[i]; r.st[1] = i + 1 < ns ? g.stages[i + 1] : -1; r.n = r.st[1] >= 0 ? 2 : 1;
      r.st[2] = r.st[3] = -1;
      sunits_all.push_back(r);
    }
  }
  // column slabs of the independent tasks: 64-column chunks that exist in dY
  const int nch_max = x3 ? 1 : 2;
  std::vector<std::pair<int, int>> slabs_all, slabs_rest;             // (first chunk, number of chunks)
  for (int c = 0; c < n_chunks; c += nch_max) slabs_all.push_back({c, (n_chunks - c) < nch_max ? (n_chunks - c) : nch_max});
  for (int c = full_slabs * 4; c < n_chunks; c += nch_max) slabs_rest.push_back({c, (n_chunks - c) < nch_max ? (n_chunks - c) : nch_max});
  auto n_tap_tasks = [&](size_t ntaps, int per) { return (int)((ntaps + per - 1) / per); };
  int base_ctas = 0;
  for (auto& u : punits) base_ctas += 2 * n_tap_tasks(u.g->taps.size(), u.nb == 2 ? 2 : 4) * full_slabs;
  for (auto& u : sunits_all) base_ctas += n_tap_tasks(u.g->taps.size(), 4) * (int)slabs_all.size();
  for (auto& u : sunits_rest) base_ctas += n_tap_tasks(u.g->taps.size(), 4) * (int)slabs_rest.size();
  // enough tasks to fill the machine, and pixel ranges small enough (~256 tiles = 32k pixels) that the range's
  // activations + gradients stay L2-resident while the wave of tasks sharing it runs
  static int range_env = -1;                                         // LU_WGRAD_RANGE: pixel tiles per task (experiment switch)
  if (range_env < 0) { const char* ce = getenv("LU_WGRAD_RANGE"); range_env = ce && atoi(ce) > 0 ? atoi(ce) : 256; }
  int split = (4 * h->num_sms + base_ctas - 1) / (base_ctas > 0 ? base_ctas : 1);
  if (split < (tiles + range_env - 1) / range_env) split = (tiles + range_env - 1) / range_env;
  if (base_ctas > 0 && (int64_t)split * base_ctas > LU_WG_MAX_TASKS) split = LU_WG_MAX_TASKS / base_ctas;
  if (split < 1) split = 1;
  if (split > tiles) split = tiles;
  auto tap_range = [&](size_t ntaps, int nt, int ti_, int& t0, int& cnt) {   // taps of a unit spread evenly over its nt tasks
    const int per = (int)ntaps / nt, extra = (int)ntaps % nt;
    t0 = ti_ * per + (ti_ < extra ? ti_ : extra); cnt = per + (ti_ < extra ? 1 : 0);
  };
  auto emit_single = [&](const Unit& u, const std::vector<std::pair<int, int>>& slabs, int sp) {
    const Group& g = *u.g;
    const int nt = n_tap_tasks(g.taps.size(), 4);
    for (auto& sl : slabs)
      for (int ti_ = 0; ti_ < nt; ++ti_) {
        int t0, cnt; tap_range(g.taps.size(), nt, ti_, t0, cnt);
        LuWgTask tk; memset(&tk, 0, sizeof tk);
        tk.stage0 = (int16_t)u.st[0]; tk.stage1 = (int16_t)u.st[1]; tk.a_is_lo = (int16_t)g.a_is_lo;
        tk.ntaps = (int16_t)cnt;
        for (int i = 0; i < cnt; ++i) {
          const int ti = g.taps[t0 + i];
          tk.off[i] = f.taps[f.astages[u.st[0]].tap_begin + ti];
          tk.kb0[i] = kb_begin[u.st[0]] + ti;
          tk.kb1[i] = u.st[1] >= 0 ? kb_begin[u.st[1]] + ti : -1;
        }
        tk.n0 = sl.first * 64; tk.nch = sl.second;
        for (int c = 0; c < sl.second; ++c) tk.ychan[c] = ychan_of(sl.first + c);
        tk.tile0 = (int)((int64_t)tiles * sp / split); tk.tile1 = (int)((int64_t)tiles * (sp + 1) / split);
        if (tk.tile1 > tk.tile0) out.push_back(tk);
      }
  };
  // Pixel split OUTERMOST: the tasks resident at any time then stream the SAME pixel range (different channel rows / taps /
  // column slabs), so activations and upstream gradients are fetched from DRAM once per range instead of once per task
  // (measured before the reorder: 84 GB of DRAM reads for a 6 GB working set).
  for (int sp = 0; sp < split; ++sp) {
    for (auto& u : punits) {
      const Group& g = *u.g;
      const int nt = n_tap_tasks(g.taps.size(), u.nb == 2 ? 2 : 4);
      if (nb_want == 3) {
        // tap pairs: consecutive taps of the list (window offsets increase along it), a last single one if their number is odd;
        // two entries (2 x 256 accumulator columns) per task
        std::vector<std::pair<int, int>> ent;
        for (size_t i = 0; i < g.taps.size(); i += 2) ent.push_back({g.taps[i], i + 1 < g.taps.size() ? g.taps[i + 1] : -1});
        const int nte = n_tap_tasks(ent.size(), 2);                  // == nt
        const uint32_t tb = f.astages[u.st[0]].tap_begin;
        for (int sl = 0; sl < full_slabs; ++sl)
          for (int ti_ = 0; ti_ < nte; ++ti_) {
            int t0, cnt; tap_range(ent.size(), nte, ti_, t0, cnt);
            LuWgPairTask tk; memset(&tk, 0, sizeof tk);
            for (int k = 0; k < 4; ++k) tk.stage[k] = (int16_t)u.st[k];
            tk.nb = 1; tk.ntaps = (int16_t)cnt;
            for (int i = 0; i < cnt; ++i) {
              const int ta = ent[t0 + i].first, tb2 = ent[t0 + i].second;
              tk.off[i] = f.taps[tb + ta];
              if (tb2 >= 0) {
                tk.disp[i] = (int)f.taps[tb + tb2] - (int)f.taps[tb + ta];
                tk.kb[i][0] = kb_begin[u.st[0]] + ta; tk.kb[i][1] = kb_begin[u.st[0]] + tb2;
                tk.kb[i][2] = kb_begin[u.st[1]] + ta; tk.kb[i][3] = kb_begin[u.st[1]] + tb2;
              } else {
                tk.kb[i][0] = kb_begin[u.st[0]] + ta; tk.kb[i][1] = kb_begin[u.st[1]] + ta;
              }
            }
            tk.n0 = sl * 256;
            for (int c = 0; c < 4; ++c) tk.ychan[c] = ychan_of(sl * 4 + c);
            tk.tile0 = (int)((int64_t)tiles * sp / split); tk.tile1 = (int)((int64_t)tiles * (sp + 1) / split);
            if (tk.tile1 > tk.tile0) pout.push_back(tk);
          }
        continue;
      }
      for (int sl = 0; sl < full_slabs; ++sl)
        for (int ti_ = 0; ti_ < nt; ++ti_) {
          int t0, cnt; tap_range(g.taps.size(), nt, ti_, t0, cnt);
          LuWgPairTask tk; memset(&tk, 0, sizeof tk);
          for (int k = 0; k < 4; ++k) tk.stage[k] = (int16_t)u.st[k];
          tk.nb = (int16_t)u.nb; tk.ntaps = (int16_t)cnt;
          for (int i = 0; i < cnt; ++i) {
            const int ti = g.taps[t0 + i];
            tk.off[i] = f.taps[f.astages[u.st[0]].tap_begin + ti];
            for (int k = 0; k < u.n; ++k) tk.kb[i][k] = kb_begin[u.st[k]] + ti;
          }
          tk.n0 = sl * 256;
          for (int c = 0; c < 4; ++c) tk.ychan[c] = ychan_of(sl * 4 + c);
          tk.tile0 = (int)((int64_t)tiles * sp / split); tk.tile1 = (int)((int64_t)tiles * (sp + 1) / split);
          if (tk.tile1 > tk.tile0) pout.push_back(tk);
        }
    }
    for (auto& u : sunits_rest) emit_single(u, slabs_rest, sp);
    for (auto& u : sunits_all) emit_single(u, slabs_all, sp);
  }
}

#ifdef LU_HOST_EMU
// TEST-ONLY (host build): replays the task lists with scalar loops, statement by statement what the two tcgen05 kernels do
// with them -- windows staged as flat [rows * pitch] arrays with the tensor map's zero fill, taps as row offsets into
// them, 128-pixel tiles, dY boxes per column chunk / plane, accumulators per tap, flush into the packed gradient; for a
// pair task the transposed product (rows = the CTA's 128 output channels, columns = the stages of both CTAs).  Lets the CPU
// suite check the task builder against the scalar mirror (tests/test_emu_wgrad_tasks.py).
static void emu_stage_window(const LuAStage& st, const LuSrcView& v, int frame, int y0, int x0, std::vector<float>& dst) {
  dst.assign((size_t)v.rows * v.pitch * 64, 0.f);
  const int64_t n = (int64_t)frame * v.frame_mul + v.frame_add;
  if (n < 0 || n >= v.dimN) return;
  for (int wy = 0; wy < v.rows; ++wy)
    for (int wx = 0; wx < v.pitch; ++wx) {
      const int yy = y0 + wy, xx = x0 + wx;
      if (yy < 0 || yy >= v.dimH || xx < 0 || xx >= v.dimW) continue;
      const uint16_t* a = v.ptr + n * v.sn + (int64_t)yy * v.sh + (int64_t)st.plane * v.sp + (int64_t)xx * v.sw + st.c;
      for (int kk = 0; kk < 64 && st.c + kk < v.dimC; ++kk) dst[((size_t)wy * v.pitch + wx) * 64 + kk] = lu_bf2f(a[kk]);
    }
}

static void emulate_wg_tasks(const LuWgradMirror& w, const std::vector<LuWgTask>& tasks, int tiles_x, int tiles_y) {
  const LuConvParams& cp = w.p;
  const int tiles_per_frame = tiles_x * tiles_y;
  std::vector<float> win[2], D;
  for (const LuWgTask& tk : tasks) {
    const LuAStage st0 = cp.astages[tk.stage0];
    const LuAStage st1 = cp.astages[tk.stage1 >= 0 ? tk.stage1 : tk.stage0];
    const LuSrcView& v = cp.src[st0.src];
    const int N = tk.nch * 64;
    const int nyp = (w.dy_planes == 2 && !tk.a_is_lo) ? 2 : 1;
    D.assign((size_t)tk.ntaps * 128 * N, 0.f);
    bool any = false;
    for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
      const int frame = tile / tiles_per_frame, rem = tile % tiles_per_frame;
      if (st0.src == w.skip_t0_src && (frame % w.T) == 0) continue;
      any = true;
      const int y0 = (rem / tiles_x) * LU_TILE_H, x0 = (rem % tiles_x) * LU_TILE_W;
      emu_stage_window(st0, v, frame, y0 + st0.dy, x0 + st0.dx, win[0]);
      if (tk.stage1 >= 0) emu_stage_window(st1, cp.src[st1.src], frame, y0 + st1.dy, x0 + st1.dx, win[1]);
      const int64_t fy = (int64_t)frame * w.dy_frame_mul + w.dy_frame_add;
      for (int ti = 0; ti < tk.ntaps; ++ti)
        for (int m = 0; m < 128; ++m) {
          const int ty = m / LU_TILE_W, tx = m % LU_TILE_W;
          const int y = y0 + ty, x = x0 + tx;
          if (y >= w.H || x >= w.W) continue;                      // dY box: zero fill outside the frame
          const size_t wi = (size_t)tk.off[ti] + (size_t)ty * v.pitch + tx;
          const uint16_t* gy = w.dY + ((fy * w.H + y) * w.W + x) * (int64_t)(w.dy_cpad * w.dy_planes);
          for (int row = 0; row < 128; ++row) {
            if (row >= 64 && tk.stage1 < 0) break;
            const float a = win[row >> 6][wi * 64 + (row & 63)];
            if (a == 0.f) continue;
            float* d = &D[((size_t)ti * 128 + row) * N];
            for (int dp = 0; dp < nyp; ++dp)
              for (int c = 0; c < tk.nch; ++c) {
                const uint16_t* g = gy + tk.ychan[c] + dp * w.dy_cpad;
                for (int j = 0; j < 64; ++j) d[c * 64 + j] += a * lu_bf2f(g[j]);
              }
          }
        }
    }
    if (!any) continue;
    for (int ti = 0; ti < tk.ntaps; ++ti)
      for (int row = 0; row < 128; ++row) {
        const int kb = row >= 64 ? (tk.stage1 >= 0 ? tk.kb1[ti] : -1) : tk.kb0[ti];
        if (kb < 0) continue;
        for (int col = 0; col < N; ++col)
          w.dWp[(int64_t)(tk.n0 + col) * cp.ktot + (int64_t)kb * LU_KBLK + (row & 63)] += D[((size_t)ti * 128 + row) * N + col];
      }
  }
}

static void emulate_wg_pair_tasks(const LuWgradMirror& w, const std::vector<LuWgPairTask>& tasks, int tiles_x, int tiles_y) {
  const LuConvParams& cp = w.p;
  const int tiles_per_frame = tiles_x * tiles_y;
  std::vector<float> win[4], D;
  for (const LuWgPairTask& tk : tasks) {
    const int ns = 2 * tk.nb, N = 64 * ns;
    const LuAStage st0 = cp.astages[tk.stage[0]];
    const LuSrcView& v = cp.src[st0.src];
    // accumulator column group (64 columns) k of entry ti: which staged window it reads and at which extra row offset
    // (two-window form: window k; tap-pair form: CTA k >> 1's window, second tap for odd k)
    auto width = [&](int ti) { return tk.disp[ti] != 0 ? 256 : N; };
    auto win_of = [&](int ti, int k) { return tk.disp[ti] != 0 ? (k >> 1) : k; };
    auto extra_of = [&](int ti, int k) { return tk.disp[ti] != 0 ? (size_t)(k & 1) * (size_t)tk.disp[ti] : (size_t)0; };
    for (int crank = 0; crank < 2; ++crank) {                        // the two CTAs: 128 output channels each
      D.assign((size_t)tk.ntaps * 128 * 256, 0.f);
      bool any = false;
      for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
        const int frame = tile / tiles_per_frame, rem = tile % tiles_per_frame;
        if (st0.src == w.skip_t0_src && (frame % w.T) == 0) continue;
        any = true;
        const int y0 = (rem / tiles_x) * LU_TILE_H, x0 = (rem % tiles_x) * LU_TILE_W;
        for (int k = 0; k < ns; ++k) {                               // B operand: the windows of BOTH CTAs, in column order
          const LuAStage st = cp.astages[tk.stage[k]];
          emu_stage_window(st, cp.src[st.src], frame, y0 + st.dy, x0 + st.dx, win[k]);
        }
        const int64_t fy = (int64_t)frame * w.dy_frame_mul + w.dy_frame_add;
        for (int ti = 0; ti < tk.ntaps; ++ti)
          for (int m = 0; m < 128; ++m) {
            const int ty = m / LU_TILE_W, tx = m % LU_TILE_W;
            const int y = y0 + ty, x = x0 + tx;
            if (y >= w.H || x >= w.W) continue;
            const size_t wi = (size_t)tk.off[ti] + (size_t)ty * v.pitch + tx;
            const uint16_t* gy = w.dY + ((fy * w.H + y) * w.W + x) * (int64_t)(w.dy_cpad * w.dy_planes);
            for (int row = 0; row < 128; ++row) {                    // A operand: this CTA's two dY chunks
              const float g = lu_bf2f(gy[tk.ychan[2 * crank + (row >> 6)] + (row & 63)]);
              if (g == 0.f) continue;
              float* d = &D[((size_t)ti * 128 + row) * 256];
              for (int col = 0; col < width(ti); ++col)
                d[col] += g * win[win_of(ti, col >> 6)][(wi + extra_of(ti, col >> 6)) * 64 + (col & 63)];
            }
          }
      }
      if (!any) continue;
      for (int ti = 0; ti < tk.ntaps; ++ti)
        for (int row = 0; row < 128; ++row)
          for (int col = 0; col < width(ti); ++col)
            w.dWp[(int64_t)(tk.n0 + crank * 128 + row) * cp.ktot + (int64_t)tk.kb[ti][col >> 6] * LU_KBLK + (col & 63)] +=
                D[((size_t)ti * 128 + row) * 256 + col];
    }
  }
}
#endif

// weight gradient of forward conv f from the upstream gradient buffer gbuf, in packed space:
// tcgen05 kernel (product) or the scalar mirror (engine=simt / host test build)
static int run_wgrad(lu_handle_s* h, ConvPlan& f, int gbuf, int T, float* grads, void* stream) {
  const ActBuf& g = h->acts[gbuf];
  float* dwp = reinterpret_cast<float*>(h->ws + h->tr_off_dwp);
  LU_MEMSET(dwp, 0, (size_t)f.npad * f.ktot * 4, stream);
  const int n_launch = f.kind == LU_EPI_LSTM ? 2 : 1;
  // LU_WGRAD_ENGINE=simt forces the scalar engine for the weight gradient only (test hook: same forward, same
  // upstream gradients, so the two weight-gradient engines can be compared without activation-kink noise)
  const char* wg_env = getenv("LU_WGRAD_ENGINE");
  const bool use_tc = h->cfg.engine == LU_ENGINE_TCGEN05 && h->cfg.a_mode == LU_AMODE_HALO &&
                      !(wg_env && strcmp(wg_env, "simt") == 0);
  // LU_WGRAD_PAIR = chunks per CTA of the CTA-pair kernel (lu_wgrad_pair_kernel: transposed product, one M = 256 MMA per
  // pair): 1 (default), 2 (two chunks where a source has >= 4: fewer shared-memory reads per MMA but as much L2 -> SM
  // traffic as the independent form), 0 (independent CTAs only).  Measured on B200 (round 2, C3 train step, weight-gradient
  // launches per step): 0 -> 88.5 ms, 1 -> 83.2 ms, 2 -> 84.0 ms.  3 = one chunk per CTA with TAP-PAIR accumulator entries (N = 256:
  // the dY operand is read from shared memory once per two taps -- 64 instead of 96 B/clk of operand reads per SM).
  static int wg_pair_env = -1;
  if (wg_pair_env < 0) { const char* ce = getenv("LU_WGRAD_PAIR"); wg_pair_env = ce ? atoi(ce) : 1; }
  const int nb_want = h->planes == 1 ? wg_pair_env : 0;
  for (int pass = 0; pass < n_launch; ++pass) {
    LuWgradMirror w; memset(&w, 0, sizeof w);
    for (int i = 0; i < f.n_views; ++i) {
      w.p.src[i] = f.views[i];
      w.p.src[i].ptr = view_ptr(h, f.view_buf[i]);
      w.p.src[i].frame_mul = 1; w.p.src[i].frame_add = 0;
    }
    w.p.astages = reinterpret_cast<const LuAStage*>(h->ws + f.off_astages);
    w.p.ktot = f.ktot; w.p.n_astages = (int)f.astages.size();
    w.kb_stage = reinterpret_cast<const uint16_t*>(h->ws + f.off_kb_stage);
    w.kb_tap = reinterpret_cast<const uint16_t*>(h->ws + f.off_kb_tap);
    w.dY = act_ptr(h, gbuf); w.dy_cpad = g.cpad; w.dy_planes = g.planes; w.dy_frame_mul = 1; w.dy_frame_add = 0;
    w.cm = f.cm; w.gate_fpad = f.fpad; w.dWp = dwp; w.npad = f.npad; w.H = f.Hout; w.W = f.Wout;
    w.frames = h->cfg.batch * T; w.chunk = 2048; w.only_src = -1; w.skip_t0_src = -1; w.T = T;
    if (f.kind == LU_EPI_LSTM) {
      if (pass == 0) {                 // all frames; h_{t-1} = h sequence shifted by one frame, nothing for t == 0
        w.p.src[1].frame_add = -1; w.skip_t0_src = 1;
      } else {                         // t == 0 frames against the initial state h_init (state buffer before the call)
        w.frames = h->cfg.batch; w.only_src = 1;
        w.p.src[1].ptr = reinterpret_cast<const uint16_t*>(h->ws + f.off_hstate[h->hcur ^ 1]);
        w.p.src[1].dimN = h->cfg.batch;
        w.dy_frame_mul = T; w.dy_frame_add = 0;
      }
    }
#ifndef LU_HOST_EMU
    if (use_tc) {
      // the task lists depend only on (conv, T, pass): built and uploaded once, then reused every step
      LuWgTask* dtasks = reinterpret_cast<LuWgTask*>(h->ws + f.off_wg_tasks) + (size_t)pass * LU_WG_MAX_TASKS;
      LuWgPairTask* dptasks = reinterpret_cast<LuWgPairTask*>(h->ws + f.off_wg_ptasks) + (size_t)pass * LU_WG_MAX_TASKS;
      if (f.wg_cached_T[pass] != T) {
        std::vector<LuWgTask> tasks;
        std::vector<LuWgPairTask> ptasks;
        build_wg_tasks(h, f, w.frames, w.only_src, nb_want, tasks, ptasks);
        LU_REQUIRE(tasks.size() <= (size_t)LU_WG_MAX_TASKS && ptasks.size() <= (size_t)LU_WG_MAX_TASKS,
                   "too many weight-gradient tasks (%zu + %zu)", tasks.size(), ptasks.size());
        cudaError_t e = cudaSuccess;
        if (!tasks.empty())
          e = cudaMemcpyAsync(dtasks, tasks.data(), tasks.size() * sizeof(LuWgTask), cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e == cudaSuccess && !ptasks.empty())
          e = cudaMemcpyAsync(dptasks, ptasks.data(), ptasks.size() * sizeof(LuWgPairTask), cudaMemcpyHostToDevice, (cudaStream_t)stream);
        LU_REQUIRE(e == cudaSuccess, "task upload: %s", cudaGetErrorString(e));
        e = cudaStreamSynchronize((cudaStream_t)stream);            // the host vectors die at scope exit
        LU_REQUIRE(e == cudaSuccess, "task upload sync: %s", cudaGetErrorString(e));
        f.wg_cached_T[pass] = T; f.wg_n_tasks[pass] = (int)tasks.size(); f.wg_n_ptasks[pass] = (int)ptasks.size();
      }
      const int n_tasks = f.wg_n_tasks[pass], n_ptasks = f.wg_n_ptasks[pass];
      if (n_tasks == 0 && n_ptasks == 0) continue;
      cudaError_t e;
      LuWgParams wp; memset(&wp, 0, sizeof wp);
      for (int i = 0; i < f.n_views; ++i) wp.tmA[i] = f.tmA[i];
      if (pass == 1) wp.tmA[1] = f.tmHstate[h->hcur ^ 1];
      wp.tmY = h->acts_tm[gbuf];
      wp.cp = w.p;
      wp.tasks = dtasks; wp.ptasks = dptasks; wp.dWp = dwp;
      wp.tiles_x = (f.Wout + LU_TILE_W - 1) / LU_TILE_W; wp.tiles_y = (f.Hout + LU_TILE_H - 1) / LU_TILE_H;
      wp.T = T; wp.skip_t0_src = w.skip_t0_src;
      wp.dy_frame_mul = (int)w.dy_frame_mul; wp.dy_frame_add = (int)w.dy_frame_add; wp.dy_planes = g.planes; wp.dy_cpad = g.cpad;
      wp.a_win_bytes = f.a_bytes;
      static int red4_env = -1;
      if (red4_env < 0) { const char* ce = getenv("LU_WGRAD_RED4"); red4_env = ce ? atoi(ce) : 1; }
      wp.flush_scalar = red4_env == 0 ? 1 : 0;
      const int budget = 232448 - 1024 - 256;
      bool& attr = h->wg_attr_set;
      if (!attr) {
        e = cudaFuncSetAttribute(lu_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        LU_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        e = cudaFuncSetAttribute(lu_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        LU_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr = true;
      }
      if (n_ptasks > 0) {
        // pair kernel: a stage = this CTA's two 16 KB dY chunks + up to two activation windows
        wp.stage_bytes = 32768 + 2 * f.a_bytes;
        wp.n_stages = budget / wp.stage_bytes;
        if (wp.n_stages > 4) wp.n_stages = 4;
        LU_REQUIRE(wp.n_stages >= 1, "weight-gradient stage does not fit shared memory");
        h->launches++;
        time_begin(h, LU_KC_WGRAD, stream);
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3((unsigned)(2 * n_ptasks)); lc.blockDim = dim3(256);
        lc.dynamicSmemBytes = (size_t)wp.n_stages * wp.stage_bytes + 1024 + 256;
        lc.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        e = cudaLaunchKernelEx(&lc, lu_wgrad_pair_kernel, wp);
        time_end(h, stream);
        LU_REQUIRE(e == cudaSuccess, "wgrad pair launch (%s) failed: %s", f.name.c_str(), cudaGetErrorString(e));
      }
      if (n_tasks > 0) {
        // stage = two activation windows + two 16 KB dY boxes (bf16: 2 column chunks; bf16x3: hi and lo plane of one chunk)
        wp.stage_bytes = 2 * f.a_bytes + 2 * 16384;
        wp.n_stages = budget / wp.stage_bytes;
        if (wp.n_stages > 4) wp.n_stages = 4;
        LU_REQUIRE(wp.n_stages >= 1, "weight-gradient stage does not fit shared memory");
        h->launches++;
        time_begin(h, LU_KC_WGRAD, stream);
        lu_wgrad_tc_kernel<<<(unsigned)n_tasks, 256, (size_t)wp.n_stages * wp.stage_bytes + 1024 + 256, (cudaStream_t)stream>>>(wp);
        time_end(h, stream);
      }
      e = cudaGetLastError();
      LU_REQUIRE(e == cudaSuccess, "wgrad launch (%s) failed: %s", f.name.c_str(), cudaGetErrorString(e));
      continue;
    }
#endif
#ifdef LU_HOST_EMU
    // TEST-ONLY: LU_WGRAD_EMU_TASKS=0|1|2 replays the task lists the tcgen05 kernels would get with that LU_WGRAD_PAIR
    if (const char* te = getenv("LU_WGRAD_EMU_TASKS")) {
      const int mode = atoi(te);
      if (mode >= 0 && mode <= 3 && h->cfg.a_mode == LU_AMODE_HALO) {
        std::vector<LuWgTask> tasks;
        std::vector<LuWgPairTask> ptasks;
        build_wg_tasks(h, f, w.frames, w.only_src, h->planes == 1 ? mode : 0, tasks, ptasks);
        const int tx = (f.Wout + LU_TILE_W - 1) / LU_TILE_W, ty = (f.Hout + LU_TILE_H - 1) / LU_TILE_H;
        emulate_wg_pair_tasks(w, ptasks, tx, ty);
        emulate_wg_tasks(w, tasks, tx, ty);
        if (getenv("LU_WGRAD_EMU_VERBOSE")) fprintf(stderr, "wgrad tasks %s pass %d: %zu pair, %zu single\n", f.name.c_str(), pass, ptasks.size(), tasks.size());
        continue;
      }
    }
#endif
    const int64_t npix = (int64_t)w.frames * w.H * w.W;
    const int64_t nchunk = (npix + w.chunk - 1) / w.chunk;
    pf(h, (int64_t)f.ktot * (f.npad / 16) * nchunk, stream, w);
  }
  LuUnpackWgrad u;
  u.dWp = dwp; u.descs = reinterpret_cast<const LuPackDesc*>(h->ws + f.off_packs); u.grads = grads; u.cm = f.cm; u.ktot = f.ktot;
  pf(h, (int64_t)f.npad * f.ktot, stream, u);
  return 0;
}

static int bwd_conv_layer(lu_handle_s* h, ConvPlan& f, int T, float* grads, std::vector<char>& gwritten, void* stream) {
  const int N = h->cfg.batch * T;
  const int gbuf = f.out_buf >= 0 ? h->gidx[f.out_buf] : h->g_logits_buf;
  const int64_t npix = (int64_t)N * f.Hout * f.Wout;
  if (f.has_bn) {
    const ActBuf& gb = h->acts[gbuf];
    double* sums = reinterpret_cast<double*>(h->ws + f.off_bwd_sums);
    LU_MEMSET(sums, 0, (size_t)f.raw_cpad * 2 * 8, stream);
    LuBnBwdReduce r;
    r.dA = act_ptr(h, gbuf); r.raw = reinterpret_cast<const float*>(h->ws + f.off_raw);
    r.scale = reinterpret_cast<const float*>(h->ws + f.off_bscale); r.shift = reinterpret_cast<const float*>(h->ws + f.off_bshift);
    r.mean = reinterpret_cast<const float*>(h->ws + f.off_save_mean); r.invstd = reinterpret_cast<const float*>(h->ws + f.off_save_invstd);
    r.sums = sums; r.cpad = gb.cpad; r.planes = gb.planes; r.raw_cpad = f.raw_cpad; r.c_real = f.cout;
    r.alpha = h->cfg.lrelu_alpha;
    rows(h, npix, f.raw_cpad / 8, stream, r);
    LuBnBwdParams bp;
    bp.sums = sums; bp.dgamma = grads + h->params[f.gamma].offset; bp.dbeta = grads + h->params[f.beta].offset;
    bp.raw_cpad = f.raw_cpad; bp.c_real = f.cout; bp.npix = npix; bp.write_grads = 1;
    bp.means = reinterpret_cast<float*>(h->ws + f.off_bwd_means);
    pf(h, f.raw_cpad, stream, bp);
    if (h->bn_sync_fn) {
      // synchronised BN: gamma / beta gradients stay local (the gradient all-reduce averages them); the input gradient
      // needs the means of g and g * xhat over the GLOBAL batch
      h->bn_sync_fn(sums, (int64_t)2 * f.raw_cpad, h->bn_sync_user);
      bp.write_grads = 0; bp.npix = npix * h->bn_sync_world;
      pf(h, f.raw_cpad, stream, bp);
    }
    LuBnBwdApply a;
    a.dA = act_ptr(h, gbuf); a.raw = r.raw; a.scale = r.scale; a.shift = r.shift; a.mean = r.mean; a.invstd = r.invstd;
    a.means = bp.means; a.cpad = gb.cpad; a.planes = gb.planes; a.raw_cpad = f.raw_cpad; a.c_real = f.cout; a.alpha = h->cfg.lrelu_alpha;
    rows(h, npix, gb.cpad / 8, stream, a);
  }
  // a conv bias in front of a training-mode BatchNorm has an exactly zero gradient (BN removes the batch mean; the
  // gradient buffer was zero-filled): only the un-normalised logits conv needs the column sum
  if (!f.has_bn) run_colsum(h, gbuf, N, grads + h->params[f.bias_param].offset, f.cout, 0, 0, stream);
  if (run_wgrad(h, f, gbuf, T, grads, stream)) return 1;
  for (int i = 0; i < f.n_in; ++i)
    if (run_dgrads(h, f, i, N, 1, 0, 1, 0, -1, gwritten, stream)) return 1;
  return 0;
}

static int bwd_lstm_layer(lu_handle_s* h, ConvPlan& f, int T, float* grads, std::vector<char>& gwritten, void* stream) {
  const int B = h->cfg.batch;
  const int gH = h->gidx[f.hseq_buf];
  const ActBuf& gh = h->acts[gH];
  const int64_t pps = (int64_t)f.Hout * f.Wout;
  if (!gwritten[gH]) {                      // no consumer wrote a gradient (cannot happen in ULSTMnet2D): zero it
    LU_MEMSET(act_ptr(h, gH), 0, gh.bytes(), stream);
    gwritten[gH] = 1;
  }
  for (int t = T - 1; t >= 0; --t) {
    LuLstmCellBwd c;
    c.dH = act_ptr(h, gH); c.gates = reinterpret_cast<const uint16_t*>(h->ws + f.off_save_gates);
    c.c_t = reinterpret_cast<const float*>(h->ws + f.off_save_c);
    c.c_prev_is_init = t == 0;
    c.c_prev = t == 0 ? reinterpret_cast<const float*>(h->ws + f.off_c_init) : c.c_t;
    c.dC = reinterpret_cast<float*>(h->ws + f.off_dc); c.dZ = act_ptr(h, f.dz_buf);
    c.pix_per_sample = pps; c.T = T; c.t = t; c.fpad = f.fpad; c.planes = h->planes; c.gate_kind = h->cfg.gate;
    c.first = t == T - 1;
    pf(h, (int64_t)B * pps * (f.fpad / 8), stream, c);
    if (t > 0)                              // dh_{t-1} += conv^T(dz_t, recurrent_kernel)
      if (run_dgrads(h, f, 1, B, T, t, T, t - 1, 1, gwritten, stream)) return 1;
  }
  run_colsum(h, f.dz_buf, B * T, grads + h->params[f.bias_param].offset, 0, f.F, f.fpad, stream);
  if (run_wgrad(h, f, f.dz_buf, T, grads, stream)) return 1;
  if (run_dgrads(h, f, 0, B * T, 1, 0, 1, 0, -1, gwritten, stream)) return 1;   // dx_t for all frames at once
  return 0;
}

static int train_loss_only(lu_handle_s* h, const float* labels, const float* cw, float* loss_out, uint16_t* gout,
                           void* stream) {
  const lu_config& c = h->cfg;
  ConvPlan& lc = h->convs[h->logits_conv];
  LU_REQUIRE(lc.cout == 3, "WeightedCELoss is defined for 3 classes (losses.py:20), network has %d", lc.cout);
  const int T = h->last_T, N = c.batch * T;
  double* acc = reinterpret_cast<double*>(h->ws + h->tr.off_loss_acc);
  LU_MEMSET(acc, 0, 16, stream);
  LuCeReduce r;
  r.raw = reinterpret_cast<const float*>(h->ws + lc.off_raw); r.labels = labels; r.acc = acc;
  r.npix = (int64_t)N * c.height * c.width; r.H = c.height; r.W = c.width; r.Hp = h->Hp; r.Wp = h->Wp;
  r.py0 = h->pad_y0; r.px0 = h->pad_x0; r.raw_cpad = lc.raw_cpad; r.chunk = 256; r.w0 = cw[0]; r.w1 = cw[1]; r.w2 = cw[2];
  pf(h, (r.npix + r.chunk - 1) / r.chunk, stream, r);
  // single-device semantics across ranks (losses.py:26 divides by the valid pixels of the WHOLE batch): with the rank-sum
  // callback set, the valid count is summed over the ranks and loss / gradients are scaled so that their mean over the
  // ranks is the single-device value
  float rank_scale = 1.f;
  if (h->bn_sync_fn) { h->bn_sync_fn(acc + 1, 1, h->bn_sync_user); rank_scale = (float)h->bn_sync_world; }
  LuCeGrad g;
  g.raw = r.raw; g.labels = labels; g.acc = acc; g.loss_out = loss_out; g.g = gout;
  g.H = c.height; g.W = c.width; g.Hp = h->Hp; g.Wp = h->Wp; g.py0 = h->pad_y0; g.px0 = h->pad_x0; g.raw_cpad = lc.raw_cpad;
  g.w0 = cw[0]; g.w1 = cw[1]; g.w2 = cw[2]; g.cpad = 0; g.planes = 1; g.rank_scale = rank_scale;
  if (gout != nullptr) {
    const ActBuf& gb = h->acts[h->g_logits_buf];
    g.cpad = gb.cpad; g.planes = gb.planes;
    pf(h, (int64_t)N * h->Hp * h->Wp, stream, g);
  } else {
    LuCeLossOnly lo; lo.acc = acc; lo.loss_out = loss_out; lo.rank_scale = rank_scale;
    pf(h, 1, stream, lo);
  }
  return 0;
}

// the contiguous range of the flat gradient buffer that holds the trainable tensors whose names start with `prefix`
static void notify_bucket(lu_handle_s* h, const char* prefix) {
  if (!h->bucket_fn) return;
  int64_t lo = -1, hi = -1;
  const size_t n = strlen(prefix);
  for (auto& p : h->params) {
    if (!p.trainable || p.name.compare(0, n, prefix) != 0) continue;
    if (lo < 0 || p.offset < lo) lo = p.offset;
    if (p.offset + p.count > hi) hi = p.offset + p.count;
  }
  if (lo >= 0) h->bucket_fn(lo, hi - lo, h->bucket_user);
}

static int train_loss_backward(lu_handle_s* h, const float* labels, const float* cw, float* loss_out, float* grads,
                               void* stream) {
  LU_REQUIRE(h && h->bound && h->dparams, "bind workspace and parameters first");
  LU_REQUIRE(labels && cw && loss_out, "null argument");
  LU_REQUIRE(h->last_T > 0, "call lu_forward first");
  if (grads == nullptr) return train_loss_only(h, labels, cw, loss_out, nullptr, stream);
  LU_REQUIRE(h->cfg.train, "handle was created without train=1");
  LU_REQUIRE(h->last_training, "lu_loss_backward needs a preceding lu_forward(training=1)");
  const int T = h->last_T;
  LU_MEMSET(grads, 0, (size_t)h->n_train * 4, stream);
  if (train_loss_only(h, labels, cw, loss_out, act_ptr(h, h->g_logits_buf), stream)) return 1;
  std::vector<char> gwritten(h->acts.size(), 0);
  for (int u = h->L - 1; u >= 0; --u) {
    const std::vector<int>& cl = h->conv_of_up[u];
    for (int j = (int)cl.size() - 1; j >= 0; --j)
      if (bwd_conv_layer(h, h->convs[cl[j]], T, grads, gwritten, stream)) return 1;
    if (h->ups[u].src_buf >= 0) {
      const ActBuf& s = h->acts[h->ups[u].src_buf];
      const int gs = h->gidx[h->ups[u].src_buf], gd = h->gidx[h->ups[u].dst_buf];
      LuUpsample2xBwd ub;
      ub.gup = act_ptr(h, gd); ub.gsrc = act_ptr(h, gs); ub.h = s.H; ub.w = s.W; ub.cpad = s.cpad; ub.planes = s.planes;
      ub.accumulate = gwritten[gs] ? 1 : 0;
      pf(h, (int64_t)h->cfg.batch * T * s.H * s.W * (s.cpad / 8), stream, ub);
      gwritten[gs] = 1;
    }
    { char pre[64]; snprintf(pre, sizeof pre, "UpLayers/%d/", u); notify_bucket(h, pre); }
  }
  for (int l = h->L - 1; l >= 0; --l) {
    const std::vector<int>& cl = h->conv_of_level[l];
    for (int j = (int)cl.size() - 1; j >= 0; --j)
      if (bwd_conv_layer(h, h->convs[cl[j]], T, grads, gwritten, stream)) return 1;
    const std::vector<int>& ll = h->lstm_of_level[l];
    for (int j = (int)ll.size() - 1; j >= 0; --j)
      if (bwd_lstm_layer(h, h->convs[ll[j]], T, grads, gwritten, stream)) return 1;
    { char pre[64]; snprintf(pre, sizeof pre, "DownLayers/%d/", l); notify_bucket(h, pre); }
  }
#ifndef LU_HOST_EMU
  cudaError_t e = cudaGetLastError();
  LU_REQUIRE(e == cudaSuccess, "backward: %s", cudaGetErrorString(e));
#endif
  return 0;
}

static int train_adam(lu_handle_s* h, const float* g, float* m, float* v, float lr, float b1, float b2, float eps,
                      int64_t step, void* stream) {
  LU_REQUIRE(h && h->dparams && g && m && v, "null argument");
  LU_REQUIRE(step >= 1, "Adam step is 1-based");
  LuAdam a;
  a.p = h->dparams; a.g = g; a.m = m; a.v = v; a.b1 = b1; a.b2 = b2; a.eps = eps;
  a.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step)));
  pf(h, h->n_train, stream, a);
  h->packed = false;
  return 0;
}
