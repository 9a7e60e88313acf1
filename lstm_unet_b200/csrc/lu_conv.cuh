// Table-driven implicit-GEMM convolution: shared epilogues, the scalar mirror kernel and the tcgen05 kernel.
//
// GEMM view: D[128 output pixels (16x8 tile), BN packed output columns] += A[pixels, 64-channel K block] * B.
// A comes straight from NHWC bf16 activation buffers (no im2col/im2row buffer is ever materialised): a K block
// is "source view, channel chunk, filter tap", and the tap only shifts which rows of a staged [rows x pitch]
// pixel window the MMA reads.  B is the pre-packed K-major bf16 weight matrix [Npad][Ktot].
#pragma once
#include "lu_defs.h"
#ifndef LU_HOST_EMU
#include <cuda_bf16.h>
#endif

#define LU_PT_STAGES 40
#define LU_PT_TAPS 160

struct LuConvParams {
  LuSrcView src[LU_MAX_SRC];
  const LuAStage* astages;
  const uint16_t* taps;
  const uint16_t* wpacked;        // [Npad][ktot]
  int32_t n_astages, ktot;
  int32_t tiles_x, tiles_y, frames, n_tiles_n, BN;
  LuEpi epi;
};

// ---- 16-wide stores -----------------------------------------------------------------------------------------
LU_HDI void lu_store16_f32(float* dst, const float* v) {
#ifdef __CUDA_ARCH__
  float4* d = reinterpret_cast<float4*>(dst);
  d[0] = make_float4(v[0], v[1], v[2], v[3]);
  d[1] = make_float4(v[4], v[5], v[6], v[7]);
  d[2] = make_float4(v[8], v[9], v[10], v[11]);
  d[3] = make_float4(v[12], v[13], v[14], v[15]);
#else
  for (int j = 0; j < 16; ++j) dst[j] = v[j];
#endif
}
LU_HDI void lu_load16_f32(const float* src, float* v) {
#ifdef __CUDA_ARCH__
  const float4* s = reinterpret_cast<const float4*>(src);
  float4 a = s[0], b = s[1], c = s[2], d = s[3];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
#else
  for (int j = 0; j < 16; ++j) v[j] = src[j];
#endif
}
LU_HDI void lu_store16_bf16(uint16_t* dst, const uint16_t* h) {
#ifdef __CUDA_ARCH__
  uint4 a, b;
  a.x = h[0] | ((uint32_t)h[1] << 16);  a.y = h[2] | ((uint32_t)h[3] << 16);
  a.z = h[4] | ((uint32_t)h[5] << 16);  a.w = h[6] | ((uint32_t)h[7] << 16);
  b.x = h[8] | ((uint32_t)h[9] << 16);  b.y = h[10] | ((uint32_t)h[11] << 16);
  b.z = h[12] | ((uint32_t)h[13] << 16); b.w = h[14] | ((uint32_t)h[15] << 16);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
#else
  for (int j = 0; j < 16; ++j) dst[j] = h[j];
#endif
}

// ---- bf16 packing of 16 values (hi plane, optional lo plane) ---------------------------------------------------
LU_HDI void lu_store16_split(uint16_t* dst, int lo_off, bool want_lo, const float* a, int fmt = 0) {
#ifdef __CUDA_ARCH__
  uint32_t h[8], l[8];
  if (fmt) {                                  // fp16 operands: one plane
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2 hv = __floats2half2_rn(a[2 * j], a[2 * j + 1]);
      h[j] = *reinterpret_cast<const uint32_t*>(&hv);
    }
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(h[4], h[5], h[6], h[7]);
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 hv = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]);
    h[j] = *reinterpret_cast<const uint32_t*>(&hv);
    if (want_lo) {
      const float2 hf = __bfloat1622float2(hv);
      const __nv_bfloat162 lv = __floats2bfloat162_rn(a[2 * j] - hf.x, a[2 * j + 1] - hf.y);
      l[j] = *reinterpret_cast<const uint32_t*>(&lv);
    }
  }
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(h[0], h[1], h[2], h[3]);
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(h[4], h[5], h[6], h[7]);
  if (want_lo) {
    reinterpret_cast<uint4*>(dst + lo_off)[0] = make_uint4(l[0], l[1], l[2], l[3]);
    reinterpret_cast<uint4*>(dst + lo_off)[1] = make_uint4(l[4], l[5], l[6], l[7]);
  }
#else
  for (int j = 0; j < 16; ++j) {
    if (fmt) { dst[j] = lu_f2half(a[j]); continue; }
    uint16_t hi, lo; lu_split(a[j], hi, lo);
    dst[j] = hi;
    if (want_lo) dst[lo_off + j] = lo;
  }
#endif
}

// ---- epilogues (one 16-column chunk of one output pixel) -------------------------------------------------
// bias/scale/shift point at the 16 per-column constants of this chunk (shared memory in the tcgen05 kernel).
// conv: v = acc + bias -> optional fp32 raw store; optional folded-BN + LeakyReLU -> bf16 (hi[,lo]) store.
LU_HDI void lu_epi_conv_chunk(const LuEpi& e, int64_t pix, int n, float* v, const float* bias, const float* scale,
                              const float* shift) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += bias[j];
  if (e.out_raw != nullptr && n < e.raw_cpad) lu_store16_f32(e.out_raw + pix * e.raw_cpad + n, v);
  if (e.out_act != nullptr && n < e.out_cpad) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = v[j] * scale[j] + shift[j];
      v[j] = a > 0.f ? a : e.alpha * a;
    }
    lu_store16_split(e.out_act + pix * (int64_t)(e.out_cpad * e.out_planes) + n, e.out_cpad, e.out_planes == 2, v, e.fmt);
  }
}

// gradient convolution: store (or accumulate into) a bf16 hi[/lo] gradient buffer
LU_HDI void lu_epi_grad_chunk(const LuEpi& e, int64_t pix, int n, float* v) {
  if (n >= e.out_cpad) return;
  uint16_t* o = e.out_act + pix * (int64_t)(e.out_cpad * e.out_planes) + n;
  if (e.accumulate) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] += lu_bf2f(o[j]);
      if (e.out_planes == 2) v[j] += lu_bf2f(o[e.out_cpad + j]);
    }
  }
  lu_store16_split(o, e.out_cpad, e.out_planes == 2, v);
}

// ConvLSTM cell (keras ConvLSTM2D defaults): z* are the pre-activations of gates i,f,c,o for 16
// channels starting at ch0; b* the matching bias slices; pix_state indexes the per-sample states, pix_out the h
// sequence buffer.
LU_HDI void lu_epi_lstm_chunk(const LuEpi& e, int64_t pix_state, int64_t pix_out, int ch0, const float* zi,
                              const float* zf, const float* zg, const float* zo, const float* bi, const float* bf,
                              const float* bg, const float* bo) {
  float* cp = e.c_state + pix_state * e.f_pad + ch0;
  float c[16], gi[16], gf[16], gg[16], go[16], hh[16];
  lu_load16_f32(cp, c);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float ai = zi[j] + bi[j], af = zf[j] + bf[j], ag = zg[j] + bg[j], ao = zo[j] + bo[j];
    if (e.gate_kind == 0) { gi[j] = lu_hard_sigmoid(ai); gf[j] = lu_hard_sigmoid(af); go[j] = lu_hard_sigmoid(ao); }
    else { gi[j] = lu_sigmoid(ai); gf[j] = lu_sigmoid(af); go[j] = lu_sigmoid(ao); }
    gg[j] = tanhf(ag);
    c[j] = gf[j] * c[j] + gi[j] * gg[j];
    hh[j] = go[j] * tanhf(c[j]);
  }
  lu_store16_f32(cp, c);
  const int64_t ctot = (int64_t)e.f_pad * e.out_planes;
  const bool lo = e.out_planes == 2;
  lu_store16_split(e.out_act + pix_out * ctot + ch0, e.f_pad, lo, hh, e.fmt);
  if (e.h_state_out != nullptr) lu_store16_split(e.h_state_out + pix_state * ctot + ch0, e.f_pad, lo, hh, e.fmt);
  if (e.save_c != nullptr) lu_store16_f32(e.save_c + pix_out * e.f_pad + ch0, c);
  if (e.save_gates != nullptr) {
    uint16_t* g = e.save_gates + pix_out * (int64_t)(4 * e.f_pad * e.out_planes) + ch0;
    lu_store16_split(g, 4 * e.f_pad, lo, gi);
    lu_store16_split(g + e.f_pad, 4 * e.f_pad, lo, gf);
    lu_store16_split(g + 2 * e.f_pad, 4 * e.f_pad, lo, gg);
    lu_store16_split(g + 3 * e.f_pad, 4 * e.f_pad, lo, go);
  }
}

// ---- scalar mirror: executes exactly the same tables / packed weights with plain loads and FMAs -------------
LU_HDI void lu_mirror_acc16(const LuConvParams& p, int frame, int y0, int x0, int m, int ncol, float* acc) {
  const int ty = m / LU_TILE_W, tx = m % LU_TILE_W;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  int kb = 0;
  for (int s = 0; s < p.n_astages; ++s) {
    const LuAStage st = p.astages[s];
    const LuSrcView& v = p.src[st.src];
    const int64_t n = (int64_t)frame * v.frame_mul + v.frame_add;
    for (int t = 0; t < st.ntaps; ++t, ++kb) {
      const int off = p.taps[st.tap_begin + t];
      const int yy = y0 + st.dy + off / v.pitch + ty;
      const int xx = x0 + st.dx + off % v.pitch + tx;
      if (yy < 0 || yy >= v.dimH || xx < 0 || xx >= v.dimW) continue;      // TMA zero-fills out-of-bounds
      const uint16_t* a = v.ptr + n * v.sn + (int64_t)yy * v.sh + (int64_t)st.plane * v.sp + (int64_t)xx * v.sw + st.c;
      const uint16_t* w = p.wpacked + (int64_t)ncol * p.ktot + (int64_t)kb * LU_KBLK;
      int cmax = v.dimC - st.c;
      if (cmax > LU_KBLK) cmax = LU_KBLK;
      for (int kk = 0; kk < cmax; ++kk) {
        const float av = lu_h162f(a[kk], p.epi.fmt);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += av * lu_h162f(w[(int64_t)j * p.ktot + kk], p.epi.fmt);
      }
    }
  }
}

// item = ((m_tile * n_tiles_n + n_tile) * 128 + m) * chunks + chunk
LU_HDI void lu_conv_mirror_item(const LuConvParams& p, int64_t item) {
  const LuEpi& e = p.epi;
  const int chunks = (e.kind == LU_EPI_LSTM) ? e.ch_tile / 16 : p.BN / 16;
  const int chunk = (int)(item % chunks); item /= chunks;
  const int m = (int)(item % 128); item /= 128;
  const int nt = (int)(item % p.n_tiles_n); item /= p.n_tiles_n;
  const int tx = (int)(item % p.tiles_x); item /= p.tiles_x;
  const int ty = (int)(item % p.tiles_y); item /= p.tiles_y;
  const int frame = (int)item;
  const int y0 = ty * LU_TILE_H, x0 = tx * LU_TILE_W;
  const int y = y0 + m / LU_TILE_W, x = x0 + m % LU_TILE_W;
  if (y >= e.H || x >= e.W) return;
  const int n0 = nt * p.BN;
  const int64_t fout = (int64_t)frame * e.out_frame_mul + e.out_frame_add;
  const int yo = y * e.oy_mul + e.oy_add, xo = x * e.ox_mul + e.ox_add;
  if (yo >= e.OH || xo >= e.OW) return;
  const int64_t pix_out = (fout * e.OH + yo) * e.OW + xo;
  if (e.kind == LU_EPI_CONV) {
    float v[16];
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + chunk * 16, v);
    const int n = n0 + chunk * 16;
    if (e.bn_sums != nullptr && n < e.raw_cpad)
      for (int j = 0; j < 16; ++j) {
        lu_atomic_add(&e.bn_sums[n + j], (double)v[j]);
        lu_atomic_add(&e.bn_sums[e.raw_cpad + n + j], (double)v[j] * (double)v[j]);
      }
    lu_epi_conv_chunk(e, pix_out, n, v, e.bias + n, e.scale ? e.scale + n : nullptr, e.shift ? e.shift + n : nullptr);
  } else if (e.kind == LU_EPI_GRAD) {
    float v[16];
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + chunk * 16, v);
    lu_epi_grad_chunk(e, pix_out, n0 + chunk * 16, v);
  } else {
    const int CH = e.ch_tile, jc = chunk * 16;
    float zi[16], zf[16], zg[16], zo[16];
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + jc, zi);
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + CH + jc, zf);
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + 2 * CH + jc, zg);
    lu_mirror_acc16(p, frame, y0, x0, m, n0 + 3 * CH + jc, zo);
    const int64_t pix_state = ((int64_t)frame * e.H + y) * e.W + x;
    const float* b = e.bias + n0 + jc;
    lu_epi_lstm_chunk(e, pix_state, pix_out, nt * CH + jc, zi, zf, zg, zo, b, b + CH, b + 2 * CH, b + 3 * CH);
  }
}

#ifndef LU_HOST_EMU
// =============================================================================================================
// tcgen05 kernel (sm_100a): TMA-staged operands, single-thread UMMA issue, accumulators in TMEM.
// =============================================================================================================
#include <cuda.h>

struct LuTcParams {
  CUtensorMap tmA[LU_MAX_SRC];
  CUtensorMap tmB;
  CUtensorMap tmBh;              // half-height weight box (BN/2 rows): each CTA of a 2-CTA cluster multicasts one half
  LuConvParams cp;
  int32_t n_a_stages, n_b_stages, a_stage_bytes, b_stage_bytes;
  int32_t b_group;               // K blocks per weight stage (one mbarrier wait / commit per group)
  int32_t b_resident;            // 1: n_b_stages holds the whole weight panel; loaded for the CTA's first tile only
  int32_t two_issuers;           // resident-weight convolutions: a second issuing thread (warp 2) takes the odd tiles / accumulator 1
  int32_t cst_per_tile;          // experiment switch (LU_CST_PER_TILE=1): re-stage the per-column constants for every tile
  int32_t acc_split;             // experiment switch (LU_ACC_SPLIT = 2 / 4, default 1): R partial accumulators per tile for narrow N
                                 // tiles -- the K = 16 MMAs of a K block go round-robin to R TMEM accumulators of BN columns each, summed
                                 // by the epilogue, so that consecutive MMAs do not wait on the same accumulator.  Parity-green; measured
                                 // no effect on B200 (the non-ConvLSTM part of the C2 step 12.4 -> 12.5 ms): dependent-MMA latency is not
                                 // what bounds the 32 / 64-channel convs
  uint32_t idesc;
  int32_t total_tiles;           // work items: tiles (cluster 1) or pairs of M tiles sharing an N tile (cluster 2)
  int32_t num_mt;                // number of M tiles (frames * tiles per frame)
  // Copies of the staging tables in kernel-parameter (constant) space: the issuing warps index them with
  // warp-uniform loop counters, so descriptors and TMA coordinates stay in uniform registers (no per-instruction
  // divergence "waterfall" around tcgen05.mma / TMA).  tap lists are de-duplicated; tables_in_params == 0 falls
  // back to the global-memory tables (direct staging, exotic shapes).
  int32_t tables_in_params;
  LuAStage st_tab[LU_PT_STAGES];
  uint16_t tap_tab[LU_PT_TAPS];
};

namespace lutc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                               int c3, int c4, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// ---- cta_group::2 forms (cluster mode 3): operands of one M = 256 MMA are spread over the CTA pair ---------------
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA loads whose completion bytes are counted on an mbarrier that may live in the PEER CTA (the pair leader's)
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// the leader's commit: one arrive on the barrier at this offset in every CTA of the mask, once the pair's MMAs retire
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle operand descriptor (cute::UMMA::SmemDescriptor layout, sm_100 version bit set):
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (ignored for swizzled K-major) |
//   [32,46) stride byte offset >> 4 = distance between 8-row groups | [46,48) version = 1 | [49,52) base offset = 0 |
//   [61,64) layout = 2 (SWIZZLE_128B).
// The 128-byte swizzle is a function of the shared-memory ADDRESS bits (measured on B200: a start address that is a
// multiple of 128 but not of 1024 bytes reads the rows TMA wrote there when the base-offset field is left 0; setting it
// to (addr >> 7) & 7 gives wrong results), which is what lets a filter tap be a mere row offset into the staged halo
// window, and lets the 8-row group stride be pitch*128 bytes for any pitch.
// high / low words of the K-major SWIZZLE_128B descriptor (layout above): hi = SBO | version | layout,
// lo = (address >> 4) | LBO.  Taps and K sub-steps only add to the low word.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// wait for outstanding tcgen05.ld; the "+f" operands pin the loaded registers behind the wait
__device__ __forceinline__ void tmem_wait16(float* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                 "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}

constexpr int kThreads = 384;      // warp 0: A producer, 1: MMA issuer, 2: TMEM allocator, 3: B producer, 4-11: epilogue
constexpr int kEpiThreads = 256;   // two warps per TMEM lane quarter: one SMSP-resident warp cannot hide the epilogue latency
constexpr int kConstFloats = 3 * 256;   // per accumulator stage: bias | scale | shift of the tile's BN columns
constexpr int kTmemCols = 512;

}  // namespace lutc

// CL = cluster mode.  1: no cluster.  2 and 3: the two CTAs of a 2-CTA cluster work on two M tiles of the SAME N tile.
//   CL == 2: each CTA multicasts one half of every weight K block into both CTAs' shared memory (half the L2->SM weight
//            traffic); the MMA and the accumulators stay per-CTA (cta_group::1).
//   CL == 3: the pair runs ONE M = 256 MMA per K step (tcgen05.mma.cta_group::2, issued by the even CTA): every CTA
//            stages its own 128-pixel window and only ITS half of the weight K block (rows [rank*BN/2, +BN/2), no
//            multicast), the tensor core reads the two halves from the two shared memories.  Per SM and K step the
//            operand read drops from (128 + BN) to (128 + BN/2) rows of 32 bytes.  All "full" barriers that the issuer
//            waits on live in the leader CTA: both producers' TMA bytes complete there, the peer's epilogue threads
//            arrive there; every "empty" / "accumulator ready" barrier is released in both CTAs by the leader's
//            multicast commit.  (Compiled, not yet run on hardware: LU_PAIR=1 selects it.)
template <int EPI, bool PTAB, int CL>
__global__ void __launch_bounds__(lutc::kThreads, 1) lu_conv_tc_kernel(const __grid_constant__ LuTcParams P) {
  using namespace lutc;
  constexpr bool PAIR = CL == 3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;   // SWIZZLE_128B needs 1024-byte aligned stages
  uint8_t* smem = smem_raw + pad;
  const int nA = P.n_a_stages, nB = P.n_b_stages;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + (uint32_t)nA * P.a_stage_bytes;
  const uint32_t bars = sB + (uint32_t)nB * P.b_stage_bytes;       // 8-byte mbarriers
  const uint32_t full_a = bars, empty_a = full_a + 8u * nA, full_b = empty_a + 8u * nA, empty_b = full_b + 8u * nB;
  const uint32_t tmem_full = empty_b + 8u * nB, tmem_empty = tmem_full + 16u, tmem_slot = tmem_empty + 16u;
  uint8_t* after_bars = smem + (size_t)nA * P.a_stage_bytes + (size_t)nB * P.b_stage_bytes + 16u * (nA + nB) + 32u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(after_bars);
  float* s_const = reinterpret_cast<float*>(after_bars + 16);        // [2][kConstFloats]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LuConvParams& cp = P.cp;
  const int BN = cp.BN;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < LU_MAX_SRC; ++i)
      if (cp.src[i].rows > 0) prefetch_tmap(&P.tmA[i]);
    prefetch_tmap(&P.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nA; ++i) { mbar_init(full_a + 8u * i, 1); mbar_init(empty_a + 8u * i, 1); }
    for (int i = 0; i < nB; ++i) { mbar_init(full_b + 8u * i, 1); mbar_init(empty_b + 8u * i, CL == 2 ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + 8u * i, 1); mbar_init(tmem_empty + 8u * i, PAIR ? 2 * kEpiThreads : kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {            // the same warp of BOTH CTAs: the columns are allocated in the two tensor memories together
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CL == 1) __syncthreads(); else cluster_sync_all();       // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_frame = cp.tiles_x * cp.tiles_y;
  // work decomposition: item -> (N tile, M tile).  cluster 1: item = tile, N fastest.  cluster 2: item = pair of
  // consecutive M tiles of one N tile; an odd tail pair re-does the last M tile in the second CTA with stores masked.
  const int crank = (CL >= 2) ? (int)cluster_ctarank() : 0;
  const int item0 = (CL >= 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = (CL >= 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // pair mode: the leader's copy of a barrier, as seen from this CTA
  auto leader = [&](uint32_t bar) { return PAIR ? map_to_rank(bar, 0u) : bar; };
  auto decode = [&](int item, int& nt, int& mt, bool& dummy) {
    if (CL >= 2) {
      nt = item % cp.n_tiles_n; mt = 2 * (item / cp.n_tiles_n) + crank;      // N fastest: concurrent clusters share the activation tiles (L2)
      dummy = mt >= P.num_mt;
      if (dummy) mt = P.num_mt - 1;
    } else { nt = item % cp.n_tiles_n; mt = item / cp.n_tiles_n; dummy = false; }
  };

  // The three issuing roles run with warp-uniform control flow (all 32 lanes walk the loops and wait on the
  // barriers); one elected lane issues the asynchronous instruction.
  constexpr bool ptab = PTAB;
  if (warp == 0) {
    // ------------------------------------------------------------------ A producer (activation windows)
    int sa = 0; uint32_t ph = 0;
    for (int tile = item0; tile < P.total_tiles; tile += item_step) {
      int nt, mt; bool dummy;
      decode(tile, nt, mt, dummy);
      const int frame = mt / tiles_per_frame, rem = mt % tiles_per_frame;
      const int y0 = (rem / cp.tiles_x) * LU_TILE_H, x0 = (rem % cp.tiles_x) * LU_TILE_W;
      for (int s = 0; s < cp.n_astages; ++s) {
        const LuAStage st = ptab ? P.st_tab[s] : cp.astages[s];
        const LuSrcView& v = cp.src[st.src];
        mbar_wait(empty_a + 8u * sa, ph ^ 1u);
        if (elect_one()) {
          if (PAIR) {          // both windows of the pair are counted on the leader's barrier
            if (crank == 0) mbar_expect_tx(full_a + 8u * sa, 2u * (uint32_t)(v.rows * v.pitch) * 128u);
            tma_load_5d_pair(sA + (uint32_t)sa * P.a_stage_bytes, &P.tmA[st.src], leader(full_a + 8u * sa), st.c, x0 + st.dx,
                             st.plane, y0 + st.dy, frame * v.frame_mul + v.frame_add);
          } else {
            mbar_expect_tx(full_a + 8u * sa, (uint32_t)(v.rows * v.pitch) * 128u);
            tma_load_5d(sA + (uint32_t)sa * P.a_stage_bytes, &P.tmA[st.src], full_a + 8u * sa, st.c, x0 + st.dx, st.plane,
                        y0 + st.dy, frame * v.frame_mul + v.frame_add);
          }
        }
        __syncwarp();
        if (++sa == nA) { sa = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ B producer (packed weight K blocks)
    int sb = 0; uint32_t ph = 0;
    const int nkb = cp.ktot / LU_KBLK, G = P.b_group;
    for (int tile = item0; tile < P.total_tiles; tile += item_step) {
      int nt, mt; bool dummy;
      decode(tile, nt, mt, dummy);
      if (P.b_resident && tile != item0) continue;               // the panel of the first tile stays in shared memory
      for (int kb = 0; kb < nkb; kb += G) {
        const int g = (nkb - kb) < G ? (nkb - kb) : G;
        mbar_wait(empty_b + 8u * sb, ph ^ 1u);               // cluster 2: released by BOTH CTAs' MMA issuers
        if (elect_one()) {
          if (!PAIR || crank == 0) mbar_expect_tx(full_b + 8u * sb, (uint32_t)(g * BN) * 128u);   // pair: both halves, on the leader
          for (int j = 0; j < g; ++j) {
            const uint32_t dst = sB + (uint32_t)sb * P.b_stage_bytes + (uint32_t)(j * (PAIR ? (BN >> 1) : BN)) * 128u;
            if (PAIR)            // this CTA's half of the K block only, at the same offset in both shared memories
              tma_load_2d_pair(dst, &P.tmBh, leader(full_b + 8u * sb), (kb + j) * LU_KBLK, nt * BN + crank * (BN >> 1));
            else if (CL == 2)
              tma_load_2d_mc(dst + (uint32_t)(crank * (BN >> 1)) * 128u, &P.tmBh, full_b + 8u * sb, (kb + j) * LU_KBLK,
                             nt * BN + crank * (BN >> 1), (uint16_t)3);
            else
              tma_load_2d(dst, &P.tmB, full_b + 8u * sb, (kb + j) * LU_KBLK, nt * BN);
          }
        }
        __syncwarp();
        if (++sb == nB) { sb = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 || (warp == 2 && !PAIR && P.two_issuers != 0)) {
    // ------------------------------------------------------------------ MMA issuer
    // One lane is elected ONCE and runs the whole loop nest (waits included); inside, everything it touches is
    // warp-uniform by construction, so ptxas keeps descriptors in uniform registers and the per-tap cost is a
    // handful of instructions.
    // Narrow N tiles (resident weights, N <= 64) are bound by THIS loop -- ~40 dependent scalar instructions around four
    // 16-clock MMAs per tap (source-level ncu sampling) -- so a second issuing thread (warp 2, idle after the TMEM
    // allocation) takes the CTA's odd tiles: issuer w owns accumulator stage w and every second group of n_astages
    // activation stages; the barriers are per stage, a commit covers the issuing thread's own MMAs.
    const int nw = (!PAIR && P.two_issuers != 0) ? 2 : 1;
    const int w = (warp == 2) ? 1 : 0;
    if ((!PAIR || crank == 0) && elect_one()) {                 // pair mode: the even CTA issues for both
      int sa = 0, sb = 0, acc = w; uint32_t pha = 0, phb = 0, phacc = 0;
      auto skip_a = [&](int n) { for (int i = 0; i < n; ++i) if (++sa == nA) { sa = 0; pha ^= 1u; } };
      skip_a(w * cp.n_astages);                                 // the stages of tile 0 belong to issuer 0
      const int G = P.b_group;
      const uint32_t b_hi = desc_hi(1024u);
      const uint32_t bn_bytes16 = (uint32_t)(PAIR ? (BN >> 1) : BN) * 8u;   // one K block of weights in this CTA, in 16-byte units
      // pair mode: M = 256 (bits [24,29) hold M >> 4)
      const uint32_t idesc = PAIR ? ((P.idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24)) : P.idesc;
      const bool resident = P.b_resident != 0;
      bool first_tile = true;
      for (int tile = item0 + w * item_step; tile < P.total_tiles; tile += nw * item_step) {
        mbar_wait(tmem_empty + 8u * acc, phacc ^ 1u);
        tc_fence_after();
        const int R = (EPI == LU_EPI_LSTM) ? 1 : P.acc_split;
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN * R);
        uint32_t accum = 0;
        int gi = 0;                                             // position inside the current weight group
        uint32_t b_lo = 0;
        for (int s = 0; s < cp.n_astages; ++s) {
          const LuAStage st = PTAB ? P.st_tab[s] : cp.astages[s];
          const uint32_t a_hi = desc_hi((uint32_t)cp.src[st.src].pitch * 128u);
          mbar_wait(full_a + 8u * sa, pha);
          tc_fence_after();
          const uint32_t a_lo = desc_lo(sA + (uint32_t)sa * P.a_stage_bytes);
          const int ntaps = st.ntaps;
          const uint32_t tb = st.tap_begin;
          for (int t = 0; t < ntaps; ++t) {
            // tap offset in rows of 128 bytes -> 16-byte units
            const uint32_t off8 = (PTAB ? (uint32_t)P.tap_tab[tb + t] : (uint32_t)cp.taps[tb + t]) * 8u;
            if (gi == 0) {
              if (!resident || first_tile) {
                mbar_wait(full_b + 8u * sb, phb);
                tc_fence_after();
              }
              b_lo = desc_lo(sB + (uint32_t)sb * P.b_stage_bytes);
            }
#pragma unroll
            for (int k = 0; k < LU_KBLK / 16; ++k) {
              // partial accumulator k mod R; the first R MMAs of a tile overwrite, everything after accumulates
              const uint32_t d_k = d_tmem + (uint32_t)((k & (R - 1)) * BN);
              const uint32_t acc_k = accum | (uint32_t)(k >= R);
              if (PAIR) mma_bf16_pair(d_k, a_lo + off8 + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc, acc_k);
              else mma_bf16(d_k, a_lo + off8 + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc, acc_k);
            }
            accum = 1;
            b_lo += bn_bytes16;
            if (++gi == G) {
              if (PAIR) tc_commit_pair(empty_b + 8u * sb, (uint16_t)3);    // both CTAs' halves of the stage
              else if (CL == 2) tc_commit_mc(empty_b + 8u * sb, (uint16_t)3);   // weight stage is shared by the cluster
              else if (!resident) tc_commit(empty_b + 8u * sb);  // frees the weight stage once its MMAs retire
              gi = 0;
              if (++sb == nB) { sb = 0; phb ^= 1u; }
            }
          }
          if (PAIR) tc_commit_pair(empty_a + 8u * sa, (uint16_t)3);   // frees both CTAs' windows
          else tc_commit(empty_a + 8u * sa);                    // frees the activation window
          if (++sa == nA) { sa = 0; pha ^= 1u; }
        }
        if (gi != 0) {                                          // partial last weight group of the tile
          if (PAIR) tc_commit_pair(empty_b + 8u * sb, (uint16_t)3);
          else if (CL == 2) tc_commit_mc(empty_b + 8u * sb, (uint16_t)3);
          else if (!resident) tc_commit(empty_b + 8u * sb);
          if (++sb == nB) { sb = 0; phb ^= 1u; }
        }
        first_tile = false;
        if (PAIR) tc_commit_pair(tmem_full + 8u * acc, (uint16_t)3);   // both halves of the M = 256 accumulator
        else tc_commit(tmem_full + 8u * acc);                   // accumulator complete -> epilogue
        if (nw == 2) { phacc ^= 1u; skip_a(cp.n_astages); }     // own accumulator stage again; the other issuer's windows skipped
        else if (++acc == 2) { acc = 0; phacc ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> HBM
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;               // the two warps of a quarter take alternate 16-column chunks
    const int m = q * 32 + lane;
    const int ep_tid = threadIdx.x - 128;
    const LuEpi& e = cp.epi;
    // training-mode BatchNorm: per-channel sum / sum of squares of the accumulators (= output - bias) of this CTA's tiles,
    // kept in the scale / shift slots of the constant staging area (unused in this mode): stage 0's for the sums, stage 1's
    // for the squares, indexed by the packed column (npad <= 512)
    const bool bn_stats = (EPI == LU_EPI_CONV) && e.bn_sums != nullptr;
    float* s_sum = s_const + 256;
    float* s_sq = s_const + kConstFloats + 256;
    if (bn_stats)
      for (int i = ep_tid; i < 512; i += kEpiThreads) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
    // Per-column constants (bias / folded-BN scale / shift) go through shared memory.  A convolution with ONE N tile has the
    // same constants for every tile: staged once for both accumulator stages instead of a global-load round trip and a
    // 256-thread barrier per 128-pixel tile (those showed as barrier + long-scoreboard stalls in the ncu capture of the
    // 32-channel decoder convs).  Measured same-box A/B (LU_CST_PER_TILE): forward convs 10.0 -> 9.7 ms per train step, the
    // inference step unchanged within noise -- it was not what bounds those launches.
    const bool cst_per_tile = (EPI != LU_EPI_GRAD) && (cp.n_tiles_n > 1 || bn_stats || P.cst_per_tile);
    if (EPI != LU_EPI_GRAD && !cst_per_tile) {
      for (int i = ep_tid; i < 3 * BN; i += kEpiThreads) {
        const int which = i / BN, j = i - which * BN;
        const float* src = which == 0 ? e.bias : (which == 1 ? e.scale : e.shift);
        const float val = (src != nullptr) ? src[j] : 0.f;
        s_const[which * 256 + j] = val;
        s_const[kConstFloats + which * 256 + j] = val;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    int acc = 0; uint32_t phacc = 0;
    for (int tile = item0; tile < P.total_tiles; tile += item_step) {
      int nt, mt; bool dummy;
      decode(tile, nt, mt, dummy);
      const int frame = mt / tiles_per_frame, rem = mt % tiles_per_frame;
      const int y = (rem / cp.tiles_x) * LU_TILE_H + m / LU_TILE_W, x = (rem % cp.tiles_x) * LU_TILE_W + m % LU_TILE_W;
      const int yo = y * e.oy_mul + e.oy_add, xo = x * e.ox_mul + e.ox_add;
      const bool valid = !dummy && (y < e.H) && (x < e.W) && (yo < e.OH) && (xo < e.OW);
      const int n0 = nt * BN;
      const int64_t fout = (int64_t)frame * e.out_frame_mul + e.out_frame_add;
      const int64_t pix_out = (fout * e.OH + yo) * e.OW + xo;
      // per-column constants of this tile -> shared memory (double-buffered with the accumulator stage)
      float* cst = s_const + acc * kConstFloats;
      if (cst_per_tile) {
        for (int i = ep_tid; i < (bn_stats ? 1 : 3) * BN; i += kEpiThreads) {
          const int which = i / BN, j = i - which * BN;
          const float* src = which == 0 ? e.bias : (which == 1 ? e.scale : e.shift);
          cst[which * 256 + j] = (src != nullptr) ? src[n0 + j] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      mbar_wait(tmem_full + 8u * acc, phacc);
      tc_fence_after();
      const int R = (EPI == LU_EPI_LSTM) ? 1 : P.acc_split;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BN * R) + ((uint32_t)(q * 32) << 16);
      if (EPI == LU_EPI_CONV) {
        for (int col = half * 16; col < BN; col += 32) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)col, v);
          tmem_wait16(v);
          for (int r = 1; r < R; ++r) {                         // partial accumulators of a split tile
            float w[16];
            tmem_ld16(taddr + (uint32_t)(r * BN + col), w);
            tmem_wait16(w);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += w[j];
          }
          if (bn_stats) {
            // column sums over the warp's 32 pixel rows by recursive halving: 31 shuffles for 16 sums + 16 squares, lane L
            // ends up with the total of value L (L < 16: sum of column col + L; else: squares of column col + L - 16)
            float a[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float w = valid ? v[j] : 0.f; a[j] = w; a[16 + j] = w * w; }
#pragma unroll
            for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n; ++i) {
                const float send = up ? a[i] : a[i + n], keep = up ? a[i + n] : a[i];
                a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            atomicAdd((lane < 16 ? s_sum : s_sq) + n0 + col + (lane & 15), a[0]);
          }
          if (valid) lu_epi_conv_chunk(e, pix_out, n0 + col, v, cst + col, cst + 256 + col, cst + 512 + col);
        }
      } else if (EPI == LU_EPI_GRAD) {
        for (int col = half * 16; col < BN; col += 32) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)col, v);
          tmem_wait16(v);
          for (int r = 1; r < R; ++r) {
            float w[16];
            tmem_ld16(taddr + (uint32_t)(r * BN + col), w);
            tmem_wait16(w);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += w[j];
          }
          if (valid) lu_epi_grad_chunk(e, pix_out, n0 + col, v);
        }
      } else {
        const int CH = e.ch_tile;
        const int64_t pix_state = ((int64_t)frame * e.H + y) * e.W + x;
        for (int jc = half * 16; jc < CH; jc += 32) {
          float zi[16], zf[16], zg[16], zo[16];
          tmem_ld16(taddr + (uint32_t)jc, zi);
          tmem_ld16(taddr + (uint32_t)(CH + jc), zf);
          tmem_ld16(taddr + (uint32_t)(2 * CH + jc), zg);
          tmem_ld16(taddr + (uint32_t)(3 * CH + jc), zo);
          tmem_wait16(zi); tmem_wait16(zf); tmem_wait16(zg); tmem_wait16(zo);
          if (valid)
            lu_epi_lstm_chunk(e, pix_state, pix_out, nt * CH + jc, zi, zf, zg, zo, cst + jc, cst + CH + jc,
                              cst + 2 * CH + jc, cst + 3 * CH + jc);
        }
      }
      tc_fence_before();
      if (PAIR && crank != 0) mbar_arrive_cluster(leader(tmem_empty + 8u * acc));   // the issuer waits for both epilogues
      else mbar_arrive(tmem_empty + 8u * acc);
      if (++acc == 2) { acc = 0; phacc ^= 1u; }
    }
    if (bn_stats) {                                  // this CTA's partial sums -> the layer's fp64 accumulators
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int n = ep_tid; n < e.raw_cpad && n < 512; n += kEpiThreads) {
        const float su = s_sum[n], sq = s_sq[n];
        if (su != 0.f || sq != 0.f) { atomicAdd(&e.bn_sums[n], (double)su); atomicAdd(&e.bn_sums[e.raw_cpad + n], (double)sq); }
      }
    }
  }

  tc_fence_before();
  if (CL == 1) __syncthreads(); else cluster_sync_all();       // no CTA may exit while its peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}
#endif  // !LU_HOST_EMU

// one task of the tcgen05 weight-gradient kernels below (plain data: the task lists are built on the host, and the host
// test build replays them with scalar loops)
struct LuWgTask {
  int16_t stage0, stage1;        // forward A stages giving rows [0,64) / [64,128); stage1 < 0: rows 64.. unused
  int16_t ntaps, a_is_lo;        // a_is_lo: bf16x3 lo plane of the activation (pairs with the hi plane of dY only)
  int32_t n0, nch;               // first packed column, number of 64-column chunks (1 or 2)
  int32_t tile0, tile1;          // pixel-tile range [tile0, tile1)
  int32_t off[4];                // tap offsets (rows) inside the window
  int32_t kb0[4], kb1[4];        // K block (64 rows of dWp) written by each tap for stage0 / stage1
  int32_t ychan[2];              // channel coordinate in dY of each 64-column chunk
};
// one task of the CTA-PAIR weight-gradient kernel: 256 output channels x (128 * nb) input channels x ntaps taps
struct LuWgPairTask {
  int16_t stage[4];              // forward A stages (64-channel chunks of ONE source with the same window geometry), in
                                 // order of the accumulator columns; CTA r of the pair stages [r*nb, (r+1)*nb)
  int16_t nb, ntaps;             // chunks per CTA (1 or 2); accumulator entries (their widths add up to <= 512 TMEM columns)
  int32_t n0;                    // first packed column of the 256-column slab: CTA r owns columns [n0 + 128 r, + 128)
  int32_t tile0, tile1;          // pixel-tile range [tile0, tile1)
  int32_t off[4];                // tap offsets (rows) inside the window
  int32_t disp[4];               // nb == 1 only: != 0 makes the entry a TAP PAIR -- a second tap `disp` window rows further is the
                                 // second 64-channel group of each CTA's B half (N = 256 instead of 128, the dY operand is read
                                 // once for both taps); 0 = a single tap
  int32_t kb[4][4];              // K block (64 rows of dWp) written by the entry's accumulator column group (col >> 6)
  int32_t ychan[4];              // channel coordinate in dY of the slab's four 64-column chunks
};

#ifndef LU_HOST_EMU
// =============================================================================================================
// Weight gradient on tcgen05: dWp[n][k] += sum over pixels  A_k[pixel] * dY[pixel][n]      (packed space)
//
// Both operands are NHWC, i.e. MN-major for a reduction over pixels: the MMA "K" rows are pixels (128-byte rows of
// 64 channels), exactly what the forward's halo windows already are.
// =============================================================================================================
struct LuWgParams {
  CUtensorMap tmA[LU_MAX_SRC];
  CUtensorMap tmY;
  LuConvParams cp;               // forward views / stage table
  const LuWgTask* tasks;
  const LuWgPairTask* ptasks;
  float* dWp;
  int32_t tiles_x, tiles_y, T, skip_t0_src;
  int32_t dy_frame_mul, dy_frame_add, dy_planes, dy_cpad;
  int32_t a_win_bytes, stage_bytes, n_stages;
  int32_t flush_scalar;          // pair kernel: 4-byte instead of 16-byte reductions in the flush (experiment switch)
};

namespace lutc {
__device__ __forceinline__ uint32_t desc_hi_mn(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t addr, uint32_t lbo_bytes) {
  return ((addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
}  // namespace lutc

// ---- independent CTAs -----------------------------------------------------------------------------------------------
// One task = (up to 2 activation stages = 128 rows of input channels, up to 4 taps, one 64/128-column slab of output
// channels, a range of pixel tiles); the 4 taps' accumulators (128 x N fp32 each) stay in TMEM for the whole pixel range
// and are flushed with fp32 atomics.  With M = N = 128 an MMA reads 8 KB of operands per 64 clocks = the whole shared
// memory bandwidth of the SM, on top of the TMA writes of the next stage: 70 % tensor-pipe activity (round-1 profile).
// Used for the layers / slabs the pair kernel below cannot take (single-chunk sources, < 256 output columns, bf16x3).
// (Round 2 measured and removed two cluster variants of THIS decomposition: operand boxes multicast to a CTA pair, and
// one M = 256 MMA per pair whose CTAs take different taps -- both 0 % on the step: the first leaves the shared-memory
// reads unchanged, the second re-stages every window once per CTA and moved 18 % more data from L2.)
__global__ void __launch_bounds__(256, 1) lu_wgrad_tc_kernel(const __grid_constant__ LuWgParams P) {
  using namespace lutc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  const int nS = P.n_stages;
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bars = s0 + (uint32_t)nS * P.stage_bytes;
  const uint32_t full = bars, empty = full + 8u * nS, done = empty + 8u * nS, tmem_slot = done + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + (size_t)nS * P.stage_bytes + 16u * nS + 8u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LuWgTask tk = P.tasks[blockIdx.x];
  const LuConvParams& cp = P.cp;
  const LuAStage st0 = cp.astages[tk.stage0];
  const LuAStage st1 = cp.astages[tk.stage1 >= 0 ? tk.stage1 : tk.stage0];
  const LuSrcView& v = cp.src[st0.src];
  const int N = tk.nch * 64;
  const int nyp = (P.dy_planes == 2 && !tk.a_is_lo) ? 2 : 1;
  const uint32_t b_off = 2u * (uint32_t)P.a_win_bytes;            // B region inside a stage
  const int tiles_per_frame = P.tiles_x * P.tiles_y;

  if (warp == 0 && lane == 0) { prefetch_tmap(&P.tmA[st0.src]); prefetch_tmap(&P.tmY); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nS; ++i) { mbar_init(full + 8u * i, 1); mbar_init(empty + 8u * i, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: activation windows + dY tiles
    int s = 0; uint32_t ph = 0;
    const uint32_t bytes = (uint32_t)(v.rows * v.pitch) * 128u * (tk.stage1 >= 0 ? 2u : 1u) + (uint32_t)(nyp * tk.nch) * 16384u;
    for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
      const int frame = tile / tiles_per_frame, rem = tile % tiles_per_frame;
      if (st0.src == P.skip_t0_src && (frame % P.T) == 0) continue;
      const int y0 = (rem / P.tiles_x) * LU_TILE_H, x0 = (rem % P.tiles_x) * LU_TILE_W;
      mbar_wait(empty + 8u * s, ph ^ 1u);
      if (elect_one()) {
        const uint32_t base = s0 + (uint32_t)s * P.stage_bytes;
        mbar_expect_tx(full + 8u * s, bytes);
        const int fa = frame * v.frame_mul + v.frame_add;
        tma_load_5d(base, &P.tmA[st0.src], full + 8u * s, st0.c, x0 + st0.dx, st0.plane, y0 + st0.dy, fa);
        if (tk.stage1 >= 0)
          tma_load_5d(base + (uint32_t)P.a_win_bytes, &P.tmA[st1.src], full + 8u * s, st1.c, x0 + st1.dx, st1.plane, y0 + st1.dy, fa);
        const int fy = frame * P.dy_frame_mul + P.dy_frame_add;
        for (int dp = 0; dp < nyp; ++dp)
          for (int c = 0; c < tk.nch; ++c)
            tma_load_5d(base + b_off + (uint32_t)(dp * tk.nch + c) * 16384u, &P.tmY, full + 8u * s, tk.ychan[c] + dp * P.dy_cpad, x0, 0, y0, fy);
      }
      __syncwarp();
      if (++s == nS) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (MN-major A and B)
    int s = 0; uint32_t ph = 0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = desc_hi_mn((uint32_t)v.pitch * 128u), b_hi = desc_hi_mn(1024u);
    const uint32_t a_lbo = tk.stage1 >= 0 ? (uint32_t)P.a_win_bytes : 0u;
    const uint32_t row2 = (uint32_t)v.pitch * 128u * 2u;           // two image rows = 16 pixels = one MMA K step
    uint32_t first = 1;
    for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
      const int frame = tile / tiles_per_frame;
      if (st0.src == P.skip_t0_src && (frame % P.T) == 0) continue;
      mbar_wait(full + 8u * s, ph);
      tc_fence_after();
      const uint32_t base = s0 + (uint32_t)s * P.stage_bytes;
      if (elect_one()) {
        for (int ti = 0; ti < tk.ntaps; ++ti) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(ti * N);
          const uint32_t a0 = base + (uint32_t)tk.off[ti] * 128u;
          for (int dp = 0; dp < nyp; ++dp) {
            const uint32_t b0 = base + b_off + (uint32_t)(dp * tk.nch) * 16384u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              mma_bf16(d_tmem, desc_lo_mn(a0 + (uint32_t)j * row2, a_lbo), a_hi, desc_lo_mn(b0 + (uint32_t)j * 2048u, 16384u), b_hi,
                       idesc, (first && dp == 0 && j == 0) ? 0u : 1u);
          }
        }
        tc_commit(empty + 8u * s);
      }
      __syncwarp();
      first = 0;
      if (++s == nS) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: TMEM -> fp32 atomics into dWp
    const int q = warp & 3;
    const int row = q * 32 + lane;                   // D row = channel row of the task
    // was anything accumulated?  (a task whose frames are all skipped leaves TMEM untouched)
    bool any = false;
    for (int tile = tk.tile0; tile < tk.tile1 && !any; ++tile)
      any = !(st0.src == P.skip_t0_src && ((tile / tiles_per_frame) % P.T) == 0);
    mbar_wait(done, 0);
    tc_fence_after();
    if (any) {
      const int half = row >> 6, kk = row & 63;
      for (int ti = 0; ti < tk.ntaps; ++ti) {
        const int kb = half ? (tk.stage1 >= 0 ? tk.kb1[ti] : -1) : tk.kb0[ti];
        const uint32_t taddr = tmem_base + (uint32_t)(ti * N) + ((uint32_t)(q * 32) << 16);
        for (int col = 0; col < N; col += 16) {
          float vv[16];
          tmem_ld16(taddr + (uint32_t)col, vv);
          tmem_wait16(vv);
          if (kb >= 0) {
            float* dst = P.dWp + (int64_t)(tk.n0 + col) * cp.ktot + (int64_t)kb * LU_KBLK + kk;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + (int64_t)j * cp.ktot, vv[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- CTA pairs: the TRANSPOSED product, one M = 256 MMA per pair (tcgen05.mma.cta_group::2) -------------------------
// D^T[output channel][input channel] per tap.  M = 256 output channels over the pair: CTA r stages ITS two 64-column
// chunks of the dY tile (the MMA's A operand, 128 rows per CTA).  N = 128 * nb input channels: CTA r stages ITS nb
// activation windows = its half of the MMA's B operand, and a tap is still a row offset into them.  Per CTA and K step
// the operand read drops from 8 KB / 64 clk to 6 KB / 64 clk (nb = 1) or 8 KB / 128 clk (nb = 2), and a stage moves
// 32 KB of dY + nb windows instead of 32 KB + 2 windows for the same MMA work (nb = 1).  Barriers as in the conv kernel's
// pair mode: all "full" barriers live in the leader (both producers' bytes complete there), "empty" / "done" are
// released in both CTAs by the leader's multicast commit.
__global__ void __launch_bounds__(256, 1) lu_wgrad_pair_kernel(const __grid_constant__ LuWgParams P) {
  using namespace lutc;
  const uint32_t crank = cluster_ctarank();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  const int nS = P.n_stages;
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bars = s0 + (uint32_t)nS * P.stage_bytes;
  const uint32_t full = bars, empty = full + 8u * nS, done = empty + 8u * nS, tmem_slot = done + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + (size_t)nS * P.stage_bytes + 16u * nS + 8u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LuWgPairTask tk = P.ptasks[blockIdx.x >> 1];
  const LuConvParams& cp = P.cp;
  const int nb = tk.nb;
  const LuAStage stw0 = cp.astages[tk.stage[crank * nb]];
  const LuAStage stw1 = cp.astages[tk.stage[crank * nb + nb - 1]];
  const LuSrcView& v = cp.src[stw0.src];
  const int N = 128 * nb;                                          // accumulator columns per tap (both CTAs' halves)
  const uint32_t w_off = 32768u;                                   // stage = [dY chunk a | dY chunk b | window(s)]
  const int tiles_per_frame = P.tiles_x * P.tiles_y;

  if (warp == 0 && lane == 0) { prefetch_tmap(&P.tmA[stw0.src]); prefetch_tmap(&P.tmY); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nS; ++i) { mbar_init(full + 8u * i, 1); mbar_init(empty + 8u * i, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer (both CTAs): own dY chunks + own windows
    int s = 0; uint32_t ph = 0;
    const uint32_t bytes_cta = (uint32_t)(v.rows * v.pitch) * 128u * (uint32_t)nb + 32768u;
    for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
      const int frame = tile / tiles_per_frame, rem = tile % tiles_per_frame;
      if (stw0.src == P.skip_t0_src && (frame % P.T) == 0) continue;
      const int y0 = (rem / P.tiles_x) * LU_TILE_H, x0 = (rem % P.tiles_x) * LU_TILE_W;
      mbar_wait(empty + 8u * s, ph ^ 1u);
      if (elect_one()) {
        const uint32_t base = s0 + (uint32_t)s * P.stage_bytes;
        const uint32_t lfull = map_to_rank(full + 8u * s, 0u);
        if (crank == 0) mbar_expect_tx(full + 8u * s, 2u * bytes_cta);            // the pair's bytes, counted on the leader
        const int fy = frame * P.dy_frame_mul + P.dy_frame_add;
        tma_load_5d_pair(base, &P.tmY, lfull, tk.ychan[2 * crank], x0, 0, y0, fy);
        tma_load_5d_pair(base + 16384u, &P.tmY, lfull, tk.ychan[2 * crank + 1], x0, 0, y0, fy);
        const int fa = frame * v.frame_mul + v.frame_add;
        tma_load_5d_pair(base + w_off, &P.tmA[stw0.src], lfull, stw0.c, x0 + stw0.dx, stw0.plane, y0 + stw0.dy, fa);
        if (nb == 2)
          tma_load_5d_pair(base + w_off + (uint32_t)P.a_win_bytes, &P.tmA[stw1.src], lfull, stw1.c, x0 + stw1.dx, stw1.plane, y0 + stw1.dy, fa);
      }
      __syncwarp();
      if (++s == nS) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: the even CTA issues for the pair
    if (crank == 0) {
      int s = 0; uint32_t ph = 0;
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t a_hi = desc_hi_mn(1024u), b_hi = desc_hi_mn((uint32_t)v.pitch * 128u);
      const uint32_t b_lbo = nb == 2 ? (uint32_t)P.a_win_bytes : 0u;
      const uint32_t row2 = (uint32_t)v.pitch * 128u * 2u;         // two image rows = 16 pixels = one MMA K step
      uint32_t first = 1;
      for (int tile = tk.tile0; tile < tk.tile1; ++tile) {
        const int frame = tile / tiles_per_frame;
        if (stw0.src == P.skip_t0_src && (frame % P.T) == 0) continue;
        mbar_wait(full + 8u * s, ph);
        tc_fence_after();
        const uint32_t base = s0 + (uint32_t)s * P.stage_bytes;
        if (elect_one()) {
          uint32_t col0 = 0;
          for (int ti = 0; ti < tk.ntaps; ++ti) {
            const bool tap_pair = tk.disp[ti] != 0;
            const uint32_t Ni = tap_pair ? 256u : (uint32_t)N;
            const uint32_t lbo = tap_pair ? (uint32_t)tk.disp[ti] * 128u : b_lbo;
            const uint32_t idesc = idesc0 | ((Ni >> 3) << 17);
            const uint32_t d_tmem = tmem_base + col0;
            const uint32_t b0 = base + w_off + (uint32_t)tk.off[ti] * 128u;
            col0 += Ni;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              mma_bf16_pair(d_tmem, desc_lo_mn(base + (uint32_t)j * 2048u, 16384u), a_hi, desc_lo_mn(b0 + (uint32_t)j * row2, lbo), b_hi,
                            idesc, (first && j == 0) ? 0u : 1u);
          }
          tc_commit_pair(empty + 8u * s, (uint16_t)3);             // frees the stage in both CTAs
        }
        __syncwarp();
        first = 0;
        if (++s == nS) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) tc_commit_pair(done, (uint16_t)3);          // both epilogues
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs): own 128 output channels
    const int q = warp & 3;
    const int row = q * 32 + lane;
    bool any = false;
    for (int tile = tk.tile0; tile < tk.tile1 && !any; ++tile)
      any = !(stw0.src == P.skip_t0_src && ((tile / tiles_per_frame) % P.T) == 0);
    mbar_wait(done, 0);
    tc_fence_after();
    if (any) {
      float* drow = P.dWp + (int64_t)(tk.n0 + (int)crank * 128 + row) * cp.ktot;
      uint32_t col0 = 0;
      for (int ti = 0; ti < tk.ntaps; ++ti) {
        const int Ni = tk.disp[ti] != 0 ? 256 : N;
        const uint32_t taddr = tmem_base + col0 + ((uint32_t)(q * 32) << 16);
        col0 += (uint32_t)Ni;
        for (int col = 0; col < Ni; col += 16) {
          float vv[16];
          tmem_ld16(taddr + (uint32_t)col, vv);
          tmem_wait16(vv);
          float* dst = drow + (int64_t)tk.kb[ti][col >> 6] * LU_KBLK + (col & 63);      // 16 consecutive floats of this row
          // a warp's 32 lanes are 32 rows of dWp = 32 different sectors per instruction whatever its width: four 16-byte
          // reductions per 16 floats instead of sixteen 4-byte ones (r2_ncu_prof_wgrad_pair.txt: 367 M atomic sectors per
          // launch = ~64 k clocks of flush per task at one sector per clock, on top of ~460 k clocks of MMAs)
          if (P.flush_scalar) {                                      // LU_WGRAD_RED4=0: the round-2 form, for A/B runs
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + j, vv[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(vv[j]), "f"(vv[j + 1]), "f"(vv[j + 2]),
                           "f"(vv[j + 3]) : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                               // no CTA may exit while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}
#endif  // !LU_HOST_EMU
