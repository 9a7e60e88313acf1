// Bandwidth-bound kernels around the convolutions: input staging, up-sampling, batch-norm, soft-max, weight
// packing, recurrent-state maintenance.  Written as index functors so the same code runs on the GPU
// (grid-stride kernel) and, in the TEST-ONLY host build (LU_HOST_EMU), as plain loops.
#pragma once
#include "lu_defs.h"

#ifdef LU_HOST_EMU
template <class F>
static void lu_parallel_for_impl(int64_t n, void* /*stream*/, F f) {
  for (int64_t i = 0; i < n; ++i) f(i);
}
#else
template <class F>
__global__ void __launch_bounds__(256) lu_pf_kernel(int64_t n, F f) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i);
}
template <class F>
static void lu_parallel_for_impl(int64_t n, void* stream, F f) {
  if (n <= 0) return;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148 * 32;
  if (blocks > cap) blocks = cap;
  lu_pf_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, f);
}
#endif
// 8 x bf16 (16 bytes) loads / stores
LU_HDI void lu_load8_bf16(const uint16_t* src, float* v) {
#ifdef __CUDA_ARCH__
  const uint4 q = *reinterpret_cast<const uint4*>(src);
  v[0] = lu_u2f(q.x << 16); v[1] = lu_u2f(q.x & 0xffff0000u); v[2] = lu_u2f(q.y << 16); v[3] = lu_u2f(q.y & 0xffff0000u);
  v[4] = lu_u2f(q.z << 16); v[5] = lu_u2f(q.z & 0xffff0000u); v[6] = lu_u2f(q.w << 16); v[7] = lu_u2f(q.w & 0xffff0000u);
#else
  for (int j = 0; j < 8; ++j) v[j] = lu_bf2f(src[j]);
#endif
}
LU_HDI void lu_load8_h16(const uint16_t* src, float* v, int fmt) {
  if (!fmt) { lu_load8_bf16(src, v); return; }
#ifdef __CUDA_ARCH__
  const uint4 q = *reinterpret_cast<const uint4*>(src);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
#else
  for (int j = 0; j < 8; ++j) v[j] = lu_half2f(src[j]);
#endif
}
LU_HDI void lu_store8_bf16(uint16_t* dst, const uint16_t* h) {
#ifdef __CUDA_ARCH__
  uint4 a;
  a.x = h[0] | ((uint32_t)h[1] << 16); a.y = h[2] | ((uint32_t)h[3] << 16);
  a.z = h[4] | ((uint32_t)h[5] << 16); a.w = h[6] | ((uint32_t)h[7] << 16);
  *reinterpret_cast<uint4*>(dst) = a;
#else
  for (int j = 0; j < 8; ++j) dst[j] = h[j];
#endif
}

LU_HDI void lu_ld8f(const float* p, float* v) {
#ifdef __CUDA_ARCH__
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
  for (int j = 0; j < 8; ++j) v[j] = p[j];
#endif
}
LU_HDI void lu_st8f(float* p, const float* v) {
#ifdef __CUDA_ARCH__
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
#else
  for (int j = 0; j < 8; ++j) p[j] = v[j];
#endif
}

// ---- row-loop kernels ---------------------------------------------------------------------------------------------
// For the passes over an NHWC tensor that need per-CHANNEL constants or per-channel sums (BatchNorm statistics / apply /
// backward, bias gradients): a thread owns ONE group of 8 channels for its whole life and walks pixels, so the constants
// are loaded once into registers, there is no index division per element, and channel sums are accumulated in registers,
// combined across the block's pixel lanes in shared memory and flushed with one atomic per channel and block.
//   struct F { struct State; static constexpr int NSUM;
//              void begin(int group, State&) const; void pixel(int64_t p, int group, State&) const;
//              void partials(const State&, float* part) const; void flush(int group, const float* part) const; }  (last two: NSUM > 0)
// Threads of a block: (pixel lane, group) with the group fastest, so a warp reads consecutive 16 / 32-byte pieces of a pixel.
#ifdef LU_HOST_EMU
template <class F>
static void lu_rows_impl(int64_t npix, int ngroups, void* /*stream*/, F f) {
  for (int g = 0; g < ngroups; ++g) {
    typename F::State st;
    f.begin(g, st);
    for (int64_t p = 0; p < npix; ++p) f.pixel(p, g, st);
    if (F::NSUM > 0) {
      float part[F::NSUM > 0 ? F::NSUM : 1];
      f.partials(st, part);
      f.flush(g, part);
    }
  }
}
#else
template <class F>
__global__ void __launch_bounds__(256) lu_rows_kernel(int64_t npix, int ngroups, F f) {
  constexpr int NS = F::NSUM > 0 ? F::NSUM : 1;
  __shared__ float red[F::NSUM > 0 ? F::NSUM * 256 : 1];
  const int gpb = ngroups < 256 ? ngroups : 256, PL = 256 / gpb;
  const int gl = (int)threadIdx.x % gpb, pl = (int)threadIdx.x / gpb;
  const int g = (int)blockIdx.y * gpb + gl;
  const bool active = pl < PL && g < ngroups;
  typename F::State st;
  if (active) {
    f.begin(g, st);
#pragma unroll 2
    for (int64_t p = (int64_t)blockIdx.x * PL + pl; p < npix; p += (int64_t)gridDim.x * PL) f.pixel(p, g, st);
  }
  if (F::NSUM > 0) {
    float part[NS];
    if (active) f.partials(st, part);
    else {
#pragma unroll
      for (int k = 0; k < NS; ++k) part[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) red[k * 256 + threadIdx.x] = part[k];
    __syncthreads();
    if (pl == 0 && g < ngroups) {
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        float s = 0.f;
        for (int q = 0; q < PL; ++q) s += red[k * 256 + q * gpb + gl];
        part[k] = s;
      }
      f.flush(g, part);
    }
  }
}
template <class F>
static void lu_rows_impl(int64_t npix, int ngroups, void* stream, F f) {
  if (npix <= 0 || ngroups <= 0) return;
  const int gpb = ngroups < 256 ? ngroups : 256, PL = 256 / gpb;
  const int gy = (ngroups + gpb - 1) / gpb;
  int64_t bx = (npix + PL - 1) / PL;
  // streaming passes: two waves of resident blocks.  Reductions: every block ends with one atomic per channel on the SAME
  // few hundred addresses (serialised in L2), so fewer, longer-lived blocks: 4 per SM, one wave
  int64_t cap = (F::NSUM > 0 ? 148 * 4 : 148 * 8 * 2) / gy;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  lu_rows_kernel<<<dim3((unsigned)bx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(npix, ngroups, f);
}
#endif

LU_HDI int lu_reflect(int i, int n) {     // tf.pad REFLECT index (no edge repeat); pad < n guaranteed
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// ---- input staging: reflect-pad (Networks.py:232) + pw x pw patch extraction of the 1-channel image ------------
// out (N,Hp,Wp,64) bf16.  Channel t < pw*pw holds the padded image at (y + t/pw - (pw-1)/2, x + t%pw - (pw-1)/2),
// zero outside the padded image (that is the SAME zero padding of the consuming convolutions).  In bf16x3 mode
// channels [32, 32+pw*pw) hold the lo parts.
// PW > 0: compile-time window size -- the loops unroll, every r[] index is a constant and the 64 channel values stay in
// registers (with a run-time window the array lives in local memory: 128 bytes of stack per thread, ptxas -v);
// PW == 0: any window size (run-time loops).  The host picks the instantiation (5 and 3 are the reference's kernels).
template <int PW>
struct LuPrepPatchesT {
  const float* x; uint16_t* out;
  int H, W, Hp, Wp, pad_y0, pad_x0, pw, x3, fmt;
  LU_HD void operator()(int64_t p) const {        // item = one pixel of the padded frame: 64 channels = 128 bytes
    const int w_ = PW > 0 ? PW : pw;
    const int xx = (int)(p % Wp); int64_t q = p / Wp;
    const int yy = (int)(q % Hp); const int64_t n = q / Hp;
    uint16_t r[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) r[j] = 0;
    const int c = (w_ - 1) / 2;
    const float* img = x + n * (int64_t)H * W;
#pragma unroll
    for (int dy = 0; dy < w_; ++dy) {
      const int py = yy + dy - c;
      if (py < 0 || py >= Hp) continue;
      const float* row = img + (int64_t)lu_reflect(py - pad_y0, H) * W;
#pragma unroll
      for (int dx = 0; dx < w_; ++dx) {
        const int px = xx + dx - c;
        if (px < 0 || px >= Wp) continue;
        const float v = row[lu_reflect(px - pad_x0, W)];
        const int t = dy * w_ + dx;
        if (x3) { uint16_t h, l; lu_split(v, h, l); r[t] = h; r[32 + t] = l; }
        else r[t] = lu_f2h16(v, fmt);
      }
    }
    uint16_t* o = out + p * 64;
#pragma unroll
    for (int g = 0; g < 8; ++g) lu_store8_bf16(o + g * 8, r + g * 8);
  }
};

// multi-channel images (in_channels > 1): reflect-pad + layout change into an ordinary NHWC 16-bit activation buffer
// x: (N,C,H,W) [channels_first] or (N,H,W,C); out (N,Hp,Wp,planes*cpad); item = (padded pixel, group of 8 channels)
struct LuPrepImage {
  const float* x; uint16_t* out;
  int C, H, W, Hp, Wp, pad_y0, pad_x0, cpad, planes, fmt, channels_first;
  LU_HD void operator()(int64_t i) const {
    const int cg = cpad / 8;
    const int c0 = (int)(i % cg) * 8; const int64_t p = i / cg;
    const int xx = (int)(p % Wp); const int64_t q = p / Wp;
    const int yy = (int)(q % Hp); const int64_t n = q / Hp;
    const int sy = lu_reflect(yy - pad_y0, H), sx = lu_reflect(xx - pad_x0, W);
    uint16_t hi[8], lo[8];
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float v = 0.f;
      if (c < C) v = channels_first ? x[((n * C + c) * H + sy) * (int64_t)W + sx] : x[((n * H + sy) * (int64_t)W + sx) * C + c];
      if (planes == 2) lu_split(v, hi[j], lo[j]);
      else hi[j] = lu_f2h16(v, fmt);
    }
    uint16_t* o = out + p * (int64_t)(cpad * planes) + c0;
    lu_store8_bf16(o, hi);
    if (planes == 2) lu_store8_bf16(o + cpad, lo);
  }
};

// ---- bilinear x2 (tf.image.resize half-pixel centres == resize_images, Networks.py:143) -------------------------
// in (N,h,w,planes*cpad) -> out (N,2h,2w,planes*cpad); value = hi + lo, re-split on store.
struct LuUpsample2x {
  const uint16_t* in; uint16_t* out;
  int h, w, cpad, planes, fmt;
  LU_HD void load(const uint16_t* b, int y, int x, float* v) const {
    const int ct = cpad * planes;
    lu_load8_h16(b + ((int64_t)y * w + x) * ct, v, fmt);
    if (planes == 2) {
      float t[8];
      lu_load8_bf16(b + ((int64_t)y * w + x) * ct + cpad, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
  }
  // item = (n, iy, ix, group of 8 channels): the 2x2 output quad of one input pixel from its 3x3 neighbourhood
  // (9 loads for 4 outputs).  out[2i] = .25 in[i-1] + .75 in[i]; out[2i+1] = .75 in[i] + .25 in[i+1] (edge clamped)
  LU_HD void operator()(int64_t i) const {
    const int cg = cpad / 8;
    const int c = (int)(i % cg) * 8; int64_t p = i / cg;
    const int ix = (int)(p % w); p /= w;
    const int iy = (int)(p % h); const int64_t n = p / h;
    const int ym = iy > 0 ? iy - 1 : 0, yp = iy + 1 < h ? iy + 1 : h - 1;
    const int xm = ix > 0 ? ix - 1 : 0, xp = ix + 1 < w ? ix + 1 : w - 1;
    const int ct = cpad * planes;
    const uint16_t* b = in + n * (int64_t)h * w * ct + c;
    float v[3][3][8];
    load(b, ym, xm, v[0][0]); load(b, ym, ix, v[0][1]); load(b, ym, xp, v[0][2]);
    load(b, iy, xm, v[1][0]); load(b, iy, ix, v[1][1]); load(b, iy, xp, v[1][2]);
    load(b, yp, xm, v[2][0]); load(b, yp, ix, v[2][1]); load(b, yp, xp, v[2][2]);
#pragma unroll
    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
      for (int qx = 0; qx < 2; ++qx) {
        const int fy = qy ? 2 : 0, fx = qx ? 2 : 0;          // the "far" row / column of this output
        uint16_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float top = 0.75f * v[1][1][j] + 0.25f * v[1][fx][j], bot = 0.75f * v[fy][1][j] + 0.25f * v[fy][fx][j];
          if (planes == 2) lu_split(0.75f * top + 0.25f * bot, hi[j], lo[j]);
          else hi[j] = lu_f2h16(0.75f * top + 0.25f * bot, fmt);
        }
        uint16_t* o = out + (((n * 2 * h + 2 * iy + qy) * 2 * w) + 2 * ix + qx) * (int64_t)ct + c;
        lu_store8_bf16(o, hi);
        if (planes == 2) lu_store8_bf16(o + cpad, lo);
      }
  }
};

// ---- batch norm (keras BatchNormalization, eps 1e-3, momentum .99) ----------------------------
// pass 1 (row-loop): per-channel sum / sum of squares of the fp32 conv output about a per-channel shift
struct LuBnStats {
  const float* raw; double* sums;      // sums[0:cpad] = sum, sums[cpad:2cpad] = sum sq (about a per-channel shift)
  const float* shift_src;              // per-channel shift (first pixel) to avoid cancellation
  int cpad, c_real;
  struct State { float sh[8], s[8], s2[8]; };
  static constexpr int NSUM = 16;
  LU_HD void begin(int g, State& st) const {
    lu_ld8f(shift_src + g * 8, st.sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) { st.s[j] = 0.f; st.s2[j] = 0.f; }
  }
  LU_HD void pixel(int64_t p, int g, State& st) const {
    float v[8];
    lu_ld8f(raw + p * cpad + g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - st.sh[j]; st.s[j] += d; st.s2[j] += d * d; }
  }
  LU_HD void partials(const State& st, float* part) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) { part[j] = st.s[j]; part[8 + j] = st.s2[j]; }
  }
  LU_HD void flush(int g, const float* part) const {
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      if (c >= c_real) continue;
      lu_atomic_add(&sums[c], (double)part[j]);
      lu_atomic_add(&sums[cpad + c], (double)part[8 + j]);
    }
  }
};
// pass 2: scale/shift from batch statistics + moving-statistics update; item = channel
struct LuBnFinalize {
  double* sums; const float* shift_src; const float* gamma; const float* beta; float* mov_mean; float* mov_var;
  float* scale; float* shift; float* save_mean; float* save_invstd;
  int64_t npix; int cpad, c_real; float eps, momentum;
  LU_HD void operator()(int64_t c) const {
    if (c >= c_real) { scale[c] = 0.f; shift[c] = 0.f; return; }
    const double n = (double)npix;
    const double m0 = sums[c] / n;
    double var = sums[cpad + c] / n - m0 * m0; if (var < 0) var = 0;
    const double mean = m0 + (double)shift_src[c];
    const double inv = 1.0 / sqrt(var + (double)eps);
    scale[c] = (float)(gamma[c] * inv);
    shift[c] = (float)(beta[c] - mean * gamma[c] * inv);
    if (save_mean) { save_mean[c] = (float)mean; save_invstd[c] = (float)inv; }
    const double unb = var * (n / (n > 1 ? n - 1 : 1));
    mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * (float)mean;
    mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * (float)unb;
    sums[c] = 0; sums[cpad + c] = 0;
  }
};
// ---- synchronised batch statistics (data-parallel training, SURVEY 8e option ii) -----------------------------------
// local moments of this rank's frames in a form that adds up across ranks: mom = [mean | mean^2 | M2] per channel
// (every rank holds the same number of pixels); item = channel
struct LuBnMoments {
  double* sums; const float* shift_src; double* mom; int64_t npix; int cpad, c_real;
  LU_HD void operator()(int64_t c) const {
    double mean = 0, m2 = 0;
    if (c < c_real) {
      const double n = (double)npix;
      const double m0 = sums[c] / n;
      double var = sums[cpad + c] / n - m0 * m0; if (var < 0) var = 0;
      mean = m0 + (double)shift_src[c]; m2 = var * n;
    }
    mom[c] = mean; mom[cpad + c] = mean * mean; mom[2 * cpad + c] = m2;
    sums[c] = 0; sums[cpad + c] = 0;
  }
};
// scale / shift / moving statistics from the moments summed over `world` ranks (Chan's pairwise combination for equal
// counts: M2 = sum M2_r + n * (sum mean_r^2 - world * mean^2)); item = channel
struct LuBnFinalizeSync {
  const double* mom; const float* gamma; const float* beta; float* mov_mean; float* mov_var;
  float* scale; float* shift; float* save_mean; float* save_invstd;
  int64_t npix; int cpad, c_real, world; float eps, momentum;
  LU_HD void operator()(int64_t c) const {
    if (c >= c_real) { scale[c] = 0.f; shift[c] = 0.f; return; }
    const double n = (double)npix, R = (double)world, N = n * R;
    const double mean = mom[c] / R;
    double m2 = mom[2 * cpad + c] + n * (mom[cpad + c] - R * mean * mean); if (m2 < 0) m2 = 0;
    const double var = m2 / N;
    const double inv = 1.0 / sqrt(var + (double)eps);
    scale[c] = (float)(gamma[c] * inv);
    shift[c] = (float)(beta[c] - mean * gamma[c] * inv);
    if (save_mean) { save_mean[c] = (float)mean; save_invstd[c] = (float)inv; }
    const double unb = var * (N / (N > 1 ? N - 1 : 1));
    mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * (float)mean;
    mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * (float)unb;
  }
};
// pass 3 (row-loop): y = lrelu(raw*scale + shift) -> 16-bit planes; a thread keeps scale / shift of its 8 channels
struct LuBnApply {
  const float* raw; const float* scale; const float* shift; uint16_t* out;
  int raw_cpad, out_cpad, planes, fmt; float alpha;
  struct State { float sc[8], sh[8]; };
  static constexpr int NSUM = 0;
  LU_HD void begin(int g, State& st) const {
    if (g * 8 < raw_cpad) { lu_ld8f(scale + g * 8, st.sc); lu_ld8f(shift + g * 8, st.sh); }
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { st.sc[j] = 0.f; st.sh[j] = 0.f; }
    }
  }
  LU_HD void pixel(int64_t p, int g, State& st) const {
    const int c = g * 8;
    float r[8];
    if (c < raw_cpad) lu_ld8f(raw + p * raw_cpad + c, r);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
    }
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = r[j] * st.sc[j] + st.sh[j];
      a = a > 0.f ? a : alpha * a;
      if (planes == 2) lu_split(a, hi[j], lo[j]);
      else hi[j] = lu_f2h16(a, fmt);
    }
    uint16_t* o = out + p * (int64_t)(out_cpad * planes) + c;
    lu_store8_bf16(o, hi);
    if (planes == 2) lu_store8_bf16(o + out_cpad, lo);
  }
  LU_HD void partials(const State&, float*) const {}
  LU_HD void flush(int, const float*) const {}
};
// inference: fold moving statistics into per-channel scale/shift (applied in the conv epilogue)
struct LuBnFold {
  const float* gamma; const float* beta; const float* mov_mean; const float* mov_var; float* scale; float* shift;
  int c_real; float eps;
  LU_HD void operator()(int64_t c) const {
    if (c >= c_real) { scale[c] = 0.f; shift[c] = 0.f; return; }
    const float inv = 1.0f / sqrtf(mov_var[c] + eps);
    scale[c] = gamma[c] * inv;
    shift[c] = beta[c] - mov_mean[c] * gamma[c] * inv;
  }
};

// ---- crop + soft-max (Networks.py:250-252) ------------------------------------------------------------------------
// raw (B*T,Hp,Wp,raw_cpad) fp32 -> logits/softmax in the API layout.  channels_first: (B,T,D,H,W), softmax over D.
// channels_last: (B,T,H,W,D) and -- reproducing the reference's Softmax(channel_axis+1) -- softmax over the BATCH axis.
struct LuSoftmaxCrop {
  const float* raw; float* logits; float* softmax;
  int B, T, H, W, Hp, Wp, py0, px0, raw_cpad, D, channels_first;
  LU_HD void operator()(int64_t i) const {
    if (channels_first) {      // item = (b,t,y,x)
      const int x = (int)(i % W); int64_t p = i / W;
      const int y = (int)(p % H); const int64_t f = p / H;
      const float* r = raw + ((f * Hp + y + py0) * Wp + x + px0) * (int64_t)raw_cpad;
      float mx = r[0];
      for (int d = 1; d < D; ++d) mx = fmaxf(mx, r[d]);
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += expf(r[d] - mx);
      for (int d = 0; d < D; ++d) {
        const int64_t o = ((f * D + d) * H + y) * W + x;
        logits[o] = r[d];
        softmax[o] = expf(r[d] - mx) / s;
      }
    } else {                   // item = (t,y,x,d); loop over b
      const int d = (int)(i % D); int64_t p = i / D;
      const int x = (int)(p % W); p /= W;
      const int y = (int)(p % H); const int t = (int)(p / H);
      float mx = -3.4e38f;
      for (int b = 0; b < B; ++b)
        mx = fmaxf(mx, raw[((((int64_t)b * T + t) * Hp + y + py0) * Wp + x + px0) * raw_cpad + d]);
      float s = 0.f;
      for (int b = 0; b < B; ++b)
        s += expf(raw[((((int64_t)b * T + t) * Hp + y + py0) * Wp + x + px0) * raw_cpad + d] - mx);
      for (int b = 0; b < B; ++b) {
        const float v = raw[((((int64_t)b * T + t) * Hp + y + py0) * Wp + x + px0) * raw_cpad + d];
        const int64_t o = ((((int64_t)b * T + t) * H + y) * W + x) * D + d;
        logits[o] = v;
        softmax[o] = expf(v - mx) / s;
      }
    }
  }
};

// ---- weight packing: Keras HWIO fp32 -> K-major bf16 [Npad][ktot] in table order ----------------------------------
struct LuPackWeights {
  const float* params; const LuPackDesc* descs; uint16_t* out; LuColMap cm; int ktot, fmt;
  LU_HD void operator()(int64_t i) const {
    const int k = (int)(i % ktot); const int n = (int)(i / ktot);
    const int kb = k / LU_KBLK, kk = k % LU_KBLK;
    const int col = lu_col_of(cm, n);
    uint16_t r = 0;
    if (col >= 0) {
      const LuPackDesc d = descs[kb];
      float w = 0.f; bool ok = false;
      if (d.kind == 0 && d.transposed) {
        if (kk < d.n_valid) { w = params[d.w_off + d.tap_off + (int64_t)(d.col_base + col) * d.cout_total + d.c_base + kk]; ok = true; }
      } else if (d.kind == 0) {
        if (kk < d.n_valid) { w = params[d.w_off + d.tap_off + (int64_t)(d.c_base + kk) * d.cout_total + col]; ok = true; }
      } else {
        int t = kk; bool z = false;
        if (d.patch_x3 && kk >= 32) { t = kk - 32; z = d.patch_hi_only != 0; }
        if (!z && t < d.pw * d.pw) {
          const int sh = (d.pw - 1) / 2 - (d.k - 1) / 2;
          const int ky = t / d.pw - sh, kx = t % d.pw - sh;
          if (ky >= 0 && ky < d.k && kx >= 0 && kx < d.k) {
            w = params[d.w_off + ((int64_t)(ky * d.k + kx) * d.cin_total + d.c_base) * d.cout_total + col]; ok = true;
          }
        }
      }
      if (ok) {
        if (fmt) r = lu_f2half(w);
        else { uint16_t hi, lo; lu_split(w, hi, lo); r = d.wpart ? lo : hi; }
      }
    }
    out[i] = r;
  }
};
struct LuPackVec {       // per-column vectors (bias, gamma...) into packed column order
  const float* src; float* out; LuColMap cm;
  LU_HD void operator()(int64_t n) const { const int col = lu_col_of(cm, (int)n); out[n] = col >= 0 ? src[col] : 0.f; }
};

// ---- recurrent state maintenance (Networks.py:77-98) ---------------------------------------------------------------
struct LuStateMask {     // h *= mask[b] (bf16 planes; mask is 0/1 in the reference's use, so hi/lo scale exactly)
  uint16_t* h; float* c; const float* mask; int64_t per_sample_h, per_sample_c; int fmt;
  LU_HD void operator()(int64_t i) const { h[i] = lu_f2h16(lu_h162f(h[i], fmt) * mask[i / per_sample_h], fmt); }
};
struct LuStateMaskC {
  float* c; const float* mask; int64_t per_sample;
  LU_HD void operator()(int64_t i) const { c[i] *= mask[i / per_sample]; }
};
// internal (B,H,W,planes*fpad) bf16 / (B,H,W,fpad) fp32  <->  API (B,F,H,W) or (B,H,W,F) fp32
struct LuStateGet {
  const uint16_t* h; const float* c; float* out; int which, H, W, F, fpad, planes, channels_first, fmt;
  LU_HD void operator()(int64_t i) const {
    int f, y, x; int64_t b;
    if (channels_first) { x = (int)(i % W); int64_t p = i / W; y = (int)(p % H); p /= H; f = (int)(p % F); b = p / F; }
    else { f = (int)(i % F); int64_t p = i / F; x = (int)(p % W); p /= W; y = (int)(p % H); b = p / H; }
    const int64_t pix = (b * H + y) * W + x;
    if (which == 1) out[i] = c[pix * fpad + f];
    else {
      float v = lu_h162f(h[pix * (int64_t)(fpad * planes) + f], fmt);
      if (planes == 2) v += lu_bf2f(h[pix * (int64_t)(fpad * planes) + fpad + f]);
      out[i] = v;
    }
  }
};
struct LuStateSet {
  uint16_t* h; float* c; const float* in; int which, H, W, F, fpad, planes, channels_first, fmt;
  LU_HD void operator()(int64_t i) const {
    int f, y, x; int64_t b;
    if (channels_first) { x = (int)(i % W); int64_t p = i / W; y = (int)(p % H); p /= H; f = (int)(p % F); b = p / F; }
    else { f = (int)(i % F); int64_t p = i / F; x = (int)(p % W); p /= W; y = (int)(p % H); b = p / H; }
    const int64_t pix = (b * H + y) * W + x;
    const float v = in ? in[i] : 0.f;
    if (which == 1) c[pix * fpad + f] = v;
    else {
      uint16_t hi, lo; lu_split(v, hi, lo);
      if (fmt) hi = lu_f2half(v);
      h[pix * (int64_t)(fpad * planes) + f] = hi;
      if (planes == 2) h[pix * (int64_t)(fpad * planes) + fpad + f] = lo;
    }
  }
};

struct LuDebugRead {     // bf16 planes NHWC (padded channels) -> fp32 NHWC (real channels)
  const uint16_t* src; float* out; int creal, cpad, planes, fmt;
  LU_HD void operator()(int64_t i) const {
    const int c = (int)(i % creal); const int64_t p = i / creal;
    float v = lu_h162f(src[p * (int64_t)(cpad * planes) + c], fmt);
    if (planes == 2) v += lu_bf2f(src[p * (int64_t)(cpad * planes) + cpad + c]);
    out[i] = v;
  }
};

// stand-alone blocks: the returned tensor as fp32 in the caller's layout, from the 16-bit activation planes (NHWC, padded
// channels) or from a convolution's fp32 raw output; item = one output element
struct LuBlockOut {
  const uint16_t* act; const float* raw; float* out;
  int C, H, W, cpad, planes, fmt, channels_first;
  LU_HD void operator()(int64_t i) const {
    int c, y, x; int64_t n;
    if (channels_first) { x = (int)(i % W); int64_t q = i / W; y = (int)(q % H); q /= H; c = (int)(q % C); n = q / C; }
    else { c = (int)(i % C); int64_t q = i / C; x = (int)(q % W); q /= W; y = (int)(q % H); n = q / H; }
    const int64_t p = (n * H + y) * (int64_t)W + x;
    float v;
    if (raw) v = raw[p * cpad + c];
    else {
      v = lu_h162f(act[p * (int64_t)(cpad * planes) + c], fmt);
      if (planes == 2) v += lu_bf2f(act[p * (int64_t)(cpad * planes) + cpad + c]);
    }
    out[i] = v;
  }
};

// generic conv-mirror launcher functor
struct LuMirrorItem {
  LuConvParams p;
  LU_HD void operator()(int64_t i) const { lu_conv_mirror_item(p, i); }
};
