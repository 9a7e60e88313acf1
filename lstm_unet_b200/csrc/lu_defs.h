// Shared POD types and portable helpers for the ConvLSTM-UNet kernels.
//
// Everything in this header compiles two ways:
//   * nvcc, sm_100a           -> the product library (liblstm_unet_b200.so)
//   * g++ with -DLU_HOST_EMU  -> a TEST-ONLY host build (tests/_emu) in which the scalar "mirror" kernels and
//     the elementwise kernels run as plain loops, so that plan/table generation, weight packing and epilogue
//     maths can be checked against the oracle in a container without a GPU.  The product library never
//     contains that path (lu_is_cuda_build() == 1) and has no CPU fallback.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef LU_HOST_EMU
#define LU_HD
#define LU_HDI inline
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#define LU_HD __host__ __device__
#define LU_HDI __host__ __device__ __forceinline__
#endif

#define LU_MAX_SRC 4
#define LU_TILE_H 16      // output pixels per tile: 16 rows x 8 columns = 128 = UMMA M
#define LU_TILE_W 8
#define LU_KBLK 64        // bf16 elements per K block (= one 128-byte swizzle row)

// ---- bf16 <-> fp32 (round to nearest even), portable -----------------------------------------------------
LU_HDI uint32_t lu_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LU_HDI float lu_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
LU_HDI uint16_t lu_f2bf(float f) {
#ifdef __CUDA_ARCH__
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));     // cvt.rn.bf16.f32: same RNE result for every finite value
#else
  uint32_t u = lu_f2u(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
#endif
}
LU_HDI float lu_bf2f(uint16_t h) { return lu_u2f(((uint32_t)h) << 16); }
// split v into hi + lo bf16 parts (lo = bf16(v - hi)); used by the bf16x3 precision mode
LU_HDI void lu_split(float v, uint16_t& hi, uint16_t& lo) {
  hi = lu_f2bf(v);
  lo = lu_f2bf(v - lu_bf2f(hi));
}

// ---- fp16 <-> fp32 (round to nearest even, subnormals kept), portable; and the handle's 16-bit operand format ------
// fmt 0 = bf16 (8-bit mantissa, fp32 range; the training / throughput format), fmt 1 = fp16 (11-bit mantissa, inference:
// same tensor-core rate, 8x smaller operand rounding error -- the 1e-3 mode at full speed)
LU_HDI uint16_t lu_f2half(float f) {
#ifdef __CUDA_ARCH__
  return __half_as_ushort(__float2half_rn(f));
#else
  uint32_t x = lu_f2u(f);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);                    // NaN
  if (x >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);                   // >= 65536 (and inf)
  const int e = (int)(x >> 23) - 127 + 15;
  uint32_t mant = x & 0x7fffffu, half;
  if (e >= 1) {
    half = ((uint32_t)e << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ++half;            // a carry into the exponent is correct (up to inf)
  } else {
    if (e < -10) return (uint16_t)sign;                                      // below half the smallest subnormal
    mant |= 0x800000u;
    const int shift = 14 - e;
    half = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half & 1u))) ++half;
  }
  return (uint16_t)(sign | half);
#endif
}
LU_HDI float lu_half2f(uint16_t h) {
#ifdef __CUDA_ARCH__
  return __half2float(__ushort_as_half(h));
#else
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, mant = h & 0x3ffu;
  if (e == 0) { const float v = (float)mant * 5.9604644775390625e-8f; return sign ? -v : v; }   // mant * 2^-24
  if (e == 31) return lu_u2f(sign | 0x7f800000u | (mant << 13));
  return lu_u2f(sign | ((e + 112u) << 23) | (mant << 13));
#endif
}
LU_HDI uint16_t lu_f2h16(float f, int fmt) { return fmt ? lu_f2half(f) : lu_f2bf(f); }
LU_HDI float lu_h162f(uint16_t h, int fmt) { return fmt ? lu_half2f(h) : lu_bf2f(h); }

// ---- activation-tile staging tables ----------------------------------------------------------------------
// One "A stage" = one TMA box of activations (a [rows x pitch] pixel window x 64 channels) that is reused by
// `ntaps` consecutive K blocks; tap i of the stage reads the 128 output pixels' operand rows starting at row
// offset taps[tap_begin + i] inside the box (halo staging) -- or at offset 0 when every tap has its own box
// (direct staging).  K blocks (and with them the 64-wide column blocks of the packed weight matrix) are
// numbered in table order.
struct LuAStage {
  int32_t c;          // channel coordinate (dim 0) of the box
  int16_t dy, dx;     // box origin relative to the tile origin, in source pixel coordinates
  uint8_t src;        // which source view / tensor map
  uint8_t plane;      // coordinate in the extra dim (row parity of a stride-2 space-to-depth view)
  uint16_t ntaps;
  uint32_t tap_begin;
};

// 5-D view (c, w, p, h, n) of an NHWC bf16 buffer; element strides.  The same numbers feed
// cuTensorMapEncodeTiled and the scalar mirror kernel.
struct LuSrcView {
  const uint16_t* ptr;
  int64_t sn, sh, sp, sw;
  int32_t dimC, dimW, dimP, dimH, dimN;
  int32_t frame_mul, frame_add;    // n coordinate = tile_frame * frame_mul + frame_add
  int32_t rows, pitch;             // box = {64, pitch, 1, rows, 1}
};

enum { LU_EPI_CONV = 0, LU_EPI_LSTM = 1, LU_EPI_GRAD = 2 };

struct LuEpi {
  int32_t kind;
  int32_t H, W;               // tile-grid spatial size (rows/cols beyond it are masked)
  // output pixel = (y*oy_mul + oy_add, x*ox_mul + ox_add) in an OH x OW frame (identity except for the parity
  // classes of a stride-2 convolution's data gradient)
  int32_t oy_mul, oy_add, ox_mul, ox_add, OH, OW;
  int32_t accumulate;         // LU_EPI_GRAD: add to the existing contents of out_act
  int32_t fmt;                // 16-bit format of the activation buffers (0 bf16, 1 fp16)
  const float* bias;          // [Npad], packed column order
  int32_t out_frame_mul, out_frame_add;
  // conv
  const float* scale;         // folded BN scale/shift in packed order (NULL: no activation output)
  const float* shift;
  uint16_t* out_act;          // NHWC bf16, channel layout [hi: cpad][lo: cpad] when planes == 2
  int32_t out_cpad, out_planes;
  float* out_raw;             // NHWC fp32 (acc + bias), NULL if unused
  int32_t raw_cpad;
  double* bn_sums;            // training-mode BatchNorm: per-channel [sum | sum of squares] of (output - bias) over the
                              // valid pixels, accumulated by the epilogue itself (no separate pass over out_raw); NULL: off
  float alpha;                // LeakyReLU slope
  // lstm
  float* c_state;             // (B,H,W,f_pad) fp32, updated in place
  uint16_t* h_state_out;      // (B,H,W,planes*f_pad) bf16 or NULL: written on the last step of a call
  int32_t f_pad, ch_tile, gate_kind;
  uint16_t* save_gates;       // training: (frames,H,W,4*f_pad) bf16 post-activation i,f,g,o (NULL otherwise)
  float* save_c;              // training: (frames,H,W,f_pad) fp32 c_t
};

// Per-K-block weight packing descriptor (see lu_pack_weights).
struct LuPackDesc {
  int64_t w_off;              // offset of the Keras kernel tensor in the flat fp32 parameter buffer
  int32_t tap_off;            // (ky*k+kx) * cin_total * cout_total
  int32_t c_base;             // first input channel covered by this block
  int32_t n_valid;            // valid channels (NORMAL) / unused (PATCH)
  int32_t cin_total, cout_total;
  int8_t wpart;               // 0: hi part of w, 1: lo part (w - hi)
  int8_t kind;                // 0 NORMAL, 1 PATCH (block channel = tap index of a 1-channel image patch)
  int8_t k, pw;               // PATCH: conv kernel size, patch window size
  int8_t patch_x3;            // PATCH: channels [32,64) are the lo parts of the taps
  int8_t patch_hi_only;       // PATCH: channels [32,64) get zero weights (the a_hi * w_lo block)
  int8_t transposed;          // data-gradient operand: block channel indexes cout, packed column indexes cin
  int8_t pad1;
  int32_t col_base;           // transposed: first input channel (concat offset) of the packed columns
};

enum { LU_COL_IDENTITY = 0, LU_COL_LSTM = 1 };
struct LuColMap {
  int32_t kind, n_real;       // conv: packed column n < n_real maps to itself
  int32_t F, ch_tile;         // lstm: n = tile*4*ch + g*ch + j  ->  g*F + tile*ch + j  (if tile*ch + j < F)
};
LU_HDI int lu_col_of(const LuColMap& m, int n) {
  if (m.kind == LU_COL_IDENTITY) return n < m.n_real ? n : -1;
  int bn = 4 * m.ch_tile;
  int tile = n / bn, r = n % bn, g = r / m.ch_tile, j = r % m.ch_tile;
  int ch = tile * m.ch_tile + j;
  return ch < m.F ? g * m.F + ch : -1;
}

LU_HDI void lu_atomic_add(double* p, double v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
LU_HDI void lu_atomic_add(float* p, float v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
}

LU_HDI float lu_hard_sigmoid(float x) { return fminf(fmaxf(0.2f * x + 0.5f, 0.0f), 1.0f); }
LU_HDI float lu_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
