// Common device/host helpers for the DwiseNeuro B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

typedef __nv_bfloat16 bf16;

// ---- error plumbing (never throws across the C ABI) -------------------------------------------
int dwn_fail(const char* fmt, ...);
int dwn_num_sms();
#define DWN_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) return dwn_fail("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)
#define DWN_REQUIRE(cond, ...)                                                                    \
  do {                                                                                            \
    if (!(cond)) return dwn_fail(__VA_ARGS__);                                                    \
  } while (0)

#define DWN_DT_F32 0
#define DWN_DT_BF16 1

// ---- vector access: 16-byte vectors of the storage type ----------------------------------------
template <typename T> struct VecT;
template <> struct VecT<float> { static constexpr int V = 4; };
template <> struct VecT<bf16> { static constexpr int V = 8; };

__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  // PRMT + LOP3, both on the ALU pipe: `u << 16` is compiled to IMAD.SHL / IMAD.U32, which competes with the FFMA2s of
  // the stencil kernels for the FMA pipe (13 % of sdw_bwd's dynamic instructions were IMADs)
  lo = __uint_as_float(__byte_perm(u, 0u, 0x1044));
  hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 16-byte vector load -> floats
__device__ __forceinline__ void ldv(const float* p, float (&o)[4]) {
  float4 r = *reinterpret_cast<const float4*>(p);
  o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
__device__ __forceinline__ void ldv(const bf16* p, float (&o)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(r.x, o[0], o[1]);
  unpack_bf16x2(r.y, o[2], o[3]);
  unpack_bf16x2(r.z, o[4], o[5]);
  unpack_bf16x2(r.w, o[6], o[7]);
}
__device__ __forceinline__ void stv(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stv(bf16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}
// 4-channel access (quad) for both storage types
__device__ __forceinline__ void ldq(const float* p, float (&o)[4]) { ldv(p, o); }
__device__ __forceinline__ void ldq(const bf16* p, float (&o)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  unpack_bf16x2(r.x, o[0], o[1]);
  unpack_bf16x2(r.y, o[2], o[3]);
}
__device__ __forceinline__ void stq(float* p, const float (&v)[4]) { stv(p, v); }
__device__ __forceinline__ void stq(bf16* p, const float (&v)[4]) {
  uint2 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = r;
}
// value as it will be seen by the consumer after storage rounding
template <typename T> __device__ __forceinline__ float rnd(float x);
template <> __device__ __forceinline__ float rnd<float>(float x) { return x; }
template <> __device__ __forceinline__ float rnd<bf16>(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- activations ---------------------------------------------------------------------------------
// fp32 pipeline: exact expf / division.  bf16 pipeline: sigmoid(v) = 0.5 + 0.5*tanh(0.5 v) with the hardware
// tanh.approx.f32 (1 MUFU; abs error of SiLU <= 2.4e-4*|v|, below the bf16 rounding of the stored result).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> struct Act;
template <> struct Act<float> {
  static __device__ __forceinline__ float sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
};
template <> struct Act<bf16> {
  static __device__ __forceinline__ float sigmoid(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
};
template <typename T> __device__ __forceinline__ float silu_t(float x) { return x * Act<T>::sigmoid(x); }
// d/dx [x*sigmoid(x)] given x
template <typename T> __device__ __forceinline__ float silu_grad_t(float x) {
  float s = Act<T>::sigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}

// Fused BN-affine + SiLU with per-channel constants prepared once per thread:
//   fp32: (p0,p1) = (scale, shift)            y = v/(1+exp(-v)),      v = p0*x+p1
//   bf16: (p0,p1) = (scale/2, shift/2)        y = h + h*tanh(h),      h = p0*x+p1
template <typename T> struct BnSilu;
template <> struct BnSilu<float> {
  static __device__ __forceinline__ void prep(float sc, float sh, float& p0, float& p1) { p0 = sc; p1 = sh; }
  static __device__ __forceinline__ float act(float x, float p0, float p1) {
    const float v = fmaf(x, p0, p1);
    return v / (1.0f + expf(-v));
  }
  // returns activation, writes derivative d silu / d v
  static __device__ __forceinline__ float act_grad(float x, float p0, float p1, float& g) {
    const float v = fmaf(x, p0, p1);
    const float s = 1.0f / (1.0f + expf(-v));
    g = s * (1.0f + v * (1.0f - s));
    return v * s;
  }
};
template <> struct BnSilu<bf16> {
  static __device__ __forceinline__ void prep(float sc, float sh, float& p0, float& p1) { p0 = 0.5f * sc; p1 = 0.5f * sh; }
  static __device__ __forceinline__ float act(float x, float p0, float p1) {
    const float h = fmaf(x, p0, p1);
    return fmaf(h, tanh_approx(h), h);
  }
  static __device__ __forceinline__ float act_grad(float x, float p0, float p1, float& g) {
    const float h = fmaf(x, p0, p1);
    const float t = tanh_approx(h);
    const float sa = fmaf(h, t, h);          // v*sigmoid(v)
    const float w = fmaf(-0.5f, t, 0.5f);    // 1 - sigmoid(v)
    g = fmaf(sa, w, 1.0f - w);               // sigmoid + silu*(1-sigmoid)
    return sa;
  }
};

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100) --------------------------------------
typedef unsigned long long f32x2;  // two packed fp32 (low word = first element)
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void ffma2(f32x2& d, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void fadd2(f32x2& d, f32x2 a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a)); }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// BN-affine + SiLU on a channel pair; (p0,p1) prepared with BnSilu<T>::prep per channel and packed
template <typename T> __device__ __forceinline__ f32x2 bnsilu2(f32x2 x, f32x2 p0, f32x2 p1);
template <> __device__ __forceinline__ f32x2 bnsilu2<bf16>(f32x2 x, f32x2 p0, f32x2 p1) {
  f32x2 h = p1;
  ffma2(h, x, p0);
  float h0, h1;
  upk2(h, h0, h1);
  f32x2 y = h;
  ffma2(y, h, pk2(tanh_approx(h0), tanh_approx(h1)));
  return y;
}
template <> __device__ __forceinline__ f32x2 bnsilu2<float>(f32x2 x, f32x2 p0, f32x2 p1) {
  float x0, x1, a0, a1, b0, b1;
  upk2(x, x0, x1); upk2(p0, a0, a1); upk2(p1, b0, b1);
  return pk2(BnSilu<float>::act(x0, a0, b0), BnSilu<float>::act(x1, a1, b1));
}
// packed pair version of BnSilu<bf16>::act_grad: returns silu, writes d silu / d v
__device__ __forceinline__ f32x2 bnsilu_grad2_bf16(f32x2 x, f32x2 p0, f32x2 p1, f32x2& g) {
  f32x2 h = p1;
  ffma2(h, x, p0);
  float h0, h1;
  upk2(h, h0, h1);
  const f32x2 t = pk2(tanh_approx(h0), tanh_approx(h1));
  f32x2 sa = h;
  ffma2(sa, h, t);                                  // v*sigmoid(v)
  f32x2 w = pk2(0.5f, 0.5f);
  ffma2(w, t, pk2(-0.5f, -0.5f));                   // 1 - sigmoid(v)
  f32x2 one_m_w = pk2(1.0f, 1.0f);
  ffma2(one_m_w, w, pk2(-1.0f, -1.0f));             // sigmoid(v)
  g = one_m_w;
  ffma2(g, sa, w);                                  // sigmoid + silu*(1-sigmoid)
  return sa;
}
// 4 channels -> two packed pairs
__device__ __forceinline__ void ldq2(const float* p, f32x2 (&o)[2]) {
  const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p);
  o[0] = q.x; o[1] = q.y;
}
__device__ __forceinline__ void ldq2(const bf16* p, f32x2 (&o)[2]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  float a, b;
  unpack_bf16x2(r.x, a, b); o[0] = pk2(a, b);
  unpack_bf16x2(r.y, a, b); o[1] = pk2(a, b);
}
__device__ __forceinline__ void stq2(float* p, const f32x2 (&v)[2]) {
  ulonglong2 q; q.x = v[0]; q.y = v[1];
  *reinterpret_cast<ulonglong2*>(p) = q;
}
__device__ __forceinline__ void stq2(bf16* p, const f32x2 (&v)[2]) {
  float a, b, c, d;
  upk2(v[0], a, b); upk2(v[1], c, d);
  uint2 r; r.x = pack_bf16x2(a, b); r.y = pack_bf16x2(c, d);
  *reinterpret_cast<uint2*>(p) = r;
}

// one channel pair
__device__ __forceinline__ f32x2 ldp2(const float* p) { return *reinterpret_cast<const f32x2*>(p); }
__device__ __forceinline__ f32x2 ldp2(const bf16* p) {
  float a, b;
  unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p), a, b);
  return pk2(a, b);
}
__device__ __forceinline__ void stp2(float* p, f32x2 v) { *reinterpret_cast<f32x2*>(p) = v; }
__device__ __forceinline__ void stp2(bf16* p, f32x2 v) {
  float a, b;
  upk2(v, a, b);
  *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b);
}

// store + return the value as the consumer will see it (bf16: rounded; fp32: unchanged).  The BatchNorm-backward sums of a
// gradient must be taken over the ROUNDED values: the folded BN1 backward (dwn_pw_algebra.cu) subtracts mean terms from a
// GEMM over the stored tensor, and only the stored tensor's own sums make that cancellation exact
__device__ __forceinline__ f32x2 stp2_rnd(float* p, f32x2 v) { stp2(p, v); return v; }
__device__ __forceinline__ f32x2 stp2_rnd(bf16* p, f32x2 v) {
  float a, b;
  upk2(v, a, b);
  const uint32_t w = pack_bf16x2(a, b);
  *reinterpret_cast<uint32_t*>(p) = w;
  unpack_bf16x2(w, a, b);
  return pk2(a, b);
}
__device__ __forceinline__ void stq2_rnd(float* p, f32x2 (&v)[2]) { stq2(p, v); }
__device__ __forceinline__ void stq2_rnd(bf16* p, f32x2 (&v)[2]) {
  float a, b, c, d;
  upk2(v[0], a, b); upk2(v[1], c, d);
  uint2 r; r.x = pack_bf16x2(a, b); r.y = pack_bf16x2(c, d);
  *reinterpret_cast<uint2*>(p) = r;
  unpack_bf16x2(r.x, a, b); unpack_bf16x2(r.y, c, d);
  v[0] = pk2(a, b); v[1] = pk2(c, d);
}

// ---- reductions ----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums of two quantities of a partial table for the BN finalize kernels.  Block = 1024 threads =
// 8 channels x 128 row slices (grid = C/8 CTAs: enough CTAs and few enough rows per thread that the whole table is
// fetched in one or two rounds of independent loads).  Fixed summation order: deterministic.
// partial row p, quantity q, channel cp lives at partial[(p*NQ + q)*Cp + cp].  Returns the sums in threads 0..7
// (channel blockIdx.x*8 + threadIdx.x); sm must hold 2*32*8 doubles.
__device__ __forceinline__ void finalize_colsum2(const float* __restrict__ partial, int P, int NQ, int qa, int qb, int Cp,
                                                 int cp, bool valid, double* sm, double& out_a, double& out_b) {
  const int sl = threadIdx.x >> 3, cl = threadIdx.x & 7, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const long rs = (long)NQ * Cp;
    const float* pa = partial + (long)qa * Cp + cp;
    const float* pb = partial + (long)qb * Cp + cp;
    int p = sl;
    for (; p + 384 < P; p += 512) {  // four independent partial rows in flight per quantity
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a4[u] += pa[(long)(p + 128 * u) * rs];
        b4[u] += pb[(long)(p + 128 * u) * rs];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)  // tail: at most three rows are left
      if (p + 128 * u < P) {
        a4[u] += pa[(long)(p + 128 * u) * rs];
        b4[u] += pb[(long)(p + 128 * u) * rs];
      }
  }
  double a = ((double)a4[0] + (double)a4[1]) + ((double)a4[2] + (double)a4[3]);
  double b = ((double)b4[0] + (double)b4[1]) + ((double)b4[2] + (double)b4[3]);
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {  // the 4 slices of a warp that share a channel
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane < 8) { sm[warp * 8 + cl] = a; sm[256 + warp * 8 + cl] = b; }
  __syncthreads();
  out_a = 0; out_b = 0;
  if (threadIdx.x < 8)
    for (int w = 0; w < 32; ++w) { out_a += sm[w * 8 + threadIdx.x]; out_b += sm[256 + w * 8 + threadIdx.x]; }
}

// BN coefficient table layout: coef[4][C] = {scale, shift, mean, rstd}
//   y = scale*x + shift ;  xhat = (x-mean)*rstd
// BN backward coefficient table: bcoef[2][C] = {c1 = sum(dy)/N, c2 = sum(dy*xhat)/N}
//   dx = gamma*rstd*(dy - c1 - xhat*c2)

// division / modulo by a launch constant: a shift when the divisor is a power of two (all real shapes),
// a 32-bit division otherwise (row indices always fit in 31 bits)
struct FastDiv {
  int d, sh;
  __host__ __device__ FastDiv() : d(1), sh(0) {}
  __host__ explicit FastDiv(int dd) : d(dd), sh(-1) {
    for (int s = 0; s < 31; ++s)
      if ((1 << s) == dd) sh = s;
  }
  __device__ __forceinline__ int div(int x) const { return sh >= 0 ? (x >> sh) : (x / d); }
  __device__ __forceinline__ int mod(int x) const { return sh >= 0 ? (x & (d - 1)) : (x % d); }
};

// Index map of F.interpolate(mode="nearest", size=out) along one axis (interpolate_shortcut, dwiseneuro.py:125-129, where
// out = ceil(in / stride)): ATen's nearest_neighbor_compute_source_index with scale = (float)in / out, i.e.
// src(dst) = min(floorf(dst * scale), in - 1).  When in == out * s it is the integer map dst * s (every real shape:
// the reference pads its clips to 64 x 64); the float form covers sizes the stride does not divide, bit-exactly.
// For out <= in the map is strictly increasing, so it has an inverse on its image (dst(), used by the backward scatter
// and by the producers that take BatchNorm statistics over exactly the gathered positions).
struct NearestMap {
  int in, out, sh;  // sh >= 0: in == out << sh (the stride is a power of two and divides the size): shifts only
  float scale;
  __host__ __device__ NearestMap() : in(1), out(1), sh(0), scale(1.f) {}
  __host__ NearestMap(int in_, int out_) : in(in_), out(out_), sh(-1), scale((float)in_ / (float)out_) {
    for (int k = 0; k < 8; ++k)
      if (in_ == (out_ << k)) sh = k;
  }
  __device__ __forceinline__ int src(int d) const {
    if (sh >= 0) return d << sh;
    const int v = (int)floorf((float)d * scale);
    return v < in - 1 ? v : in - 1;
  }
  __device__ __forceinline__ int dst(int x) const {  // -1: position x is not gathered
    if (sh >= 0) return (x & ((1 << sh) - 1)) == 0 ? (x >> sh) : -1;
    const int d0 = (int)((float)x / scale);
    for (int d = (d0 > 0 ? d0 - 1 : 0); d <= d0 + 1 && d < out; ++d) {
      const int v = (int)floorf((float)d * scale);
      if ((v < in - 1 ? v : in - 1) == x) return d;
    }
    return -1;
  }
};
static inline int dwn_ceil_div(int a, int b) { return (a + b - 1) / b; }

static inline int dwn_largest_divisor_le(int n, int cap) {
  int best = 1;
  for (int d = 1; d <= cap && d <= n; ++d)
    if (n % d == 0) best = d;
  return best;
}
