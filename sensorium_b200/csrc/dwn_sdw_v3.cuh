// Pipelined bf16 spatial depth-wise kernels (forward + backward), sm_100a.
//   * raw bf16 tiles are streamed global -> shared with cp.async (no register staging) and double-buffered, so the
//     DRAM latency of tile k+1 is hidden behind the BN+SiLU pass and the stencil of tile k;
//   * the stencil runs on packed fp32x2 FMAs (FFMA2), two channels per instruction;
//   * grid is 1-D with the channel chunk fastest, so CTAs that share a (plane, band) tile are co-resident and each
//     128-byte line of E_raw is fetched from HBM once.
#pragma once
#include "dwn_common.cuh"
#include "dwn_reduce.cuh"
#include <type_traits>

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
// small always-valid copies (L1-allocating variant: .cg only exists for 16 bytes)
__device__ __forceinline__ void cp_async16_ok(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
// Per-thread software pipeline through shared memory (trunk kernels): every thread owns DEPTH private slots of NV
// 16-byte vectors and issues its own cp.async copies DEPTH loop iterations ahead, so the bytes in flight are
// DEPTH x slot x threads instead of what fits in registers.  Layout [DEPTH][NV][nthr]: conflict-free 16-byte accesses.
template <int DEPTH, int NV>
struct ThreadPipe {
  static_assert((DEPTH & (DEPTH - 1)) == 0, "DEPTH must be a power of two");
  uint4* base;
  int nthr;
  __device__ __forceinline__ ThreadPipe(void* smem, int nthr_, int tid) : base(reinterpret_cast<uint4*>(smem) + tid), nthr(nthr_) {}
  __device__ __forceinline__ uint4* slot(int k, int v) const { return base + ((k & (DEPTH - 1)) * NV + v) * nthr; }
  static size_t bytes(int nthr) { return (size_t)DEPTH * NV * nthr * sizeof(uint4); }
};
__device__ __forceinline__ void quad_from(const uint4& q, float (&o)[4]) {
  o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
}
__device__ __forceinline__ void quad_from_bf16(const uint4& q, float (&o)[4]) {  // first 8 bytes hold 4 bf16
  unpack_bf16x2(q.x, o[0], o[1]);
  unpack_bf16x2(q.y, o[2], o[3]);
}
template <typename T> __device__ __forceinline__ void pipe_issue_quad(uint4* dst, const T* src);
template <> __device__ __forceinline__ void pipe_issue_quad<float>(uint4* dst, const float* src) { cp_async16_ok(dst, src); }
template <> __device__ __forceinline__ void pipe_issue_quad<bf16>(uint4* dst, const bf16* src) { cp_async8(dst, src); }
template <typename T> __device__ __forceinline__ void pipe_read_quad(const uint4* src, float (&o)[4]);
template <> __device__ __forceinline__ void pipe_read_quad<float>(const uint4* src, float (&o)[4]) { quad_from(*src, o); }
template <> __device__ __forceinline__ void pipe_read_quad<bf16>(const uint4* src, float (&o)[4]) {
  const uint2 q = *reinterpret_cast<const uint2*>(src);
  unpack_bf16x2(q.x, o[0], o[1]);
  unpack_bf16x2(q.y, o[2], o[3]);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// =================================================================================================
// forward:  S_raw = dwconv3x3_stride_S( SiLU(BN1(E_raw)) ),  + per-channel sum / sumsq partials
//   RPI = tile rows covered by one pass of the 256 threads over the raw tile (256 / (W * CC/8), 1 or 2).
//   All per-vector addresses are loop-invariant per thread plus a compile-time multiple of a per-row step.
// =================================================================================================
template <int S, int THO, int RPI, int CC>
__global__ void __launch_bounds__(256, 2)
sdw_fwd_v3_kernel(const bf16* __restrict__ in, const float* __restrict__ coef, const float* __restrict__ wgt,
                  bf16* __restrict__ out, float* __restrict__ partial, int NP, int H, int C, int nchunks, int nbsh) {
  // CC (channels per CTA) is a template parameter: with (CC/4)*Wo == 256 the tile geometry W, Wo, WP and every
  // shared-memory offset become compile-time constants (immediate LDS/STS offsets, no per-tap integer adds)
  constexpr int Wo = 1024 / CC, W = Wo * S, WP = W + 2;
  constexpr int cvn = CC / 8, cvsh = (cvn == 2 ? 1 : cvn == 4 ? 2 : cvn == 8 ? 3 : 4);
  constexpr int NR = (THO - 1) * S + 3;
  constexpr int NIT = NR / RPI;
  static_assert(NR % RPI == 0, "tile rows must be a multiple of the rows per pass");
  static_assert(W * cvn == 256 / RPI, "RPI must match the tile geometry");
  extern __shared__ __align__(16) unsigned char smem_v3[];
  const int Ho = H / S;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  constexpr int NVEC = NR * (256 / RPI);  // 16-byte vectors per raw tile (W*cvn == 256/RPI)
  // three raw tiles in a ring: while tile k is activated and convolved, tiles k+1 and k+2 are in flight
  // (two buffers kept only ~40 KB in flight per SM, below what the HBM latency needs)
  bf16* raw0 = reinterpret_cast<bf16*>(smem_v3);
  // activated halo tile in bf16 (what torch's autocast feeds the depth-wise conv: SiLU output is bf16 there too).  An
  // fp32 tile kept the shared-memory pipe at 82 % of its wavefront peak (ncu, round 2): STS and LDS bytes are halved
  bf16* act = raw0 + (size_t)3 * NVEC * 8;
  // ---- loop-invariant coordinates of this thread inside one pass
  constexpr int vpr = 256 / RPI;                   // vectors per tile row
  const int r_first = tid / vpr;                   // 0 (RPI=1) or 0/1 (RPI=2)
  const int wq = (tid & (vpr - 1)) >> cvsh;
  const int lcv = tid & (cvn - 1);
  const int act_off0 = (r_first * WP + wq + 1) * CC + lcv * 8;
  constexpr int act_step = RPI * WP * CC;
  const long g_off0 = ((long)r_first * W + wq) * C + c0 + lcv * 8;
  const long g_step = (long)RPI * W * C;
  f32x2 lp0[4], lp1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a0, a1, b0, b1;
    BnSilu<bf16>::prep(coef[c0 + lcv * 8 + 2 * j], coef[C + c0 + lcv * 8 + 2 * j], a0, b0);
    BnSilu<bf16>::prep(coef[c0 + lcv * 8 + 2 * j + 1], coef[C + c0 + lcv * 8 + 2 * j + 1], a1, b1);
    lp0[j] = pk2(a0, a1);
    lp1[j] = pk2(b0, b1);
  }
  constexpr int cqn = CC >> 2;
  const int cq = tid % cqn, wo = tid / cqn;
  const bf16* act_rd = act + (wo * S) * CC + cq * 4;
  constexpr int row_step = WP * CC;
  f32x2 w2[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w2[k][0] = pk2(wgt[(c0 + cq * 4 + 0) * 9 + k], wgt[(c0 + cq * 4 + 1) * 9 + k]);
    w2[k][1] = pk2(wgt[(c0 + cq * 4 + 2) * 9 + k], wgt[(c0 + cq * 4 + 3) * 9 + k]);
  }
  f32x2 st2[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
  for (int i = tid; i < NR * 2 * CC; i += 256) {  // zero halo columns once
    const int r = i / (2 * CC), rem = i % (2 * CC);
    act[(r * WP + ((rem / CC) ? (W + 1) : 0)) * CC + (rem % CC)] = __float2bfloat16_rn(0.f);
  }
  const int nbm = (1 << nbsh) - 1;
  const int ntiles = NP << nbsh;
  const long plane = (long)H * W * C;
  auto issue = [&](int t, bf16* buf) {
    const int p = t >> nbsh, hi0 = (t & nbm) * (THO * S) - 1;
    const bf16* src = in + (long)p * plane + (long)hi0 * W * C + g_off0;
    bf16* dst = buf + tid * 8;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int hi = hi0 + r_first + it * RPI;
      const bool ok = (unsigned)hi < (unsigned)H;
      cp_async16(dst + it * 2048, ok ? src + it * g_step : in, ok);
    }
  };
  int t = worker, k = 0;
  if (t < ntiles) issue(t, raw0);
  cp_async_commit();
  if (t + nworkers < ntiles) issue(t + nworkers, raw0 + (size_t)NVEC * 8);
  cp_async_commit();
  for (; t < ntiles; t += nworkers, k = (k == 2 ? 0 : k + 1)) {
    const bf16* cur = raw0 + (size_t)k * NVEC * 8;
    bf16* nxt = raw0 + (size_t)(k == 0 ? 2 : k - 1) * NVEC * 8;  // buffer (k+2)%3: its tile was consumed last iteration
    if (t + 2 * nworkers < ntiles) issue(t + 2 * nworkers, nxt);
    cp_async_commit();
    cp_async_wait<2>();
    __syncthreads();  // raw tile landed for everybody; everybody is done with the previous act tile
    const int p = t >> nbsh, ho0 = (t & nbm) * THO;
    const int hi0 = ho0 * S - 1;
    // ---- BN1 + SiLU pass: raw bf16 -> act fp32 (zero rows outside the image: padding applies after the activation)
    {
      const bf16* rp = cur + tid * 8;
      bf16* dst = act + act_off0;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int hi = hi0 + r_first + it * RPI;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)hi < (unsigned)H) {
          const uint4 q = *reinterpret_cast<const uint4*>(rp + it * 2048);
          const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
          uint32_t ow[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo, hi2;
            unpack_bf16x2(qq[j], lo, hi2);
            f32x2 h = lp1[j];
            ffma2(h, pk2(lo, hi2), lp0[j]);        // h = x*p0 + p1
            float h0, h1;
            upk2(h, h0, h1);
            f32x2 y = h;
            ffma2(y, h, pk2(tanh_approx(h0), tanh_approx(h1)));  // y = h + h*tanh(h)
            float y0, y1;
            upk2(y, y0, y1);
            ow[j] = pack_bf16x2(y0, y1);
          }
          o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        *reinterpret_cast<uint4*>(dst + it * act_step) = o;   // one 16-byte store per thread, contiguous across lanes
      }
    }
    __syncthreads();
    // ---- stencil: sliding 3-row register window, packed fp32x2 FMAs
    f32x2 R[3][3][2];
    auto load_row = [&](int r) {
      const bf16* src = act_rd + r * row_step;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint2 q = *reinterpret_cast<const uint2*>(src + kw * CC);
        float a0, a1, a2, a3;
        unpack_bf16x2(q.x, a0, a1);
        unpack_bf16x2(q.y, a2, a3);
        R[r % 3][kw][0] = pk2(a0, a1);
        R[r % 3][kw][1] = pk2(a2, a3);
      }
    };
    bf16* op = out + (((long)p * Ho + ho0) * Wo + wo) * C + c0 + cq * 4;
    const long ostep = (long)Wo * C;
    if (S == 1) { load_row(0); load_row(1); } else { load_row(0); }
#pragma unroll
    for (int hl = 0; hl < THO; ++hl) {
      if (S == 1) { load_row(hl + 2); } else { load_row(2 * hl + 1); load_row(2 * hl + 2); }
      f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          ffma2(a0, R[(hl * S + kh) % 3][kw][0], w2[kh * 3 + kw][0]);
          ffma2(a1, R[(hl * S + kh) % 3][kw][1], w2[kh * 3 + kw][1]);
        }
      float o[4];
      upk2(a0, o[0], o[1]);
      upk2(a1, o[2], o[3]);
      stq(op + hl * ostep, o);
      fadd2(st2[0][0], a0);
      fadd2(st2[0][1], a1);
      ffma2(st2[1][0], a0, a0);
      ffma2(st2[1][1], a1, a1);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (partial) {
    float st[2][4];
    upk2(st2[0][0], st[0][0], st[0][1]); upk2(st2[0][1], st[0][2], st[0][3]);
    upk2(st2[1][0], st[1][0], st[1][1]); upk2(st2[1][1], st[1][2], st[1][3]);
    block_reduce_channels<2, 4>(st, reinterpret_cast<float*>(smem_v3), cqn, Wo, partial + (long)worker * 2 * C, C, c0);
  }
}

// =================================================================================================
// backward: dE_pre = SiLU'(BN1 E) * dwconv3x3^T( BN2bwd(dS_pre) ), dw[9], BN1-backward partial sums
//   partial[P][11][C] = { sum dehat, sum dehat*xhat1, dw[0..8] }
// The staged (output-sized) tile is Wo*CC/8 = 128/S vectors per row, i.e. RPI = 2*S rows per 256-thread pass.
// =================================================================================================
template <int S, int THI, int CC>
__global__ void __launch_bounds__(256, 2)
sdw_bwd_v3_kernel(const bf16* __restrict__ dsh, const bf16* __restrict__ s_raw, const bf16* __restrict__ e_raw,
                  const float* __restrict__ coef2, const float* __restrict__ bcoef2, const float* __restrict__ coef1,
                  const float* __restrict__ wgt, bf16* __restrict__ dE, float* __restrict__ partial, int NP, int H,
                  int C, int nchunks, int nbsh) {
  // CC is a template parameter: (CC/4)*W == 256 makes W, Wo, WP and all smem offsets compile-time constants
  constexpr int W = 1024 / CC, Wo = W / S, WP = Wo + 2;
  constexpr int cvn = CC / 8, cvsh = (cvn == 2 ? 1 : cvn == 4 ? 2 : cvn == 8 ? 3 : 4);
  constexpr int NR = S == 1 ? THI + 2 : THI / 2 + 1;
  constexpr int RPI = 2 * S;
  constexpr int NIT = (NR + RPI - 1) / RPI;
  constexpr int VPR = 256 / RPI;  // vectors per staged row
  extern __shared__ __align__(16) unsigned char smem_v3[];
  const int Ho = H / S;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  constexpr int NVEC = NIT * 256;
  constexpr int EVEC = THI * 128;  // E tile: THI rows x (W*CC/8 = 128) 16-byte vectors
  bf16* rawD = reinterpret_cast<bf16*>(smem_v3);
  bf16* rawS = rawD + (size_t)NVEC * 8;
  bf16* rawE = rawS + (size_t)NVEC * 8;  // [2][EVEC*8]: the E tile of item k+1 streams in while item k is convolved
  float* tile = reinterpret_cast<float*>(rawE + (size_t)2 * EVEC * 8);
  float* sco = tile + (size_t)NR * WP * CC;
  const int r_first = tid / VPR;
  const int two = (tid & (VPR - 1)) >> cvsh;  // output column handled in the staging passes
  const int lcv = tid & (cvn - 1);
  constexpr int cqn = CC >> 2;
  // stride 2: the first 128 threads take the even input columns, the last 128 the odd ones, so the column parity
  // (which decides the taps: kw = 1 for even, kw = 0 and 2 for odd columns) is uniform per warp and is a branch
  // instead of per-tap selects
  const int cq = tid % cqn, widx = tid / cqn;
  const int wi = (S == 1) ? widx : ((widx % (W / 2)) * 2 + widx / (W / 2));
  const int cch = c0 + cq * 4;
  for (int i = tid; i < CC; i += 256) {
    const int cc = c0 + i;
    const float sc = coef2[cc], mu = coef2[2 * C + cc], rs = coef2[3 * C + cc];
    const float k1 = bcoef2[cc], k2 = bcoef2[C + cc];
    sco[i] = sc;
    sco[CC + i] = -sc * (k1 - mu * rs * k2);
    sco[2 * CC + i] = -sc * rs * k2;
    float q0, q1;
    BnSilu<bf16>::prep(coef1[cc], coef1[C + cc], q0, q1);
    sco[3 * CC + i] = q0;
    sco[4 * CC + i] = q1;
    sco[5 * CC + i] = coef1[2 * C + cc];
    sco[6 * CC + i] = coef1[3 * C + cc];
  }
  for (int i = tid; i < NR * WP * CC; i += 256) tile[i] = 0.f;  // halo columns stay zero
  f32x2 w2[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w2[k][0] = pk2(wgt[(cch + 0) * 9 + k], wgt[(cch + 1) * 9 + k]);
    w2[k][1] = pk2(wgt[(cch + 2) * 9 + k], wgt[(cch + 3) * 9 + k]);
  }
  const bool odd_w = (wi & 1) != 0;
  const int colA = (S == 1) ? 0 : (odd_w ? (wi + 1) / 2 : wi / 2);
  const int colB = (wi - 1) / 2;
  f32x2 st2[11][2];
#pragma unroll
  for (int q = 0; q < 11; ++q) { st2[q][0] = 0ull; st2[q][1] = 0ull; }
  const int nbm = (1 << nbsh) - 1;
  const int ntiles = NP << nbsh;
  constexpr int row_step = WP * CC;
  const int tile_off0 = (r_first * WP + two + (S == 1 ? 1 : 0)) * CC + lcv * 8;
  const int h0 = (tid & 4) ? 4 : 0;  // conflict-free order of the two 16-byte STS (see sdw_fwd_v3_kernel)
  const long g_off0 = (long)two * C + c0 + lcv * 8;
  const long orow = (long)Wo * C;
  const long erow = (long)W * C;
  auto issue = [&](int t) {
    const int p = t >> nbsh, hi0 = (t & nbm) * THI;
    const int ho_first = (S == 1) ? hi0 - 1 : hi0 / 2;
    const long base = (long)p * Ho * orow + g_off0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int r = r_first + it * RPI;
      const int ho = ho_first + r;
      const bool ok = (r < NR) && ((unsigned)ho < (unsigned)Ho);
      const long off = ok ? base + (long)ho * orow : 0;
      cp_async16(rawD + (size_t)(tid + it * 256) * 8, dsh + off, ok);
      cp_async16(rawS + (size_t)(tid + it * 256) * 8, s_raw + off, ok);
    }
  };
  // E tile staging: vector i = tid + it*256 -> row 2*it + (tid>>7), column vector tid&127
  const int e_r0 = tid >> 7;
  const long e_goff = (long)((tid & 127) >> cvsh) * C + c0 + lcv * 8;
  auto issue_e = [&](int t, int buf) {
    const int p = t >> nbsh, hi0 = (t & nbm) * THI;
    const bf16* eb = e_raw + ((long)p * H + hi0 + e_r0) * erow + e_goff;
    bf16* dst = rawE + (size_t)buf * EVEC * 8;
#pragma unroll
    for (int it = 0; it < THI / 2; ++it) cp_async16(dst + (size_t)(tid + it * 256) * 8, eb + (long)(2 * it) * erow, true);
  };
  int t = worker, kbuf = 0;
  if (t < ntiles) issue(t);
  cp_async_commit();
  if (t < ntiles) issue_e(t, 0);
  cp_async_commit();
  for (; t < ntiles; t += nworkers, kbuf ^= 1) {
    const int p = t >> nbsh, hi0 = (t & nbm) * THI;
    const int ho_first = (S == 1) ? hi0 - 1 : hi0 / 2;
    bf16* dp = dE + (((long)p * H + hi0) * W + wi) * C + cch;
    cp_async_wait<1>();  // the raw dS/S tiles of this item (the E tile, committed after them, may still be in flight)
    __syncthreads();     // raw tiles visible; previous stencil finished with `tile` and with the other E buffer
    // ---- BN2 backward on the staged tile: dS_raw = a*g - d*x - b (zero outside the image)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int r = r_first + it * RPI;
      if (r < NR) {
        const int ho = ho_first + r;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        if ((unsigned)ho < (unsigned)Ho) {
          const uint4 qg = *reinterpret_cast<const uint4*>(rawD + (size_t)(tid + it * 256) * 8);
          const uint4 qx = *reinterpret_cast<const uint4*>(rawS + (size_t)(tid + it * 256) * 8);
          const uint32_t gg[4] = {qg.x, qg.y, qg.z, qg.w}, xx[4] = {qx.x, qx.y, qx.z, qx.w};
          const f32x2* ca = reinterpret_cast<const f32x2*>(sco + lcv * 8);            // a
          const f32x2* cb = reinterpret_cast<const f32x2*>(sco + CC + lcv * 8);       // -b
          const f32x2* cd = reinterpret_cast<const f32x2*>(sco + 2 * CC + lcv * 8);   // -d
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g0, g1, x0, x1;
            unpack_bf16x2(gg[j], g0, g1);
            unpack_bf16x2(xx[j], x0, x1);
            f32x2 r = cb[j];                    // dS_raw = a*g - d*x - b
            ffma2(r, ca[j], pk2(g0, g1));
            ffma2(r, cd[j], pk2(x0, x1));
            upk2(r, v[2 * j], v[2 * j + 1]);
          }
          o0 = make_float4(v[0], v[1], v[2], v[3]);
          o1 = make_float4(v[4], v[5], v[6], v[7]);
        }
        float* dst = tile + tile_off0 + it * RPI * row_step;
        *reinterpret_cast<float4*>(dst + h0) = h0 ? o1 : o0;
        *reinterpret_cast<float4*>(dst + (4 - h0)) = h0 ? o0 : o1;
      }
    }
    cp_async_wait<0>();  // E tile of this item landed (it was issued one item ago)
    __syncthreads();     // tile + E ready, raw dS/S buffers free
    if (t + nworkers < ntiles) issue(t + nworkers);
    cp_async_commit();
    if (t + nworkers < ntiles) issue_e(t + nworkers, kbuf ^ 1);
    cp_async_commit();
    const bf16* rawEk = rawE + (size_t)kbuf * EVEC * 8;
    const bf16* esm = rawEk + (size_t)wi * CC + cq * 4;  // row hl at + hl*W*CC
    // ---- transposed stencil + weight gradient, one input row at a time (E row prefetched one ahead)
    const ulonglong2 qa0 = *reinterpret_cast<const ulonglong2*>(sco + 3 * CC + cq * 4);
    const ulonglong2 qa1 = *reinterpret_cast<const ulonglong2*>(sco + 4 * CC + cq * 4);
    const float* tbase = tile + cq * 4 + ((S == 1) ? (wi + 2) * CC : 0);
    auto row_body = [&](const int hl, auto par_c) {
      constexpr int PAR = decltype(par_c)::value;
      f32x2 e2[2], sg0, sg1;
      ldq2(esm + hl * (W * CC), e2);
      const f32x2 ea0 = bnsilu_grad2_bf16(e2[0], qa0.x, qa1.x, sg0);
      const f32x2 ea1 = bnsilu_grad2_bf16(e2[1], qa0.y, qa1.y, sg1);
      f32x2 acc0 = 0ull, acc1 = 0ull;
      if (S == 1) {
        const float* trow = tbase + hl * row_step;  // tap (kh,kw) sits at trow + (2-kh)*row_step - kw*CC
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float* tr = trow + (2 - kh) * row_step;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(tr - kw * CC);
            ffma2(acc0, w2[kh * 3 + kw][0], q.x);
            ffma2(acc1, w2[kh * 3 + kw][1], q.y);
            ffma2(st2[2 + kh * 3 + kw][0], ea0, q.x);
            ffma2(st2[2 + kh * 3 + kw][1], ea1, q.y);
          }
        }
      } else {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if ((PAR == 0) != (kh == 1)) continue;
          const int r = (kh == 1) ? hl / 2 : (kh == 0 ? (hl + 1) / 2 : (hl - 1) / 2);
          const float* tr = tbase + r * row_step;
          if (odd_w) {  // warp-uniform
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(tr + colA * CC);
            ffma2(acc0, w2[kh * 3 + 0][0], q.x);
            ffma2(acc1, w2[kh * 3 + 0][1], q.y);
            ffma2(st2[2 + kh * 3 + 0][0], ea0, q.x);
            ffma2(st2[2 + kh * 3 + 0][1], ea1, q.y);
            const ulonglong2 q2 = *reinterpret_cast<const ulonglong2*>(tr + colB * CC);
            ffma2(acc0, w2[kh * 3 + 2][0], q2.x);
            ffma2(acc1, w2[kh * 3 + 2][1], q2.y);
            ffma2(st2[2 + kh * 3 + 2][0], ea0, q2.x);
            ffma2(st2[2 + kh * 3 + 2][1], ea1, q2.y);
          } else {
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(tr + colA * CC);
            ffma2(acc0, w2[kh * 3 + 1][0], q.x);
            ffma2(acc1, w2[kh * 3 + 1][1], q.y);
            ffma2(st2[2 + kh * 3 + 1][0], ea0, q.x);
            ffma2(st2[2 + kh * 3 + 1][1], ea1, q.y);
          }
        }
      }
      f32x2 o2[2];
      o2[0] = fmul2(acc0, sg0);
      o2[1] = fmul2(acc1, sg1);
      stq2_rnd(dp + hl * erow, o2);  // o2 <- the stored (rounded) values
      // statistics on the raw E: sum(o) and sum(o*e); sum(o*xhat) = rstd*(sum(o*e) - mean*sum(o)) is formed once
      // per CTA after the tile loop
      fadd2(st2[0][0], o2[0]);
      fadd2(st2[0][1], o2[1]);
      ffma2(st2[1][0], o2[0], e2[0]);
      ffma2(st2[1][1], o2[1], e2[1]);
    };
    if (S == 1) {
#pragma unroll 1
      for (int hl = 0; hl < THI; ++hl) row_body(hl, std::integral_constant<int, 0>{});
    } else {
#pragma unroll 1
      for (int hl = 0; hl < THI; hl += 2) {
        row_body(hl, std::integral_constant<int, 0>{});
        row_body(hl + 1, std::integral_constant<int, 1>{});
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  float st[11][4];
#pragma unroll
  for (int q = 0; q < 11; ++q) { upk2(st2[q][0], st[q][0], st[q][1]); upk2(st2[q][1], st[q][2], st[q][3]); }
#pragma unroll
  for (int j = 0; j < 4; ++j)  // sum(o*xhat1) from the raw-E sums
    st[1][j] = coef1[3 * C + cch + j] * (st[1][j] - coef1[2 * C + cch + j] * st[0][j]);
  block_reduce_channels<11, 4>(st, reinterpret_cast<float*>(smem_v3), cqn, W, partial + (long)worker * 11 * C, C, c0);
}

// =================================================================================================
// backward, stride 1 (v5).  Same staging as sdw_bwd_v3_kernel<1,...>, different stencil mapping: the v3 stencil
// re-read all nine fp32 taps of every output from shared memory (9 x LDS.128 per 4 channels) and ncu showed the
// stride-1 launches limited by shared-memory load wavefronts (~80 % of peak, short-scoreboard stalls) at 47 % of HBM.
// Here a thread owns ONE channel pair and two columns (wcol, wcol + W/2) in turn and keeps a sliding 3x3 register
// window over the tile rows, so each output costs 3 x LDS.64: a third of the shared-memory traffic per channel.
// =================================================================================================
template <int THI, int CC>
__global__ void __launch_bounds__(256, 2)
sdw_bwd_v5_kernel(const bf16* __restrict__ dsh, const bf16* __restrict__ s_raw, const bf16* __restrict__ e_raw,
                  const float* __restrict__ coef2, const float* __restrict__ bcoef2, const float* __restrict__ coef1,
                  const float* __restrict__ wgt, bf16* __restrict__ dE, float* __restrict__ partial, int NP, int H,
                  int C, int nchunks, int nbsh) {
  constexpr int W = 1024 / CC, WP = W + 2;
  constexpr int cvn = CC / 8, cvsh = (cvn == 2 ? 1 : cvn == 4 ? 2 : cvn == 8 ? 3 : 4);
  constexpr int NR = THI + 2;
  constexpr int RPI = 2;
  constexpr int NIT = (NR + RPI - 1) / RPI;
  constexpr int VPR = 256 / RPI;  // vectors per staged row
  extern __shared__ __align__(16) unsigned char smem_v3[];
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  constexpr int NVEC = NIT * 256;
  constexpr int EVEC = THI * 128;  // E tile: THI rows x (W*CC/8 = 128) 16-byte vectors
  bf16* rawD = reinterpret_cast<bf16*>(smem_v3);
  bf16* rawS = rawD + (size_t)NVEC * 8;
  bf16* rawE = rawS + (size_t)NVEC * 8;
  float* tile = reinterpret_cast<float*>(rawE + (size_t)EVEC * 8);
  float* sco = tile + (size_t)NR * WP * CC;
  // staging coordinates (16-byte vectors of 8 channels)
  const int r_first = tid / VPR;
  const int two = (tid & (VPR - 1)) >> cvsh;
  const int lcv = tid & (cvn - 1);
  // stencil coordinates (one channel pair, two columns)
  constexpr int cpn = CC / 2, NCOL = 256 / cpn;
  static_assert(NCOL * 2 == W, "two column halves per thread");
  const int cp = tid % cpn, wcol = tid / cpn;
  const int cch = c0 + cp * 2;
  for (int i = tid; i < CC; i += 256) {
    const int cc = c0 + i;
    const float sc = coef2[cc], mu = coef2[2 * C + cc], rs = coef2[3 * C + cc];
    const float k1 = bcoef2[cc], k2 = bcoef2[C + cc];
    sco[i] = sc;
    sco[CC + i] = -sc * (k1 - mu * rs * k2);
    sco[2 * CC + i] = -sc * rs * k2;
    float q0, q1;
    BnSilu<bf16>::prep(coef1[cc], coef1[C + cc], q0, q1);
    sco[3 * CC + i] = q0;
    sco[4 * CC + i] = q1;
  }
  for (int i = tid; i < NR * WP * CC; i += 256) tile[i] = 0.f;  // halo columns stay zero
  f32x2 w2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w2[k] = pk2(wgt[cch * 9 + k], wgt[(cch + 1) * 9 + k]);
  f32x2 st2[11];
#pragma unroll
  for (int q = 0; q < 11; ++q) st2[q] = 0ull;
  const int nbm = (1 << nbsh) - 1;
  const int ntiles = NP << nbsh;
  constexpr int row_step = WP * CC;
  const int tile_off0 = (r_first * WP + two + 1) * CC + lcv * 8;
  const int h0 = (tid & 4) ? 4 : 0;  // conflict-free order of the two 16-byte STS
  const long g_off0 = (long)two * C + c0 + lcv * 8;
  const long orow = (long)W * C;
  auto issue = [&](int t) {
    const int p = t >> nbsh, hi0 = (t & nbm) * THI;
    const long base = (long)p * H * orow + g_off0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int r = r_first + it * RPI;
      const int ho = hi0 - 1 + r;
      const bool ok = (r < NR) && ((unsigned)ho < (unsigned)H);
      const long off = ok ? base + (long)ho * orow : 0;
      cp_async16(rawD + (size_t)(tid + it * 256) * 8, dsh + off, ok);
      cp_async16(rawS + (size_t)(tid + it * 256) * 8, s_raw + off, ok);
    }
  };
  const int e_r0 = tid >> 7;
  const long e_goff = (long)((tid & 127) >> cvsh) * C + c0 + lcv * 8;
  int t = worker;
  if (t < ntiles) issue(t);
  cp_async_commit();
  for (; t < ntiles; t += nworkers) {
    const int p = t >> nbsh, hi0 = (t & nbm) * THI;
    cp_async_wait<0>();
    __syncthreads();  // raw dS/S tiles landed; previous stencil finished with `tile` and `rawE`
    {  // E rows of this tile: in flight while the dS tile is transformed
      const bf16* eb = e_raw + ((long)p * H + hi0 + e_r0) * orow + e_goff;
#pragma unroll
      for (int it = 0; it < THI / 2; ++it) cp_async16(rawE + (size_t)(tid + it * 256) * 8, eb + (long)(2 * it) * orow, true);
      cp_async_commit();
    }
    // ---- BN2 backward on the staged tile: dS_raw = a*g - d*x - b (zero outside the image)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int r = r_first + it * RPI;
      if (r < NR) {
        const int ho = hi0 - 1 + r;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        if ((unsigned)ho < (unsigned)H) {
          const uint4 qg = *reinterpret_cast<const uint4*>(rawD + (size_t)(tid + it * 256) * 8);
          const uint4 qx = *reinterpret_cast<const uint4*>(rawS + (size_t)(tid + it * 256) * 8);
          const uint32_t gg[4] = {qg.x, qg.y, qg.z, qg.w}, xx[4] = {qx.x, qx.y, qx.z, qx.w};
          const f32x2* ca = reinterpret_cast<const f32x2*>(sco + lcv * 8);            // a
          const f32x2* cb = reinterpret_cast<const f32x2*>(sco + CC + lcv * 8);       // -b
          const f32x2* cd = reinterpret_cast<const f32x2*>(sco + 2 * CC + lcv * 8);   // -d
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g0, g1, x0, x1;
            unpack_bf16x2(gg[j], g0, g1);
            unpack_bf16x2(xx[j], x0, x1);
            f32x2 rr = cb[j];
            ffma2(rr, ca[j], pk2(g0, g1));
            ffma2(rr, cd[j], pk2(x0, x1));
            upk2(rr, v[2 * j], v[2 * j + 1]);
          }
          o0 = make_float4(v[0], v[1], v[2], v[3]);
          o1 = make_float4(v[4], v[5], v[6], v[7]);
        }
        float* dst = tile + tile_off0 + it * RPI * row_step;
        *reinterpret_cast<float4*>(dst + h0) = h0 ? o1 : o0;
        *reinterpret_cast<float4*>(dst + (4 - h0)) = h0 ? o0 : o1;
      }
    }
    cp_async_wait<0>();  // E tile landed
    __syncthreads();     // tile + E ready, raw dS/S buffers free
    if (t + nworkers < ntiles) issue(t + nworkers);
    cp_async_commit();
    // ---- transposed stencil + weight gradient: sliding 3x3 register window down the tile rows
    const f32x2 qa0 = ldp2(sco + 3 * CC + cp * 2);
    const f32x2 qa1 = ldp2(sco + 4 * CC + cp * 2);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int wi = wcol + half * NCOL;
      const bf16* esm = rawE + (size_t)wi * CC + cp * 2;  // row hl at + hl*W*CC
      const float* tb = tile + (wi + 2) * CC + cp * 2;    // tap (kh,kw) of input row hl: tile row hl+2-kh, column -kw
      bf16* dp = dE + (((long)p * H + hi0) * W + wi) * C + cch;
      f32x2 R[3][3];
      auto load_row = [&](const int r) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) R[r % 3][kw] = ldp2(tb + r * row_step - kw * CC);
      };
      load_row(0);
      load_row(1);
#pragma unroll
      for (int hl = 0; hl < THI; ++hl) {
        load_row(hl + 2);
        const f32x2 e2 = ldp2(esm + hl * (W * CC));
        f32x2 sg;
        const f32x2 ea = bnsilu_grad2_bf16(e2, qa0, qa1, sg);
        f32x2 acc = 0ull;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const f32x2 q = R[(hl + 2 - kh) % 3][kw];
            ffma2(acc, w2[kh * 3 + kw], q);
            ffma2(st2[2 + kh * 3 + kw], ea, q);
          }
        const f32x2 o = stp2_rnd(dp + hl * orow, fmul2(acc, sg));  // the stored (rounded) value
        // statistics on the raw E: sum(o) and sum(o*e); sum(o*xhat) is formed once per CTA after the tile loop
        fadd2(st2[0], o);
        ffma2(st2[1], o, e2);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  float st[11][2];
#pragma unroll
  for (int q = 0; q < 11; ++q) upk2(st2[q], st[q][0], st[q][1]);
#pragma unroll
  for (int j = 0; j < 2; ++j)  // sum(o*xhat1) from the raw-E sums
    st[1][j] = coef1[3 * C + cch + j] * (st[1][j] - coef1[2 * C + cch + j] * st[0][j]);
  block_reduce_channels<11, 2>(st, reinterpret_cast<float*>(smem_v3), cpn, NCOL, partial + (long)worker * 11 * C, C, c0);
}
