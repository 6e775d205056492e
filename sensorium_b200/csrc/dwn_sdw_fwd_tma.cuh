// TMA-staged spatial depth-wise forward (bf16 pipeline, sm_100a).
//
// Same arithmetic, in the same order, as sdw_fwd_v3_kernel of dwn_sdw_v3.cuh (S_raw and the BN2 partial sums are
// bit-identical for the same tile -> worker map); what changes is the staging and the tile geometry:
//   * ONE elected thread issues one cp.async.bulk.tensor.4d copy per item (box = CC channels x W columns x NR rows, rows
//     outside the image are filled by the tensor map) into a ring of NB raw buffers tracked by mbarriers; the other 255
//     threads spend no instruction on addresses (v3: NIT cp.async + their address / predicate chains per thread and item);
//   * stride 2: items of 4 output rows (9 staged rows, halo overhead 1.125 instead of 1.25 at 2 rows) and an activated tile
//     whose even / odd columns live in two planes, so the stride-2 taps of a warp read contiguous shared memory (v3: 33 % of
//     the shared-memory wavefronts were bank conflicts, profiles/r2_ncu_full_stencils_after.csv);
//   * where shared memory allows, the activated tile is double-buffered: one CTA barrier per item instead of two.
#pragma once
#include "dwn_sdw_tma.cuh"

template <int S, int THO, int CC>
struct SdwFwdGeom {
  static constexpr int Wo = 1024 / CC, W = Wo * S, WP = W + 2;
  static constexpr int cvn = CC / 8;
  static constexpr int NR = (THO - 1) * S + 3;
  static constexpr int VPR = W * cvn;  // 16-byte vectors per staged row: 128 (stride 1) or 256 (stride 2)
  static constexpr int RPI = 256 / VPR;
  static constexpr int NIT = (NR + RPI - 1) / RPI;
  static constexpr int RAW_BYTES = NR * VPR * 16;  // == the TMA box, a multiple of 2048
  // activated tile, bf16.  stride 1: [NR][W + 2][CC] (tile column a = w + 1, columns 0 and W + 1 stay zero).
  // stride 2: per row an even plane (a = 0, 2, ... W: EP pixels) followed by an odd plane (a = 1, 3, ... W - 1: Wo pixels);
  // with 64-byte pixels (CC = 32) the even plane is padded by one pixel so that the two pixels a quarter-warp of the
  // activation pass writes (one per plane) fall into different halves of the 128-byte bank window
  static constexpr int EP = Wo + 1 + (CC == 32 ? 1 : 0);
  static constexpr int ROWP = S == 1 ? WP * CC : (EP + Wo) * CC;  // elements per tile row
  static constexpr int ACT_BYTES = (NR * ROWP * 2 + 127) / 128 * 128;
  static constexpr int LIM = 115712 - 64;  // two CTAs per SM
  static constexpr int NB = (3 * RAW_BYTES + 2 * ACT_BYTES <= LIM) ? 3 : (2 * RAW_BYTES + 2 * ACT_BYTES <= LIM) ? 2
                          : (3 * RAW_BYTES + ACT_BYTES <= LIM) ? 3 : 2;
  static constexpr int NT = (3 * RAW_BYTES + 2 * ACT_BYTES <= LIM || 2 * RAW_BYTES + 2 * ACT_BYTES <= LIM) ? 2 : 1;
  static constexpr bool FITS = NB * RAW_BYTES + NT * ACT_BYTES <= LIM;
  static constexpr int OFF_ACT = NB * RAW_BYTES;
  static constexpr int OFF_BAR = OFF_ACT + NT * ACT_BYTES;
  static constexpr int SMEM = OFF_BAR + 64;
  static_assert(NR % RPI == 0, "staged rows must be a multiple of the rows per activation pass");
  static_assert(SMEM >= 256 * 2 * 4 * 4, "reduction scratch");
};

template <int S, int THO, int CC>
__global__ void __launch_bounds__(256, 2)
sdw_fwd_v6_kernel(const __grid_constant__ CUtensorMap map_in, const float* __restrict__ coef, const float* __restrict__ wgt,
                  bf16* __restrict__ out, float* __restrict__ partial, int NP, int H, int C, int nchunks, int nbsh) {
  using G = SdwFwdGeom<S, THO, CC>;
  constexpr int Wo = G::Wo, W = G::W, WP = G::WP, cvn = G::cvn, NR = G::NR, RPI = G::RPI, NIT = G::NIT, NB = G::NB, NT = G::NT;
  constexpr int ROWP = G::ROWP, EP = G::EP;
  extern __shared__ __align__(128) unsigned char smem_f6[];
  bf16* act0 = reinterpret_cast<bf16*>(smem_f6 + G::OFF_ACT);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_f6 + G::OFF_BAR);  // [NB]
  const int Ho = H / S;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  if (tid == 0) {
    for (int b = 0; b < NB; ++b) bk_mbar_init(full + b, 1);
    bk_mbar_init_fence();
  }
  // ---- activation pass: loop-invariant coordinates of this thread (vector tid + it*256 of the staged tile)
  constexpr int vpr = G::VPR;
  const int r_first = tid / vpr;  // 0 (stride 2) or 0 / 1 (stride 1)
  const int wq = (tid % vpr) / cvn;
  const int lcv = tid % cvn;
  int act_off0;
  if (S == 1) {
    act_off0 = (r_first * WP + wq + 1) * CC + lcv * 8;
  } else {
    const int a = wq + 1;
    act_off0 = ((a & 1) ? (EP + (a - 1) / 2) : (a / 2)) * CC + lcv * 8;
  }
  constexpr int act_step = RPI * ROWP;
  f32x2 lp0[4], lp1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a0, a1, b0, b1;
    BnSilu<bf16>::prep(coef[c0 + lcv * 8 + 2 * j], coef[C + c0 + lcv * 8 + 2 * j], a0, b0);
    BnSilu<bf16>::prep(coef[c0 + lcv * 8 + 2 * j + 1], coef[C + c0 + lcv * 8 + 2 * j + 1], a1, b1);
    lp0[j] = pk2(a0, a1);
    lp1[j] = pk2(b0, b1);
  }
  // ---- stencil: thread = (channel quad, output column)
  constexpr int cqn = CC >> 2;
  const int cq = tid % cqn, wo = tid / cqn;
  // tap kw of output column wo: stride 1 tile column wo + kw; stride 2 tile column 2 wo + kw = even[wo], odd[wo], even[wo + 1]
  constexpr int tap1 = S == 1 ? CC : EP * CC, tap2 = S == 1 ? 2 * CC : CC;
  const int rd_off = wo * CC + cq * 4;
  f32x2 w2[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w2[k][0] = pk2(wgt[(c0 + cq * 4 + 0) * 9 + k], wgt[(c0 + cq * 4 + 1) * 9 + k]);
    w2[k][1] = pk2(wgt[(c0 + cq * 4 + 2) * 9 + k], wgt[(c0 + cq * 4 + 3) * 9 + k]);
  }
  f32x2 st2[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
  // zero halo columns, once: they are never written by the activation pass
  for (int nt = 0; nt < NT; ++nt) {
    bf16* act = act0 + (size_t)nt * (G::ACT_BYTES / 2);
    if (S == 1) {
      for (int i = tid; i < NR * 2 * CC; i += 256) {
        const int r = i / (2 * CC), rem = i % (2 * CC);
        act[r * ROWP + ((rem / CC) ? (W + 1) : 0) * CC + (rem % CC)] = __float2bfloat16_rn(0.f);
      }
    } else {
      for (int i = tid; i < NR * CC; i += 256) act[(i / CC) * ROWP + (i % CC)] = __float2bfloat16_rn(0.f);
    }
  }
  const int nbm = (1 << nbsh) - 1;
  const int ntiles = NP << nbsh;
  auto issue = [&](int t, int b) {
    const int p = t >> nbsh, hi0 = (t & nbm) * (THO * S) - 1;
    bk_mbar_expect_tx(full + b, G::RAW_BYTES);
    tma_load_4d(smem_f6 + (size_t)b * G::RAW_BYTES, &map_in, full + b, c0, 0, hi0, p);
  };
  __syncthreads();  // barriers initialised, halo columns zeroed
  int t = worker;
  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (t + j * nworkers < ntiles) issue(t + j * nworkers, j);
  }
  int b = 0;
  uint32_t ph = 0;  // parity of the current use of buffer b
  for (uint32_t k = 0; t < ntiles; t += nworkers, ++k) {
    const int p = t >> nbsh, ho0 = (t & nbm) * THO;
    const int hi0 = ho0 * S - 1;
    bf16* act = act0 + (NT == 2 ? (size_t)(k & 1) * (G::ACT_BYTES / 2) : 0);
    if (NT == 1 && k > 0) __syncthreads();  // every warp has finished the previous item's stencil: the tile can be overwritten
    bk_mbar_wait(full + b, ph);
    // ---- BN1 + SiLU pass: staged bf16 -> activated bf16 tile (zero rows outside the image: padding applies after the activation)
    {
      const bf16* rp = reinterpret_cast<const bf16*>(smem_f6 + (size_t)b * G::RAW_BYTES) + tid * 8;
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
            ffma2(h, pk2(lo, hi2), lp0[j]);  // h = x*p0 + p1
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
        *reinterpret_cast<uint4*>(dst + it * act_step) = o;
      }
    }
    __syncthreads();  // tile complete; the staged buffer is free
    if (tid == 0 && t + NB * nworkers < ntiles) issue(t + NB * nworkers, b);
    if (++b == NB) { b = 0; ph ^= 1; }
    // ---- stencil: sliding 3-row register window, packed fp32x2 FMAs
    const bf16* act_rd = act + rd_off;
    f32x2 R[3][3][2];
    auto load_row = [&](int r) {
      const bf16* src = act_rd + r * ROWP;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint2 q = *reinterpret_cast<const uint2*>(src + (kw == 0 ? 0 : kw == 1 ? tap1 : tap2));
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
  __syncthreads();  // nothing in flight: every issued copy was waited for
  if (partial) {
    float st[2][4];
    upk2(st2[0][0], st[0][0], st[0][1]); upk2(st2[0][1], st[0][2], st[0][3]);
    upk2(st2[1][0], st[1][0], st[1][1]); upk2(st2[1][1], st[1][2], st[1][3]);
    block_reduce_channels<2, 4>(st, reinterpret_cast<float*>(smem_f6), cqn, Wo, partial + (long)worker * 2 * C, C, c0);
  }
}

// returns 1 if the shape is not eligible (the caller falls back to the cp.async kernels), 0 on success, -1 on error
template <int S>
static int sdw_fwd_v6_launch(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP,
                             int H, int W, int C, int tho_pref, cudaStream_t st) {
  if (H % S != 0 || W % S != 0) return 1;
  const int Ho = H / S, Wo = W / S;
  if (Wo < 8 || Wo > 32 || 1024 % Wo != 0) return 1;
  const int CC = 1024 / Wo;
  if (C % CC != 0) return 1;
  // rows per item (measured, tests/gpu_checks/check_sdw_fwd_tma.py --time): stride 1: 16 where two CTAs still fit (-6 % vs 8),
  // stride 2: 4 (-9 % vs 2)
  int THO = tho_pref > 0 ? tho_pref : (S == 1 ? 16 : 4);
  if (S == 1 && THO > 8 && CC > 64) THO = 8;
  while (THO > (S == 1 ? 4 : 2) && Ho % THO != 0) THO /= 2;
  if (Ho % THO != 0) return 1;
  const int nbsh = sdw_ilog2_exact(Ho / THO);
  if (nbsh < 0) return 1;
  const int NR = (THO - 1) * S + 3;
  CUtensorMap map;
  if (sdw_make_map4(&map, in, C, W, H, NP, CC, W, NR) != 0) return -1;
  const int nchunks = C / CC;
  dim3 grid(P * nchunks), block(256);
#define LAUNCH(THO_, CC_)                                                                                       \
  {                                                                                                             \
    using G = SdwFwdGeom<S, THO_, CC_>;                                                                         \
    if constexpr (G::FITS) {                                                                                    \
      auto k = sdw_fwd_v6_kernel<S, THO_, CC_>;                                                                 \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);                            \
      k<<<grid, block, G::SMEM, st>>>(map, coef, wgt, (bf16*)out, partial, NP, H, C, nchunks, nbsh);            \
    } else return 1;                                                                                            \
  }
#define LAUNCH_CC(THO_)                                                                                         \
  {                                                                                                             \
    if (CC == 32) LAUNCH(THO_, 32) else if (CC == 64) LAUNCH(THO_, 64) else if (CC == 128) LAUNCH(THO_, 128)   \
    else return 1;                                                                                              \
  }
  if constexpr (S == 1) {
    if (THO == 4) LAUNCH_CC(4) else if (THO == 8) LAUNCH_CC(8) else if (THO == 16) LAUNCH_CC(16) else return 1;
  } else {
    if (THO == 2) LAUNCH_CC(2) else if (THO == 4) LAUNCH_CC(4) else return 1;
  }
#undef LAUNCH_CC
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}
