// Backward kernels of the DwiseNeuro core and cortex (channels-last).  Formulas: SURVEY.md §7.4.
//   BN (train):  dx = gamma*rstd*(dy - c1 - xhat*c2),  c1 = sum(dy)/N, c2 = sum(dy*xhat)/N
//   SiLU:        s'(u) = sig(u)*(1 + u*(1 - sig(u)))
// Every kernel that feeds a BatchNorm backward also emits the deterministic per-CTA partial sums
// (sum dy, sum dy*xhat) of the *next* BN in the chain, so each tensor is read once per stage.
#include "dwn_common.cuh"
#include "dwn_reduce.cuh"
#include "dwn_sdw_v3.cuh"
#include "dwn_bulk.cuh"
#include "dwn_sdw_tma.cuh"
#include <cstdlib>
#include <type_traits>

__global__ void stem_bwd_finalize_kernel(const float* __restrict__ partial, int P, int cin, const double* __restrict__ mom,
                                         double count, const float* __restrict__ w, const float* __restrict__ coef,
                                         float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         int C0);

// =================================================================================================
// BN backward finalize: partial[P][NQ][C] (quantities q0, q0+1 = sum dy, sum dy*xhat)
//   -> dgamma, dbeta (parameter gradients) and bcoef[2][C] = sums / N
// =================================================================================================
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int P, int NQ, int q0,
                                                              double count, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ bcoef,
                                                              int C) {
  __shared__ double s_red[512];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);  // block = (8 channels, 128 row slices), see finalize_colsum2
  double da, db;
  finalize_colsum2(partial, P, NQ, q0, q0 + 1, C, c < C ? c : 0, c < C, s_red, da, db);
  if (threadIdx.x >= 8 || c >= C) return;
  if (dbeta) dbeta[c] = (float)da;
  if (dgamma) dgamma[c] = (float)db;
  bcoef[c] = (float)(da / count);
  bcoef[C + c] = (float)(db / count);
}

// two BatchNorms whose sums live in the same partial table (the projection BN and the shortcut BN of a block, the two BNs
// of a cortex layer): one launch, blockIdx.y selects the set
__global__ void __launch_bounds__(1024) bn_bwd_finalize2_kernel(const float* __restrict__ partial, int P, int NQ, int q0a,
                                                               int q0b, double count, float* __restrict__ dgA,
                                                               float* __restrict__ dbA, float* __restrict__ bcA,
                                                               float* __restrict__ dgB, float* __restrict__ dbB,
                                                               float* __restrict__ bcB, int C) {
  __shared__ double s_red[512];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  const bool second = blockIdx.y != 0;
  const int q0 = second ? q0b : q0a;
  double da, db;
  finalize_colsum2(partial, P, NQ, q0, q0 + 1, C, c < C ? c : 0, c < C, s_red, da, db);
  if (threadIdx.x >= 8 || c >= C) return;
  float* dbeta = second ? dbB : dbA;
  float* dgamma = second ? dgB : dgA;
  float* bcoef = second ? bcB : bcA;
  dbeta[c] = (float)da;
  dgamma[c] = (float)db;
  bcoef[c] = (float)(da / count);
  bcoef[C + c] = (float)(db / count);
}
extern "C" int dwn_bn_bwd_finalize2(const float* partial, int P, int NQ, int q0a, int q0b, double count, float* dgammaA,
                                    float* dbetaA, float* bcoefA, float* dgammaB, float* dbetaB, float* bcoefB, int C,
                                    void* stream) {
  bn_bwd_finalize2_kernel<<<dim3((C + 7) / 8, 2), 1024, 0, (cudaStream_t)stream>>>(partial, P, NQ, q0a, q0b, count, dgammaA,
                                                                                   dbetaA, bcoefA, dgammaB, dbetaB, bcoefB, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_bn_bwd_finalize(const float* partial, int P, int NQ, int q0, double count, float* dgamma,
                                   float* dbeta, float* bcoef, int C, void* stream) {
  bn_bwd_finalize_kernel<<<(C + 7) / 8, 1024, 0, (cudaStream_t)stream>>>(partial, P, NQ, q0, count, dgamma, dbeta,
                                                                           bcoef, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// residual epilogue backward, pass 1 (reductions)          (forward: dwn_block_out)
//   partial[P][4][Co] = { sum dp*dO, sum dp*dO*yhat4, sum dO, sum dO*xhat_sc }
// =================================================================================================
template <typename T>
__global__ void block_bwd_reduce_kernel(const float* __restrict__ dO, const T* __restrict__ y_raw,
                                        const float* __restrict__ coef4, const float* __restrict__ dp,
                                        const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                        float* __restrict__ partial, int B, int Tn, int Ho, int Wo, int Ci, int Co,
                                        NearestMap mh, NearestMap mw, int cqc, FastDiv dw, FastDiv dh, FastDiv dt) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  const int ci = c % Ci;
  float m4[4], r4[4], ms[4], rs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m4[j] = coef4[2 * Co + c + j];
    r4[j] = coef4[3 * Co + c + j];
    ms[j] = coef_sc[2 * Co + c + j];
    rs[j] = coef_sc[3 * Co + c + j];
  }
  float st[4][4] = {};
  const int Hi = mh.in, Wi = mw.in;
  const int Mo = B * Tn * Ho * Wo;
  // the three streamed quads go through a per-thread cp.async pipeline (ThreadPipe, dwn_sdw_v3.cuh)
  constexpr int DEPTH = 8;
  ThreadPipe<DEPTH, 3> pipe(smem, blockDim.x, tid);
  const int m0 = blockIdx.x * ln + lane, mstep = gridDim.x * ln;
  auto issue = [&](int k) {
    const int m = m0 + k * mstep;
    if (m < Mo) {
      const int wq = dw.mod(m), r1 = dw.div(m);
      const int hq = dh.mod(r1), bt = dh.div(r1);
      cp_async16_ok(pipe.slot(k, 0), dO + (long)m * Co + c);
      pipe_issue_quad<T>(pipe.slot(k, 1), y_raw + (long)m * Co + c);
      cp_async16_ok(pipe.slot(k, 2), xin + (((long)bt * Hi + mh.src(hq)) * Wi + mw.src(wq)) * Ci + ci);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) issue(k);
  int kk = 0;
  for (int m = m0; m < Mo; m += mstep, ++kk) {
    const int b = dt.div(dh.div(dw.div(m)));
    float g[4], y[4], x[4];
    const float d = dp ? dp[b] : 1.0f;  // before the wait: its latency overlaps the pipeline wait
    cp_async_wait<DEPTH - 1>();
    quad_from(*pipe.slot(kk, 0), g);
    pipe_read_quad<T>(pipe.slot(kk, 1), y);
    quad_from(*pipe.slot(kk, 2), x);
    issue(kk + DEPTH);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d4 = d * g[j];
      st[0][j] += d4;
      st[1][j] = fmaf(d4, (y[j] - m4[j]) * r4[j], st[1][j]);
      st[2][j] += g[j];
      st[3][j] = fmaf(g[j], (x[j] - ms[j]) * rs[j], st[3][j]);
    }
  }
  cp_async_wait<0>();
  block_reduce_channels<4, 4>(st, smem, cqc, ln, partial + (long)blockIdx.x * 4 * Co, Co, blockIdx.y * cqc * 4);
}

extern "C" int dwn_block_bwd_reduce(const float* dO, const void* y_raw, const float* coef4, const float* dp,
                                    const float* xin, const float* coef_sc, float* partial, int P, int B, int Tn, int Ho,
                                    int Wo, int Ci, int Co, int stride, int Hi, int Wi, int dtype, void* stream) {
  DWN_REQUIRE(Ho == dwn_ceil_div(Hi, stride) && Wo == dwn_ceil_div(Wi, stride),
              "dwn_block_bwd_reduce: output %dx%d is not ceil(%dx%d / %d)", Ho, Wo, Hi, Wi, stride);
  const NearestMap mh(Hi, Ho), mw(Wi, Wo);
  int cqc = dwn_largest_divisor_le(Co / 4, 64), ln = 256 / cqc;
  dim3 grid(P, (Co / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * 16 * sizeof(float);
  if (ThreadPipe<8, 3>::bytes(block.x) > sm) sm = ThreadPipe<8, 3>::bytes(block.x);
  cudaFuncSetAttribute(block_bwd_reduce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(block_bwd_reduce_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (dtype == DWN_DT_F32)
    block_bwd_reduce_kernel<float><<<grid, block, sm, (cudaStream_t)stream>>>(dO, (const float*)y_raw, coef4, dp, xin,
                                                                              coef_sc, partial, B, Tn, Ho, Wo, Ci, Co,
                                                                              mh, mw, cqc, FastDiv(Wo), FastDiv(Ho), FastDiv(Tn));
  else
    block_bwd_reduce_kernel<bf16><<<grid, block, sm, (cudaStream_t)stream>>>(dO, (const bf16*)y_raw, coef4, dp, xin,
                                                                             coef_sc, partial, B, Tn, Ho, Wo, Ci, Co,
                                                                             mh, mw, cqc, FastDiv(Wo), FastDiv(Ho), FastDiv(Tn));
  DWN_LAUNCH_CHECK();
  return 0;
}

// pass 2: dY_raw = gamma4*rstd4*(dp*dO - c1 - yhat*c2)   (A operand of the pwl dgrad / wgrad GEMMs)
template <typename T>
__global__ void block_bwd_dy_kernel(const float* __restrict__ dO, const T* __restrict__ y_raw,
                                    const float* __restrict__ coef4, const float* __restrict__ bcoef4,
                                    const float* __restrict__ dp, T* __restrict__ dY, int Mo, FastDiv drows, int Co,
                                    int cqc) {
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  float a[4], bb[4], dd[4];  // dY = a*(dp*g) - dd*y - bb
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float sc = coef4[c + j], mu = coef4[2 * Co + c + j], rs = coef4[3 * Co + c + j];
    const float c1 = bcoef4[c + j], c2 = bcoef4[Co + c + j];
    a[j] = sc;
    dd[j] = sc * rs * c2;
    bb[j] = sc * (c1 - mu * rs * c2);
  }
#pragma unroll 4
  for (int m = blockIdx.x * ln + lane; m < Mo; m += gridDim.x * ln) {
    float g[4], y[4], o[4];
    ldq(dO + (long)m * Co + c, g);
    ldq(y_raw + (long)m * Co + c, y);
    const float d = dp ? dp[drows.div(m)] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fmaf(a[j], d * g[j], -fmaf(dd[j], y[j], bb[j]));
    stq(dY + (long)m * Co + c, o);
  }
}

extern "C" int dwn_block_bwd_dy(const float* dO, const void* y_raw, const float* coef4, const float* bcoef4,
                                const float* dp, void* dY, long Mo, long rows_per_b, int Co, int dtype, void* stream) {
  int cqc = dwn_largest_divisor_le(Co / 4, 64), ln = 256 / cqc;
  dim3 grid(592, (Co / 4) / cqc), block(cqc * ln);
  if (dtype == DWN_DT_F32)
    block_bwd_dy_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(dO, (const float*)y_raw, coef4, bcoef4, dp,
                                                                         (float*)dY, (int)Mo, FastDiv((int)rows_per_b), Co,
                                                                         cqc);
  else
    block_bwd_dy_kernel<bf16><<<grid, block, 0, (cudaStream_t)stream>>>(dO, (const bf16*)y_raw, coef4, bcoef4, dp,
                                                                        (bf16*)dY, (int)Mo, FastDiv((int)rows_per_b), Co,
                                                                        cqc);
  DWN_LAUNCH_CHECK();
  return 0;
}

// gradient w.r.t. the block input: point-wise dgrad + shortcut path (nearest scatter, cyclic-tile sum, BN_sc bwd)
template <int STEM_CIN>  // > 0: this is block 0 -> accumulate the stem reductions G[c][k], S[c] instead of storing dX0
__global__ void block_in_bwd_kernel(const float* __restrict__ dXpw, const float* __restrict__ dO,
                                    const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                    const float* __restrict__ bcoef_sc, const float* __restrict__ colbias,
                                    float* __restrict__ dXin, int B, int Tn, int Hi, int Wi, int Ci, int Co, NearestMap mh,
                                    NearestMap mw, int cqc, FastDiv dw, FastDiv dh, const float* __restrict__ x_in,
                                    float* __restrict__ stem_partial) {
  extern __shared__ float smem[];
  constexpr int SQ = STEM_CIN > 0 ? STEM_CIN + 1 : 1;
  float sst[SQ][4] = {};
  const int plane = Tn * Hi * Wi;
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  const int Mi = B * Tn * Hi * Wi;
  const int Ho = mh.out, Wo = mw.out;
  const int nrep = (Co - c + Ci - 1) / Ci;  // output channels fed by input channel c: c, c+Ci, ...
  // dx_sc = sum_rep a*g - dd*x - bb   (BN_sc backward, cyclic channel tile summed)
  float a[2][4], bb[4] = {0.f, 0.f, 0.f, 0.f}, dd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[r][j] = 0.f;
      if (r < nrep) {
        const int cc = c + r * Ci + j;
        const float sc = coef_sc[cc], mu = coef_sc[2 * Co + cc], rs = coef_sc[3 * Co + cc];
        const float c1 = bcoef_sc[cc], c2 = bcoef_sc[Co + cc];
        a[r][j] = sc;
        dd[j] += sc * rs * c2;
        bb[j] += sc * (c1 - mu * rs * c2);
      }
    }
  float cb[4] = {0.f, 0.f, 0.f, 0.f};
  if (colbias) ldq(colbias + c, cb);
  // the streamed quads go through a per-thread cp.async pipeline (ThreadPipe, dwn_sdw_v3.cuh):
  // slot = { dXpw quad, x quad, dO quad, dO quad of the second channel tile }.  The stem input scalars are shared by
  // the threads of a pixel and stay plain (broadcast) loads.
  constexpr int DEPTH = 4;
  constexpr int NV = 4;
  ThreadPipe<DEPTH, NV> pipe(smem, blockDim.x, tid);
  const int m0 = blockIdx.x * ln + lane, mstep = gridDim.x * ln;
  auto issue = [&](int k) {
    const int m = m0 + k * mstep;
    if (m < Mi) {
      cp_async16_ok(pipe.slot(k, 0), dXpw + (long)m * Ci + c);
      const int wq = dw.mod(m), r1 = dw.div(m);
      const int hq = dh.mod(r1), bt = dh.div(r1);
      const int ho = mh.dst(hq), wo = mw.dst(wq);  // >= 0: this position feeds the (nearest-gathered) shortcut
      if (ho >= 0 && wo >= 0) {
        const long mo = ((long)bt * Ho + ho) * Wo + wo;
        cp_async16_ok(pipe.slot(k, 1), xin + (long)m * Ci + c);
        cp_async16_ok(pipe.slot(k, 2), dO + mo * Co + c);
        if (nrep > 1) cp_async16_ok(pipe.slot(k, 3), dO + mo * Co + c + Ci);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) issue(k);
  int kk = 0;
  for (int m = m0; m < Mi; m += mstep, ++kk) {
    float o[4];
    float xs[STEM_CIN > 0 ? STEM_CIN : 1];
    if (STEM_CIN > 0) {  // issued before the wait so that their latency overlaps it
      const int b = m / plane, pos = m - b * plane;
#pragma unroll
      for (int q = 0; q < STEM_CIN; ++q) xs[q] = __ldg(&x_in[((long)b * STEM_CIN + q) * plane + pos]);
    }
    cp_async_wait<DEPTH - 1>();
    quad_from(*pipe.slot(kk, 0), o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] -= cb[j];
    const int wq = dw.mod(m), r1 = dw.div(m);
    const int hq = dh.mod(r1);
    if (mh.dst(hq) >= 0 && mw.dst(wq) >= 0) {
      float x[4], g[4];
      quad_from(*pipe.slot(kk, 1), x);
      quad_from(*pipe.slot(kk, 2), g);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += fmaf(a[0][j], g[j], -fmaf(dd[j], x[j], bb[j]));
      if (nrep > 1) {
        quad_from(*pipe.slot(kk, 3), g);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(a[1][j], g[j], o[j]);
      }
    }
    issue(kk + DEPTH);
    if (STEM_CIN > 0) {
#pragma unroll
      for (int k = 0; k < STEM_CIN; ++k) {
        const float xv = xs[k];
#pragma unroll
        for (int j = 0; j < 4; ++j) sst[k][j] = fmaf(o[j], xv, sst[k][j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sst[SQ - 1][j] += o[j];
    } else {
      stq(dXin + (long)m * Ci + c, o);
    }
  }
  cp_async_wait<0>();
  if (STEM_CIN > 0)
    block_reduce_channels<SQ, 4>(sst, smem, cqc, ln, stem_partial + (long)blockIdx.x * SQ * Ci, Ci, blockIdx.y * cqc * 4);
}

extern "C" int dwn_block_in_bwd(const float* dXpw, const float* dO, const float* xin, const float* coef_sc,
                                const float* bcoef_sc, const float* colbias, float* dXin, int B, int Tn, int Hi, int Wi,
                                int Ci, int Co, int stride, void* stream) {
  DWN_REQUIRE(Co <= 2 * Ci, "dwn_block_in_bwd: channel tiling factor > 2 unsupported (Co=%d Ci=%d)", Co, Ci);
  int cqc = dwn_largest_divisor_le(Ci / 4, 64), ln = 256 / cqc;
  dim3 grid(592, (Ci / 4) / cqc), block(cqc * ln);
  const size_t sm = ThreadPipe<4, 4>::bytes(block.x);
  cudaFuncSetAttribute(block_in_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  {  // grid-stride kernel: exactly one resident wave (592 CTAs were 1.33 waves at 3 CTAs per SM: a third of the run
     // time had two thirds of the SMs idle)
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, block_in_bwd_kernel<0>, (int)block.x, sm) == cudaSuccess &&
        occ > 0) {
      int gx = occ * dwn_num_sms() / (int)grid.y;
      if (gx >= 1) grid.x = gx;
    }
  }
  block_in_bwd_kernel<0><<<grid, block, sm, (cudaStream_t)stream>>>(
      dXpw, dO, xin, coef_sc, bcoef_sc, colbias, dXin, B, Tn, Hi, Wi, Ci, Co, NearestMap(Hi, dwn_ceil_div(Hi, stride)),
      NearestMap(Wi, dwn_ceil_div(Wi, stride)), cqc, FastDiv(Wi), FastDiv(Hi), nullptr, nullptr);
  DWN_LAUNCH_CHECK();
  return 0;
}

// block 0 variant: the gradient w.r.t. the stem output is consumed on the fly by the stem reductions
// (partial[P][6][Ci], same layout as dwn_stem_bwd) and never written to HBM.  in_channels == 5 only.
// P = grid rows; dwn_block_in_bwd_stem_rows() returns the value that fills exactly one resident wave.
extern "C" int dwn_block_in_bwd_stem_rows(int Ci) {
  int cqc = dwn_largest_divisor_le(Ci / 4, 64), ln = 256 / cqc;
  const int threads = cqc * ln, ny = (Ci / 4) / cqc;
  size_t sm = (size_t)threads * 24 * sizeof(float);
  if (ThreadPipe<4, 4>::bytes(threads) > sm) sm = ThreadPipe<4, 4>::bytes(threads);
  cudaFuncSetAttribute(block_in_bwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, block_in_bwd_kernel<5>, threads, sm) != cudaSuccess || occ < 1)
    return 592;
  const int p = occ * dwn_num_sms() / ny;
  return p >= 1 ? p : 592;
}
extern "C" int dwn_block_in_bwd_stem(const float* dXpw, const float* dO, const float* xin, const float* coef_sc,
                                     const float* bcoef_sc, const float* colbias, const float* x_in, float* stem_partial,
                                     int P, int B, int Tn, int Hi, int Wi, int Ci, int Co, int stride, void* stream) {
  DWN_REQUIRE(Co <= 2 * Ci, "dwn_block_in_bwd_stem: channel tiling factor > 2 unsupported");
  int cqc = dwn_largest_divisor_le(Ci / 4, 64), ln = 256 / cqc;
  DWN_REQUIRE(P > 0, "dwn_block_in_bwd_stem: P must be positive");
  dim3 grid(P, (Ci / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * 24 * sizeof(float);
  if (ThreadPipe<4, 4>::bytes(block.x) > sm) sm = ThreadPipe<4, 4>::bytes(block.x);
  cudaFuncSetAttribute(block_in_bwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  block_in_bwd_kernel<5><<<grid, block, sm, (cudaStream_t)stream>>>(
      dXpw, dO, xin, coef_sc, bcoef_sc, colbias, nullptr, B, Tn, Hi, Wi, Ci, Co, NearestMap(Hi, dwn_ceil_div(Hi, stride)),
      NearestMap(Wi, dwn_ceil_div(Wi, stride)), cqc, FastDiv(Wi), FastDiv(Hi), x_in, stem_partial);
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_stem_bwd_finalize(const float* partial, int P, const double* mom, const float* w, const float* coef,
                                     float* dw, float* dgamma, float* dbeta, int B, int cin, long plane, int C0,
                                     void* stream) {
  stem_bwd_finalize_kernel<<<(C0 + 7) / 8, 256, 0, (cudaStream_t)stream>>>(partial, P, cin, mom, (double)B * plane, w, coef,
                                                                            dw, dgamma, dbeta, C0);
  DWN_LAUNCH_CHECK();
  return 0;
}

// pool backward: dX[bt][hw][c] = dP[bt][c] / HW
__global__ void pool_bwd_kernel(const float* __restrict__ dP, float* __restrict__ dX, long BT, int HW, int C) {
  const int cq4 = C / 4;
  const float inv = 1.0f / (float)HW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < BT * HW * cq4; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cq4) * 4;
    const long bt = i / ((long)cq4 * HW);
    float g[4];
    ldq(dP + bt * C + c, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] *= inv;
    stq(dX + (i / cq4) * C + c, g);
  }
}
extern "C" int dwn_pool_bwd(const float* dP, float* dX, long BT, int HW, int C, void* stream) {
  long n = BT * HW * (C / 4);
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  pool_bwd_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(dP, dX, BT, HW, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// squeeze-excite backward (dwiseneuro.py:38-43).  Pp[b][k][n] = sum_{m in b} a[m][k]*dY[m][n] comes from
// the per-sample wgrad GEMM; it yields both dW_pwl and the gate gradient (u = a*g, Y = u W^T):
//   dg[b][k] = sum_n W[n][k]*Pp[b][k][n],   dW[n][k] = sum_b g[b][k]*Pp[b][k][n]
// =================================================================================================
// a1: dpre2[b][k] = g(1-g) * sum_n Wt[k][n]*Pp[b][k][n]   (warp per k, lanes over n: both operands row-contiguous)
__global__ void __launch_bounds__(256) se_bwd_a1_kernel(const float* __restrict__ Pp, const float* __restrict__ wt,
                                                       const float* __restrict__ gate, float* __restrict__ dpre2, int C,
                                                       int Co) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= C) return;
  const float* P = Pp + ((long)b * C + k) * Co;
  const float* w = wt + (long)k * Co;
  float s = 0.f;
  for (int n = lane; n < Co; n += 32) s = fmaf(w[n], P[n], s);
  s = warp_sum(s);
  if (lane == 0) {
    const float g = gate[(long)b * C + k];
    dpre2[(long)b * C + k] = s * g * (1.0f - g);
  }
}

// a2: one CTA (1024 threads) per sample: dh -> dhpre -> dmean.  Both weight passes issue their global loads in explicit
// batches of 8 with select-predication (no early-exit branch between loads): the loops the compiler produced from a plain
// `#pragma unroll 8` waited for every load before issuing the next (67 us at C = 1792 in the ncu launch list, all latency).
__global__ void __launch_bounds__(1024) se_bwd_a2_kernel(const float* __restrict__ dpre2, const float* __restrict__ hpre,
                                                        const float* __restrict__ w1, const float* __restrict__ w2,
                                                        float* __restrict__ dhpre, float* __restrict__ dmean, int C,
                                                        int RD) {
  extern __shared__ float sm[];  // dpre2[C], dhp[RD], part[nw][RD]
  float* s_dp2 = sm;
  float* s_dhp = sm + C;
  float* s_part = s_dhp + RD;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < C; k += blockDim.x) s_dp2[k] = dpre2[(long)b * C + k];
  __syncthreads();
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  // dh[r] = sum_k dpre2[k] * w2[k][r]: warp = a contiguous slice of k, lanes = r (two r per lane and pass)
  const int kc = (C + nw - 1) / nw;
  const int kbeg = min(C, wid * kc), kend = min(C, kbeg + kc);
  for (int r0 = 0; r0 < RD; r0 += 64) {
    const int ra = r0 + lane, rb = r0 + 32 + lane;
    const bool oka = ra < RD, okb = rb < RD;
    float sa = 0.f, sb = 0.f;
    for (int k0 = kbeg; k0 < kend; k0 += 8) {
      float wa[8], wb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = min(k0 + j, kend - 1);
        const bool okk = k0 + j < kend;
        wa[j] = (oka && okk) ? __ldg(&w2[(long)kk * RD + ra]) : 0.f;
        wb[j] = (okb && okk) ? __ldg(&w2[(long)kk * RD + rb]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = s_dp2[min(k0 + j, kend - 1)];
        sa = fmaf(d, wa[j], sa);
        sb = fmaf(d, wb[j], sb);
      }
    }
    if (oka) s_part[wid * RD + ra] = sa;
    if (okb) s_part[wid * RD + rb] = sb;
  }
  __syncthreads();
  for (int r = tid; r < RD; r += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += s_part[w * RD + r];
    const float u = hpre[(long)b * RD + r];
    const float sg = 1.0f / (1.0f + expf(-u));
    const float d = s * sg * (1.0f + u * (1.0f - sg));
    s_dhp[r] = d;
    dhpre[(long)b * RD + r] = d;
  }
  __syncthreads();
  // dmean[k] = sum_r dhp[r] * w1[r][k]: thread = k (coalesced rows of w1), r in batches of 8
  for (int k = tid; k < C; k += blockDim.x) {
    float s = 0.f;
    for (int r0 = 0; r0 < RD; r0 += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = (r0 + j < RD) ? __ldg(&w1[(long)min(r0 + j, RD - 1) * C + k]) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(s_dhp[min(r0 + j, RD - 1)], wv[j], s);
    }
    dmean[(long)b * C + k] = s;
  }
}

// parameter gradients (sums over the batch): dWpwl [Co][C], dW2 [C][RD], db2 [C], dW1 [RD][C], db1 [RD]
__global__ void se_bwd_b_kernel(const float* __restrict__ Pp, const float* __restrict__ gate,
                                const float* __restrict__ hpre, const float* __restrict__ mean,
                                const float* __restrict__ dpre2, const float* __restrict__ dhpre,
                                float* __restrict__ dwpwl, float* __restrict__ dw2, float* __restrict__ db2,
                                float* __restrict__ dw1, float* __restrict__ db1, int B, int C, int Co, int RD) {
  const long n_pwl = (long)C * Co, n_w2 = (long)C * RD, n_w1 = (long)RD * C;
  const long total = n_pwl + n_w2 + C + n_w1 + RD;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < n_pwl) {
      const int k = (int)(i / Co), n = (int)(i % Co);
      for (int b = 0; b < B; ++b) s = fmaf(gate[(long)b * C + k], Pp[((long)b * C + k) * Co + n], s);
      dwpwl[(long)n * C + k] = s;
      continue;
    }
    long j = i - n_pwl;
    if (j < n_w2) {
      const int k = (int)(j / RD), r = (int)(j % RD);
      for (int b = 0; b < B; ++b) {
        const float u = hpre[(long)b * RD + r];
        s = fmaf(dpre2[(long)b * C + k], u / (1.0f + expf(-u)), s);
      }
      dw2[j] = s;
      continue;
    }
    j -= n_w2;
    if (j < C) {
      for (int b = 0; b < B; ++b) s += dpre2[(long)b * C + j];
      db2[j] = s;
      continue;
    }
    j -= C;
    if (j < n_w1) {
      const int r = (int)(j / C), k = (int)(j % C);
      for (int b = 0; b < B; ++b) s = fmaf(dhpre[(long)b * RD + r], mean[(long)b * C + k], s);
      dw1[j] = s;
      continue;
    }
    j -= n_w1;
    for (int b = 0; b < B; ++b) s += dhpre[(long)b * RD + j];
    db1[j] = s;
  }
}

// wt = projection weight transposed to [C(mid)][Co] (tiny; prepared by the caller)
extern "C" int dwn_se_bwd(const float* Pp, const float* wt, const float* gate, const float* hpre, const float* mean,
                          const float* w1, const float* w2, float* dpre2, float* dhpre, float* dmean, float* dwpwl,
                          float* dw2, float* db2, float* dw1, float* db1, int B, int C, int Co, int RD, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g1((C + 7) / 8, B);
  se_bwd_a1_kernel<<<g1, 256, 0, st>>>(Pp, wt, gate, dpre2, C, Co);
  DWN_LAUNCH_CHECK();
  se_bwd_a2_kernel<<<B, 1024, ((size_t)C + RD + 32 * (size_t)RD) * sizeof(float), st>>>(dpre2, hpre, w1, w2, dhpre, dmean,
                                                                                       C, RD);
  DWN_LAUNCH_CHECK();
  long total = (long)C * Co + (long)C * RD + C + (long)RD * C + RD;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  se_bwd_b_kernel<<<gx, 256, 0, st>>>(Pp, gate, hpre, mean, dpre2, dhpre, dwpwl, dw2, db2, dw1, db1, B, C, Co, RD);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// temporal dw backward, pass 1 (statistics only): dthat = (da + dmean[b]/Nsp) * SiLU'(BN3(Tm_raw)),
// partial[P][2][C] = { sum dthat, sum dthat*xhat3 }.  dthat is NOT stored: pass 2 reads da and Tm_raw anyway and
// recomputes it in registers, which saves one write + keeps the value in fp32.
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256, 4)
tdw_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ tm, const float* __restrict__ coef3,
                      const float* __restrict__ dmean, float inv_nsp, float* __restrict__ partial, int Nsp, int C,
                      int cqc) {
  // grid (J, channel chunks, B): the per-sample SE term dmean[b][c]/Nsp is a thread constant.
  // 4 channels per thread keep the register count low enough for 4 CTAs / SM (latency hiding by occupancy).
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  const int b = blockIdx.z;
  float q0[4], q1[4], mu[4], rs[4], dm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    BnSilu<T>::prep(coef3[c + j], coef3[C + c + j], q0[j], q1[j]);
    mu[j] = coef3[2 * C + c + j];
    rs[j] = coef3[3 * C + c + j];
    dm[j] = dmean[(long)b * C + c + j] * inv_nsp;
  }
  float st[2][4] = {};
  const T* gp = da + (long)b * Nsp * C + c;
  const T* xp = tm + (long)b * Nsp * C + c;
#pragma unroll 4
  for (int r = blockIdx.x * ln + lane; r < Nsp; r += gridDim.x * ln) {
    float g[4], x[4];
    ldq(gp + (long)r * C, g);
    ldq(xp + (long)r * C, x);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sg;
      BnSilu<T>::act_grad(x[j], q0[j], q1[j], sg);
      const float d = (g[j] + dm[j]) * sg;
      st[0][j] += d;
      st[1][j] = fmaf(d, (x[j] - mu[j]) * rs[j], st[1][j]);
    }
  }
  block_reduce_channels<2, 4>(st, smem, cqc, ln, partial + ((long)b * gridDim.x + blockIdx.x) * 2 * C, C,
                              blockIdx.y * cqc * 4);
}

// Bulk-staged variant (bf16/fp32, C/8 <= 256): CTA (j, b) streams contiguous chunks of R rows x C channels of da and
// Tm_raw through a 3-stage cp.async.bulk ring and reduces them from shared memory with 16-byte LDS.  The statistics
// are taken on the raw x (sum d, sum d*x); sum d*xhat = rstd*(sum d*x - mean*sum d) is formed once per CTA.
template <typename T>
__global__ void __launch_bounds__(256, 2)
tdw_bwd_reduce_bulk_kernel(const T* __restrict__ da, const T* __restrict__ tm, const float* __restrict__ coef3,
                           const float* __restrict__ dmean, float inv_nsp, float* __restrict__ partial, int Nsp, int C,
                           int cvc, int R) {
  constexpr int NS = 3;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int tid = threadIdx.x;
  const size_t chunk_elems = (size_t)R * C;
  T* buf = reinterpret_cast<T*>(smraw);  // [NS][2][R*C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + (size_t)NS * 2 * chunk_elems * sizeof(T));
  const int cv = tid % cvc, lane = tid / cvc, ln = blockDim.x / cvc;
  const int c = cv * 8;
  const int b = blockIdx.y, j = blockIdx.x, J = gridDim.x;
  const int nch = (Nsp + R - 1) / R;
  const T* gbase = da + (long)b * Nsp * C;
  const T* xbase = tm + (long)b * Nsp * C;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) bk_mbar_init(&bars[s], 1);
    bk_mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int ch, int stage) {  // one thread
    const int rows = min(R, Nsp - ch * R);
    const uint32_t bytes = (uint32_t)((size_t)rows * C * sizeof(T));
    bk_mbar_expect_tx(&bars[stage], 2 * bytes);
    bk_bulk_g2s(buf + (size_t)(stage * 2) * chunk_elems, gbase + (long)ch * R * C, bytes, &bars[stage]);
    bk_bulk_g2s(buf + (size_t)(stage * 2 + 1) * chunk_elems, xbase + (long)ch * R * C, bytes, &bars[stage]);
  };
  if (tid == 0)
    for (int s = 0; s < NS; ++s)
      if (j + s * J < nch) issue(j + s * J, s);
  float q0[8], q1[8], dm[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    BnSilu<T>::prep(coef3[c + e], coef3[C + c + e], q0[e], q1[e]);
    dm[e] = dmean[(long)b * C + c + e] * inv_nsp;
  }
  float st[2][8] = {};
  int it = 0;
  for (int ch = j; ch < nch; ch += J, ++it) {
    const int stage = it % NS;
    bk_mbar_wait(&bars[stage], (uint32_t)((it / NS) & 1));
    const int rows = min(R, Nsp - ch * R);
    const T* gs = buf + (size_t)(stage * 2) * chunk_elems + c;
    const T* xs = gs + chunk_elems;
#pragma unroll 2
    for (int r = lane; r < rows; r += ln) {
      float g[8], x[8];
      ldv(gs + (size_t)r * C, g);
      ldv(xs + (size_t)r * C, x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float sg;
        BnSilu<T>::act_grad(x[e], q0[e], q1[e], sg);
        const float d = (g[e] + dm[e]) * sg;
        st[0][e] += d;
        st[1][e] = fmaf(d, x[e], st[1][e]);
      }
    }
    __syncthreads();  // everybody is done with this stage
    if (tid == 0 && ch + NS * J < nch) issue(ch + NS * J, stage);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    st[1][e] = coef3[3 * C + c + e] * (st[1][e] - coef3[2 * C + c + e] * st[0][e]);
  block_reduce_channels<2, 8>(st, reinterpret_cast<float*>(smraw), cvc, ln, partial + ((long)b * J + j) * 2 * C, C, 0);
}

template <typename T>
static bool tdw_bwd_reduce_bulk_launch(const void* da, const void* tm, const float* coef3, const float* dmean, int Nsp,
                                       float* partial, int J, int B, int C, cudaStream_t st) {
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte vector: the kernel reads 8 channels per thread
  if (V != 8 && V != 4) return false;
  if (C % 8 != 0 || C / 8 > 256 || ((size_t)C * sizeof(T)) % 16 != 0) return false;
  const int cvc = C / 8, ln = 256 / cvc;
  int R = (int)(14336 / ((size_t)C * sizeof(T)));
  if (R < 1) R = 1;
  if (R > Nsp) R = Nsp;
  size_t sm = (size_t)3 * 2 * R * C * sizeof(T) + 64;
  const size_t sm_red = (size_t)cvc * ln * 16 * sizeof(float);
  if (sm_red > sm) sm = sm_red;
  auto k = tdw_bwd_reduce_bulk_kernel<T>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k<<<dim3(J, B), cvc * ln, sm, st>>>((const T*)da, (const T*)tm, coef3, dmean, 1.0f / Nsp, partial, Nsp, C, cvc, R);
  return true;
}

// partial must hold B*J rows of [2][C]
extern "C" int dwn_tdw_bwd_reduce(const void* da, const void* tm, const float* coef3, const float* dmean, int Nsp,
                                  float* partial, int J, int B, int C, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DWN_REQUIRE(C % 4 == 0, "dwn_tdw_bwd_reduce: C %% 4 != 0");
  if (dtype == DWN_DT_BF16 && tdw_bwd_reduce_bulk_launch<bf16>(da, tm, coef3, dmean, Nsp, partial, J, B, C, st)) {
    DWN_LAUNCH_CHECK();
    return 0;
  }
  int cqc = dwn_largest_divisor_le(C / 4, 128), ln = 256 / cqc;
  if (ln < 1) ln = 1;
  dim3 grid(J, (C / 4) / cqc, B), block(cqc * ln);
  if (dtype == DWN_DT_F32)
    tdw_bwd_reduce_kernel<float><<<grid, block, block.x * 8 * sizeof(float), st>>>((const float*)da, (const float*)tm, coef3,
                                                                                  dmean, 1.0f / Nsp, partial, Nsp, C, cqc);
  else
    tdw_bwd_reduce_kernel<bf16><<<grid, block, block.x * 8 * sizeof(float), st>>>((const bf16*)da, (const bf16*)tm, coef3,
                                                                                 dmean, 1.0f / Nsp, partial, Nsp, C, cqc);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// temporal dw backward, pass 2 (thread = channel pair x position, whole T column in registers, packed fp32x2):
//   dthat = (da + dmean[b]/Nsp) * SiLU'(BN3(Tm_raw))           (recomputed, see pass 1)
//   dTm = gamma3*rstd3*(dthat - c1 - xhat3*c2) = a3*dthat - d3*x - b3
//   dS_act[u] = sum_k w[k]*dTm[u-k+2] ;  dw[k] += s_act[u]*dTm[u-k+2]
//   dshat[u] = dS_act[u]*SiLU'(BN2(S_raw[u]))  -> written in place over dthat
//   partial[P][7][C] = { sum dshat, sum dshat*xhat2, dw[0..4] }
// =================================================================================================
template <typename T, int TT>
__global__ void __launch_bounds__(128, 4)
tdw_bwd_kernel(T* __restrict__ dth, const T* __restrict__ tm, const T* __restrict__ s_raw,
               const float* __restrict__ coef3, const float* __restrict__ bcoef3, const float* __restrict__ coef2,
               const float* __restrict__ wgt, const float* __restrict__ dmean, float inv_nsp,
               float* __restrict__ partial, int B, int Tn, int HW, int C, int cpc) {
  // thread = one channel PAIR x position (2 channels keep the whole-T register column small enough for
  // 4 CTAs / SM; a warp still covers 64 consecutive channels = 128 bytes per row)
  constexpr int TA = TT > 0 ? TT : 32;
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cp = tid % cpc, lane = tid / cpc, ln = blockDim.x / cpc;
  const int c = (blockIdx.y * cpc + cp) * 2;
  const int tn = TT > 0 ? TT : Tn;
  f32x2 a3, b3, d3, w2[5];
  float p0[2], p1[2], mu2[2], rs2[2], q0[2], q1[2];
  {
    float av[2], bv[2], dv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cc = c + e;
      const float sc = coef3[cc], mu = coef3[2 * C + cc], rs = coef3[3 * C + cc];
      const float k1 = bcoef3[cc], k2 = bcoef3[C + cc];
      av[e] = sc;
      dv[e] = -sc * rs * k2;
      bv[e] = -sc * (k1 - mu * rs * k2);
      BnSilu<T>::prep(coef2[cc], coef2[C + cc], p0[e], p1[e]);
      BnSilu<T>::prep(sc, coef3[C + cc], q0[e], q1[e]);
      mu2[e] = coef2[2 * C + cc];
      rs2[e] = coef2[3 * C + cc];
    }
    a3 = pk2(av[0], av[1]);
    b3 = pk2(bv[0], bv[1]);
    d3 = pk2(dv[0], dv[1]);
#pragma unroll
    for (int k = 0; k < 5; ++k) w2[k] = pk2(wgt[c * 5 + k], wgt[(c + 1) * 5 + k]);
  }
  f32x2 st2[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) st2[q] = 0ull;
  const long npos = (long)B * HW;
  const long tstride = (long)HW * C;
  const long bstride = (long)Tn * tstride;
  for (long pos = (long)blockIdx.x * ln + lane; pos < npos; pos += (long)gridDim.x * ln) {
    const long b = pos / HW, hw = pos - b * HW;
    const long base = b * bstride + hw * C + c;
    T* gp = dth + base;
    const T* xp = tm + base;
    const T* sp = s_raw + base;
    const float dm0 = dmean[b * C + c] * inv_nsp, dm1 = dmean[b * C + c + 1] * inv_nsp;
    f32x2 dT[TA], sv[TA];
#pragma unroll
    for (int t = 0; t < TA; ++t)
      if (t < tn) sv[t] = ldp2(sp + t * tstride);
#pragma unroll
    for (int t = 0; t < TA; ++t) {
      if (t < tn) {
        const f32x2 g = ldp2(gp + t * tstride), x = ldp2(xp + t * tstride);
        float g0, g1, x0, x1, s0, s1;
        upk2(g, g0, g1);
        upk2(x, x0, x1);
        BnSilu<T>::act_grad(x0, q0[0], q1[0], s0);
        BnSilu<T>::act_grad(x1, q0[1], q1[1], s1);
        const f32x2 dth2 = pk2((g0 + dm0) * s0, (g1 + dm1) * s1);
        f32x2 v = b3;  // dTm = a3*dthat + d3*x + b3   (d3, b3 carry the minus signs)
        ffma2(v, a3, dth2);
        ffma2(v, d3, x);
        dT[t] = v;
      }
    }
#pragma unroll
    for (int u = 0; u < TA; ++u) {
      if (u < tn) {
        float s[2], sa[2], sg[2];
        upk2(sv[u], s[0], s[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) sa[j] = BnSilu<T>::act_grad(s[j], p0[j], p1[j], sg[j]);
        const f32x2 sa2 = pk2(sa[0], sa[1]);
        f32x2 acc = 0ull;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int t = u - k + 2;
          if (t >= 0 && t < tn) {
            ffma2(acc, w2[k], dT[t]);
            ffma2(st2[2 + k], sa2, dT[t]);
          }
        }
        const f32x2 o = stp2_rnd(gp + u * tstride, fmul2(acc, pk2(sg[0], sg[1])));  // BN2-backward sums over the stored values
        fadd2(st2[0], o);
        ffma2(st2[1], o, pk2((s[0] - mu2[0]) * rs2[0], (s[1] - mu2[1]) * rs2[1]));
      }
    }
  }
  float st[7][2];
#pragma unroll
  for (int q = 0; q < 7; ++q) upk2(st2[q], st[q][0], st[q][1]);
  block_reduce_channels<7, 2>(st, smem, cpc, ln, partial + (long)blockIdx.x * 7 * C, C, blockIdx.y * cpc * 2);
}

// Bulk-staged variant of pass 2 (bf16): one work item = one (b, hw) position x CCH channels x all T.  The 3*T row
// segments of an item (da, Tm_raw, S_raw; CCH*2 bytes each, contiguous in the channels-last layout) are fetched by
// cp.async.bulk into a 2-stage shared-memory ring, so the bytes in flight do not depend on registers or occupancy;
// thread = channel pair, whole-T column in registers, same arithmetic as tdw_bwd_kernel.
// Warp-specialised: the LAST warp is the producer (waits for a stage to be released, issues its copies); the CCH/64
// consumer warps wait on the stage's "full" barrier, compute, and release it with one arrival per warp on its "empty"
// barrier — no CTA-wide barrier in the item loop (ncu: 2.8 warps per issue slot were parked at __syncthreads).
template <int TT>
__global__ void __launch_bounds__(256, 2)
tdw_bwd_bulk_kernel(bf16* __restrict__ dth, const bf16* __restrict__ tm, const bf16* __restrict__ s_raw,
                    const float* __restrict__ coef3, const float* __restrict__ bcoef3, const float* __restrict__ coef2,
                    const float* __restrict__ wgt, const float* __restrict__ dmean, float inv_nsp,
                    float* __restrict__ partial, int B, int Tn, int HW, int C, int CCH) {
  constexpr int TA = TT > 0 ? TT : 32;
  constexpr int NS = 2;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int tn = TT > 0 ? TT : Tn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * CCH, c = c0 + tid * 2;
  bf16* buf = reinterpret_cast<bf16*>(smraw);  // [NS][3][tn][CCH]
  const size_t stage_elems = (size_t)3 * tn * CCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + NS * stage_elems * sizeof(bf16));   // full[NS], empty[NS]
  uint64_t* ebars = bars + NS;
  const long npos = (long)B * HW;
  const long tstride = (long)HW * C;
  const long bstride = (long)Tn * tstride;
  const int G = gridDim.x;
  const int ncw = CCH / 64;                      // consumer warps; warp ncw is the producer
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { bk_mbar_init(&bars[s], 1); bk_mbar_init(&ebars[s], (uint32_t)ncw); }
    bk_mbar_init_fence();
  }
  __syncthreads();
  const uint32_t seg_bytes = (uint32_t)(CCH * sizeof(bf16));
  auto issue = [&](long pos, int stage) {  // all lanes of the producer warp
    const long b = pos / HW, hw = pos - b * HW;
    const long base = b * bstride + hw * C + c0;
    if (lane == 0) bk_mbar_expect_tx(&bars[stage], 3u * (uint32_t)tn * seg_bytes);
    __syncwarp();
    bf16* dst = buf + (size_t)stage * stage_elems;
    for (int i = lane; i < 3 * tn; i += 32) {
      const int a = i / tn, t = i - a * tn;
      const bf16* src = (a == 0 ? (const bf16*)dth : (a == 1 ? tm : s_raw)) + base + (long)t * tstride;
      bk_bulk_g2s(dst + (size_t)i * CCH, src, seg_bytes, &bars[stage]);
    }
  };
  long pos = blockIdx.x;
  f32x2 st2[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) st2[q] = 0ull;
  if (warp == ncw) {
    // ================= producer warp =================
    int it = 0;
    for (; pos < npos; pos += G, ++it) {
      const int stage = it & 1;
      if (it >= NS) bk_mbar_wait(&ebars[stage], (uint32_t)(((it >> 1) - 1) & 1));
      issue(pos, stage);
    }
  } else {
  f32x2 a3, b3, d3, w2[5], p0_2, p1_2, q0_3, q1_3, mu2, rs2;
  {
    float av[2], bv[2], dv[2], p0[2], p1[2], q0[2], q1[2], m2[2], r2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cc = c + e;
      const float sc = coef3[cc], mu = coef3[2 * C + cc], rs = coef3[3 * C + cc];
      const float k1 = bcoef3[cc], k2 = bcoef3[C + cc];
      av[e] = sc;
      dv[e] = -sc * rs * k2;
      bv[e] = -sc * (k1 - mu * rs * k2);
      BnSilu<bf16>::prep(coef2[cc], coef2[C + cc], p0[e], p1[e]);
      BnSilu<bf16>::prep(sc, coef3[C + cc], q0[e], q1[e]);
      m2[e] = coef2[2 * C + cc];
      r2[e] = coef2[3 * C + cc];
    }
    a3 = pk2(av[0], av[1]); b3 = pk2(bv[0], bv[1]); d3 = pk2(dv[0], dv[1]);
    p0_2 = pk2(p0[0], p0[1]); p1_2 = pk2(p1[0], p1[1]);
    q0_3 = pk2(q0[0], q0[1]); q1_3 = pk2(q1[0], q1[1]);
    mu2 = pk2(-m2[0] * r2[0], -m2[1] * r2[1]);  // xhat2 = s*rs2 + (-mu2*rs2)
    rs2 = pk2(r2[0], r2[1]);
#pragma unroll
    for (int k = 0; k < 5; ++k) w2[k] = pk2(wgt[c * 5 + k], wgt[(c + 1) * 5 + k]);
  }
  // the per-sample SE term of the NEXT item is fetched one item ahead (no dependent global load per item)
  f32x2 dm_next = 0ull;
  if (pos < npos) {
    const long b = pos / HW;
    dm_next = pk2(dmean[b * C + c] * inv_nsp, dmean[b * C + c + 1] * inv_nsp);
  }
  int it = 0;
  for (; pos < npos; pos += G, ++it) {
    const int stage = it & 1;
    const f32x2 dm = dm_next;
    if (pos + G < npos) {
      const long bn = (pos + G) / HW;
      dm_next = pk2(dmean[bn * C + c] * inv_nsp, dmean[bn * C + c + 1] * inv_nsp);
    }
    const long b = pos / HW, hw = pos - b * HW;
    bf16* gp = dth + b * bstride + hw * C + c;
    bk_mbar_wait(&bars[stage], (uint32_t)((it >> 1) & 1));
    const bf16* sg_ = buf + (size_t)stage * stage_elems + tid * 2;  // da rows
    const bf16* sx_ = sg_ + (size_t)tn * CCH;                        // Tm_raw rows
    const bf16* ss_ = sx_ + (size_t)tn * CCH;                        // S_raw rows
    f32x2 dT[TA];
#pragma unroll
    for (int t = 0; t < TA; ++t) {
      if (t < tn) {
        const f32x2 g = ldp2(sg_ + (size_t)t * CCH), x = ldp2(sx_ + (size_t)t * CCH);
        f32x2 s3;
        bnsilu_grad2_bf16(x, q0_3, q1_3, s3);
        f32x2 gd = g;
        fadd2(gd, dm);
        const f32x2 dth2 = fmul2(gd, s3);
        f32x2 v = b3;  // dTm = a3*dthat + d3*x + b3   (d3, b3 carry the minus signs)
        ffma2(v, a3, dth2);
        ffma2(v, d3, x);
        dT[t] = v;
      }
    }
#pragma unroll
    for (int u = 0; u < TA; ++u) {
      if (u < tn) {
        const f32x2 s = ldp2(ss_ + (size_t)u * CCH);
        f32x2 sgr;
        const f32x2 sa2 = bnsilu_grad2_bf16(s, p0_2, p1_2, sgr);
        f32x2 acc = 0ull;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int t = u - k + 2;
          if (t >= 0 && t < tn) {
            ffma2(acc, w2[k], dT[t]);
            ffma2(st2[2 + k], sa2, dT[t]);
          }
        }
        const f32x2 o = stp2_rnd(gp + (long)u * tstride, fmul2(acc, sgr));  // BN2-backward sums over the stored values
        fadd2(st2[0], o);
        f32x2 xh = mu2;
        ffma2(xh, s, rs2);
        ffma2(st2[1], o, xh);
      }
    }
    __syncwarp();     // every lane of this warp has consumed its columns of the stage
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem_u32(&ebars[stage])) : "memory");
  }
  }  // consumer warps
  float st[7][2];
#pragma unroll
  for (int q = 0; q < 7; ++q) upk2(st2[q], st[q][0], st[q][1]);
  // the producer warp's zeros land in a second (unused) lane slice of the scratch: ln = 1 sums slice 0 only
  block_reduce_channels<7, 2>(st, reinterpret_cast<float*>(smraw), CCH / 2, 1, partial + (long)blockIdx.x * 7 * C, C, c0);
}

static bool tdw_bwd_bulk_launch(void* dth, const void* tm, const void* s_raw, const float* coef3, const float* bcoef3,
                                const float* coef2, const float* wgt, const float* dmean, float* partial, int P, int B,
                                int Tn, int HW, int C, cudaStream_t st) {
  if (C % 16 != 0) return false;
  int CCH = 0;  // channels per CTA: a divisor of C, multiple of 64 (whole warps of channel pairs), <= 448 (7 consumer
  //               warps + the producer warp = 256 threads)
  for (int cand = 448; cand >= 64; cand -= 64)
    if (C % cand == 0) { CCH = cand; break; }
  if (CCH == 0) return false;
  const size_t stage = (size_t)3 * Tn * CCH * sizeof(bf16);
  size_t sm = 2 * stage + 64;
  const size_t sm_red = (size_t)(CCH / 2 + 32) * 14 * sizeof(float) * 2;  // two lane slices (consumers, producer warp)
  if (sm_red > sm) sm = sm_red;
  if (sm > 110 * 1024) return false;
  dim3 grid(P, C / CCH), block(CCH / 2 + 32);  // + the producer warp
#define GOB(TTV)                                                                                                   \
  {                                                                                                                \
    auto k = tdw_bwd_bulk_kernel<TTV>;                                                                             \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                                 \
    k<<<grid, block, sm, st>>>((bf16*)dth, (const bf16*)tm, (const bf16*)s_raw, coef3, bcoef3, coef2, wgt, dmean,  \
                               1.0f / ((float)Tn * (float)HW), partial, B, Tn, HW, C, CCH);                        \
  }
  if (Tn == 16) GOB(16) else if (Tn == 8) GOB(8) else GOB(0)
#undef GOB
  return true;
}

extern "C" int dwn_tdw_bwd(void* dth, const void* tm, const void* s_raw, const float* coef3, const float* bcoef3,
                           const float* coef2, const float* wgt, const float* dmean, float* partial, int P, int B, int Tn,
                           int HW, int C, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DWN_REQUIRE(Tn <= 32, "dwn_tdw_bwd: T > 32 unsupported");
  DWN_REQUIRE(C % 2 == 0, "dwn_tdw_bwd: C %% 2 != 0");
  if (dtype == DWN_DT_BF16 &&
      tdw_bwd_bulk_launch(dth, tm, s_raw, coef3, bcoef3, coef2, wgt, dmean, partial, P, B, Tn, HW, C, st)) {
    DWN_LAUNCH_CHECK();
    return 0;
  }
  int cpc = dwn_largest_divisor_le(C / 2, 128);
  int ln = 128 / cpc;
  if (ln < 1) ln = 1;
  dim3 grid(P, (C / 2) / cpc), block(cpc * ln);
  size_t sm = (size_t)block.x * 14 * sizeof(float);
#define GO(TY, TTV)                                                                                              \
  tdw_bwd_kernel<TY, TTV><<<grid, block, sm, st>>>((TY*)dth, (const TY*)tm, (const TY*)s_raw, coef3, bcoef3, coef2, wgt, \
                                                   dmean, 1.0f / ((float)Tn * (float)HW), partial, B, Tn, HW, C, cpc)
  if (dtype == DWN_DT_F32) {
    if (Tn == 16) GO(float, 16); else if (Tn == 8) GO(float, 8); else GO(float, 0);
  } else {
    if (Tn == 16) GO(bf16, 16); else if (Tn == 8) GO(bf16, 8); else GO(bf16, 0);
  }
#undef GO
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// spatial dw backward.  CTA tile: one (b,t) plane, THI input rows, all W, CC channels.
//   smem: dS_raw tile = gamma2*rstd2*(dshat - c1 - xhat2*c2) with zero halo (BN2 backward applied on load)
//   thread (channel quad, wi): for each input row: load E_raw once -> ehat, e_act, SiLU';
//     dE_act = sum_taps w*dS_raw ;  dw[tap] += e_act*dS_raw   (same smem values serve dgrad and wgrad)
//     dehat = dE_act*SiLU'(ehat) -> dE_pre ;  partial[P][11][C] = { sum dehat, sum dehat*xhat1, dw[0..8] }
// =================================================================================================
template <typename T, int S, int THI>
__global__ void __launch_bounds__(256, 2)
sdw_bwd_kernel(const T* __restrict__ dsh, const T* __restrict__ s_raw, const T* __restrict__ e_raw,
               const float* __restrict__ coef2, const float* __restrict__ bcoef2, const float* __restrict__ coef1,
               const float* __restrict__ wgt, T* __restrict__ dE, float* __restrict__ partial, int NP, int H, int W,
               int C, int CC, int nchunks, int wpsh, int cvsh) {
  // 1-D grid, channel chunk fastest (see sdw_fwd_kernel).  wpsh = log2(Wo+2) is never a power of two, so only
  // the channel-vector split uses a shift; the (row, col) split uses one division per vector.
  constexpr int V = VecT<T>::V;
  constexpr int NR = S == 1 ? THI + 2 : THI / 2 + 1;
  extern __shared__ float tile[];
  const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S, WP = Wo + 2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  const int cvn = CC / V;
  const int lcv = tid % cvn;
  const int cqn = CC / 4;
  const int cq = tid % cqn, wi = tid / cqn;
  const int cch = c0 + cq * 4;
  // per-channel constants live in shared memory (registers are needed for the 44 accumulators + 36 weights):
  //   sco[0..2] : BN2 backward on load  dS_raw = la*g - ld*x - lb
  //   sco[3..6] : BN1+SiLU constants p0, p1 and mean1, rstd1
  float* sco = tile + (size_t)NR * (Wo + 2) * CC;
  for (int i = tid; i < CC; i += nthr) {
    const int cc = c0 + i;
    const float sc = coef2[cc], mu = coef2[2 * C + cc], rs = coef2[3 * C + cc];
    const float k1 = bcoef2[cc], k2 = bcoef2[C + cc];
    sco[i] = sc;
    sco[CC + i] = sc * (k1 - mu * rs * k2);
    sco[2 * CC + i] = sc * rs * k2;
    float q0, q1;
    BnSilu<T>::prep(coef1[cc], coef1[C + cc], q0, q1);
    sco[3 * CC + i] = q0;
    sco[4 * CC + i] = q1;
    sco[5 * CC + i] = coef1[2 * C + cc];
    sco[6 * CC + i] = coef1[3 * C + cc];
  }
  float wr[9][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[k][j] = wgt[(cch + j) * 9 + k];
  // S == 2: w-direction tap slots are thread constants (wi parity)
  const bool odd_w = (wi & 1) != 0;
  const int colA = (S == 1) ? 0 : (odd_w ? (wi + 1) / 2 : wi / 2);
  const int colB = (wi - 1) / 2;  // only used when odd_w (kw = 2)
  float st[11][4] = {};
  const int nb = (H + THI - 1) / THI;  // the last band may be partial (H not a multiple of THI)
  const int ntiles = NP * nb;
  const int nvec = NR * WP * cvn;
  for (int t = worker; t < ntiles; t += nworkers) {
    const int p = t / nb, hi0 = (t % nb) * THI;
    const int ho_first = (S == 1) ? hi0 - 1 : hi0 / 2;
    // prefetch the first E row of this tile while the dS tile is staged
    float e_nxt[4];
    ldq(e_raw + (((long)p * H + hi0) * W + wi) * C + cch, e_nxt);
    __syncthreads();
    // ---- load dS_raw tile (cols: S==1 -> wo+1 in [0,Wo+1]; S==2 -> wo in [0,Wo])
#pragma unroll 2
    for (int i = tid; i < nvec; i += nthr) {
      const int pos = cvsh >= 0 ? (i >> cvsh) : (i / cvn);
      const int r = pos / WP;
      const int col = pos - r * WP;
      const int ho = ho_first + r;
      const int wo = (S == 1) ? col - 1 : col;
      float v[V];
      if (ho >= 0 && ho < Ho && wo >= 0 && wo < Wo) {
        float g[V], x[V];
        const long off = (((long)p * Ho + ho) * Wo + wo) * C + c0 + lcv * V;
        ldv(dsh + off, g);
        ldv(s_raw + off, x);
        const float* ca = sco + lcv * V;
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = fmaf(ca[j], g[j], -fmaf(ca[2 * CC + j], x[j], ca[CC + j]));
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = 0.f;
      }
      float* dst = tile + ((r * WP + col) * CC + lcv * V);
#pragma unroll
      for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __syncthreads();
    // rows are processed in pairs so the stride-2 tap parity is a compile-time constant without fully
    // unrolling the band (full unrolling costs >200 registers)
    auto row_body = [&](const int hl, auto par_c) {
      constexpr int PAR = decltype(par_c)::value;
      const int hi = hi0 + hl;
      if (hi >= H) return;  // partial last band
      float e[4], ea[4], sg[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
      const long eoff = (((long)p * H + hi) * W + wi) * C + cch;
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = e_nxt[j];
      if (hl + 1 < THI && hi + 1 < H) ldq(e_raw + eoff + (long)W * C, e_nxt);
      {
        const float4 q0 = *reinterpret_cast<const float4*>(sco + 3 * CC + cq * 4);
        const float4 q1 = *reinterpret_cast<const float4*>(sco + 4 * CC + cq * 4);
        const float a0[4] = {q0.x, q0.y, q0.z, q0.w}, a1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) ea[j] = BnSilu<T>::act_grad(e[j], a0[j], a1[j], sg[j]);
      }
      if (S == 1) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int r = hl - kh + 2;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 q = *reinterpret_cast<const float4*>(tile + ((r * WP + wi - kw + 2) * CC + cq * 4));
            const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j] = fmaf(wr[kh * 3 + kw][j], v[j], acc[j]);
              st[2 + kh * 3 + kw][j] = fmaf(ea[j], v[j], st[2 + kh * 3 + kw][j]);
            }
          }
        }
      } else {
        // valid kh: even row -> kh=1 (tile row hl/2); odd row -> kh=0 (row (hl+1)/2), kh=2 (row (hl-1)/2)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if ((PAR == 0) != (kh == 1)) continue;
          const int r = (kh == 1) ? hl / 2 : (kh == 0 ? (hl + 1) / 2 : (hl - 1) / 2);
          {
            const float4 q = *reinterpret_cast<const float4*>(tile + ((r * WP + colA) * CC + cq * 4));
            const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float wv = odd_w ? wr[kh * 3 + 0][j] : wr[kh * 3 + 1][j];
              acc[j] = fmaf(wv, v[j], acc[j]);
              const float pr = ea[j] * v[j];
              st[2 + kh * 3 + 0][j] += odd_w ? pr : 0.f;
              st[2 + kh * 3 + 1][j] += odd_w ? 0.f : pr;
            }
          }
          if (odd_w) {
            const float4 q = *reinterpret_cast<const float4*>(tile + ((r * WP + colB) * CC + cq * 4));
            const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j] = fmaf(wr[kh * 3 + 2][j], v[j], acc[j]);
              st[2 + kh * 3 + 2][j] = fmaf(ea[j], v[j], st[2 + kh * 3 + 2][j]);
            }
          }
        }
      }
      float o[4];
      {
        const float4 qm = *reinterpret_cast<const float4*>(sco + 5 * CC + cq * 4);
        const float4 qr = *reinterpret_cast<const float4*>(sco + 6 * CC + cq * 4);
        const float mu1[4] = {qm.x, qm.y, qm.z, qm.w}, rs1[4] = {qr.x, qr.y, qr.z, qr.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = rnd<T>(acc[j] * sg[j]);
          st[0][j] += o[j];
          st[1][j] = fmaf(o[j], (e[j] - mu1[j]) * rs1[j], st[1][j]);
        }
      }
      stq(dE + eoff, o);
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
  __syncthreads();
  block_reduce_channels<11, 4>(st, tile, cqn, W, partial + (long)worker * 11 * C, C, c0);
}

static inline int ilog2_exact_b(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

template <typename T, int S>
static int sdw_bwd_launch(const void* dsh, const void* s_raw, const void* e_raw, const float* coef2, const float* bcoef2,
                          const float* coef1, const float* wgt, void* dE, float* partial, int P, int NP, int H, int W,
                          int C, cudaStream_t st) {
  constexpr int V = VecT<T>::V;
  const int Wo = (W + S - 1) / S;
  int CC = 128;  // largest power of two that divides C with (CC/4)*W <= 256 threads
  while (CC >= 8 && (C % CC != 0 || (CC / 4) * W > 256)) CC /= 2;
  DWN_REQUIRE(CC >= 8 && C % CC == 0 && (CC / 4) * W <= 256, "dwn_sdw_bwd: unsupported C=%d W=%d", C, W);
  int THI = (H % 8 == 0) ? 8 : (H % 4 == 0 ? 4 : 2);  // odd H: bands of 2 rows, the last one partial
  const int NR = S == 1 ? THI + 2 : THI / 2 + 1;
  size_t sm = ((size_t)NR * (Wo + 2) + 7) * CC * sizeof(float);
  size_t sm_red = (size_t)(CC / 4) * W * 11 * 4 * sizeof(float);
  if (sm_red > sm) sm = sm_red;
  const int nchunks = C / CC;
  const int cvsh = ilog2_exact_b(CC / V);
  dim3 grid(P * nchunks), block((CC / 4) * W);
#define LAUNCH(THI_)                                                                                              \
  {                                                                                                               \
    auto k = sdw_bwd_kernel<T, S, THI_>;                                                                          \
    if (sm > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);            \
    k<<<grid, block, sm, st>>>((const T*)dsh, (const T*)s_raw, (const T*)e_raw, coef2, bcoef2, coef1, wgt, (T*)dE, \
                               partial, NP, H, W, C, CC, nchunks, -1, cvsh);                                      \
  }
  switch (THI) {
    case 8: LAUNCH(8) break;
    case 4: LAUNCH(4) break;
    default: LAUNCH(2) break;
  }
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}

// stride 1 runs the register-window variant (v5), stride 2 the v3 kernel
using sdw_bwd_fn = void (*)(const bf16*, const bf16*, const bf16*, const float*, const float*, const float*, const float*,
                            bf16*, float*, int, int, int, int, int);
template <int S, int THI, int CC>
static sdw_bwd_fn sdw_bwd_pick() {
  if constexpr (S == 1) return sdw_bwd_v5_kernel<THI, CC>;
  else return sdw_bwd_v3_kernel<S, THI, CC>;
}

template <int S>
static int sdw_bwd_v3_launch(const void* dsh, const void* s_raw, const void* e_raw, const float* coef2,
                             const float* bcoef2, const float* coef1, const float* wgt, void* dE, float* partial, int P,
                             int NP, int H, int W, int C, cudaStream_t st) {
  if (H % S != 0 || W % S != 0) return 1;  // ceil-sized outputs take the generic kernel
  const int Wo = W / S;
  int CC = 1024 / W;
  if (CC > 128) CC = 128;
  while (CC >= 8 && (C % CC != 0)) CC /= 2;
  if (CC < 8 || C % CC != 0 || (CC / 4) * W != 256) return 1;
  const int cvsh = ilog2_exact_b(CC / 8);
  if (cvsh < 0 || Wo * (CC / 8) != 128 / S || W * (CC / 8) != 128) return 1;
  const int THI = (H % 8 == 0) ? 8 : (H % 4 == 0 ? 4 : 0);
  if (THI == 0) return 1;
  const int nbsh = ilog2_exact_b(H / THI);
  if (nbsh < 0) return 1;
  const int NR = S == 1 ? THI + 2 : THI / 2 + 1;
  const int RPI = 2 * S, NIT = (NR + RPI - 1) / RPI;
  // the v3 kernel (stride 2) double-buffers its E tile; the v5 kernel (stride 1) keeps one
  const size_t e_bufs = S == 2 ? 2 : 1;
  size_t sm = 2 * (size_t)NIT * 256 * 16 + e_bufs * (size_t)THI * 128 * 16 +
              ((size_t)NR * (Wo + 2) + 7) * CC * sizeof(float);
  const size_t sm_red = (size_t)256 * 11 * 4 * sizeof(float);
  if (sm_red > sm) sm = sm_red;
  const int nchunks = C / CC;
  dim3 grid(P * nchunks), block(256);
#define LAUNCH(THI_, CC_)                                                                                          \
  {                                                                                                                \
    auto k = sdw_bwd_pick<S, THI_, CC_>();                                                                         \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                                 \
    k<<<grid, block, sm, st>>>((const bf16*)dsh, (const bf16*)s_raw, (const bf16*)e_raw, coef2, bcoef2, coef1, wgt, \
                               (bf16*)dE, partial, NP, H, C, nchunks, nbsh);                                       \
  }
  if (THI == 8) {
    if (CC == 16) LAUNCH(8, 16) else if (CC == 32) LAUNCH(8, 32) else if (CC == 64) LAUNCH(8, 64)
    else if (CC == 128) LAUNCH(8, 128) else return 1;
  } else {
    if (CC == 16) LAUNCH(4, 16) else if (CC == 32) LAUNCH(4, 32) else if (CC == 64) LAUNCH(4, 64)
    else if (CC == 128) LAUNCH(4, 128) else return 1;
  }
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_sdw_bwd(const void* dsh, const void* s_raw, const void* e_raw, const float* coef2,
                           const float* bcoef2, const float* coef1, const float* wgt, void* dE, float* partial, int P,
                           int NP, int H, int W, int C, int stride, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DWN_REQUIRE(stride == 1 || stride == 2, "dwn_sdw_bwd: stride %d unsupported", stride);
  if (dtype == DWN_DT_F32)
    return stride == 1
               ? sdw_bwd_launch<float, 1>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st)
               : sdw_bwd_launch<float, 2>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st);
  // TMA-staged kernels first (DWN_SDW_TMA=0 selects the cp.async kernels, DWN_SDW_THI the stride-2 tile height: A/B runs)
  const char* env_tma = getenv("DWN_SDW_TMA");
  const char* env_thi = getenv("DWN_SDW_THI");
  const int tma_mode = env_tma ? atoi(env_tma) : 1;
  const int thi_pref = env_thi ? atoi(env_thi) : 0;
  int rc = 1;
  // DWN_SDW_TMA: 0 = cp.async kernels, 1 = TMA kernels (stride 1: one-pass v7), 2 = TMA kernels (stride 1: two-pass v6)
  if (tma_mode == 1 && stride == 1)
    rc = sdw_bwd_v7_launch(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, thi_pref, st);
  if (tma_mode && rc > 0)
    rc = stride == 1 ? sdw_bwd_v6_launch<1>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, 8, st)
                     : sdw_bwd_v6_launch<2>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C,
                                            thi_pref, st);
  if (rc <= 0) return rc;
  rc = stride == 1
               ? sdw_bwd_v3_launch<1>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st)
               : sdw_bwd_v3_launch<2>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st);
  if (rc <= 0) return rc;
  return stride == 1
             ? sdw_bwd_launch<bf16, 1>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st)
             : sdw_bwd_launch<bf16, 2>(dsh, s_raw, e_raw, coef2, bcoef2, coef1, wgt, dE, partial, P, NP, H, W, C, st);
}

// BN backward apply in place: g <- gamma*rstd*(g - c1 - xhat*c2).  Thread = fixed channel vector (coefficients
// live in registers), rows strided over the grid: pure 16-byte streaming.
template <typename T>
__global__ void bn_bwd_apply_kernel(T* __restrict__ g, const T* __restrict__ x, const float* __restrict__ coef,
                                    const float* __restrict__ bcoef, long M, int C, int cvc) {
  constexpr int V = VecT<T>::V;
  const int tid = threadIdx.x;
  const int cv = tid % cvc, lane = tid / cvc, ln = blockDim.x / cvc;
  const int c = (blockIdx.y * cvc + cv) * V;
  float a[V], bb[V], dd[V];  // dx = a*g - bb - dd*x  with a = scale, dd = scale*rstd*c2, bb = scale*(c1 - mean*rstd*c2)
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float sc = coef[c + j], mu = coef[2 * C + c + j], rs = coef[3 * C + c + j];
    const float c1 = bcoef[c + j], c2 = bcoef[C + c + j];
    a[j] = sc;
    dd[j] = sc * rs * c2;
    bb[j] = sc * (c1 - mu * rs * c2);
  }
  for (long m = (long)blockIdx.x * ln + lane; m < M; m += (long)gridDim.x * ln) {
    const long off = m * C + c;
    float gv[V], xv[V];
    ldv(g + off, gv);
    ldv(x + off, xv);
#pragma unroll
    for (int j = 0; j < V; ++j) gv[j] = fmaf(a[j], gv[j], -fmaf(dd[j], xv[j], bb[j]));
    stv(g + off, gv);
  }
}

extern "C" int dwn_bn_bwd_apply(void* g, const void* x, const float* coef, const float* bcoef, long M, int C, int dtype,
                                void* stream) {
  const int V = dtype == DWN_DT_F32 ? 4 : 8;
  DWN_REQUIRE(C % V == 0, "dwn_bn_bwd_apply: C %% %d != 0", V);
  int cvc = dwn_largest_divisor_le(C / V, 64), ln = 256 / cvc;
  const int ny = (C / V) / cvc;
  int gx = (148 * 8 + ny - 1) / ny;
  dim3 grid(gx, ny), block(cvc * ln);
  if (dtype == DWN_DT_F32)
    bn_bwd_apply_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((float*)g, (const float*)x, coef, bcoef, M, C, cvc);
  else
    bn_bwd_apply_kernel<bf16><<<grid, block, 0, (cudaStream_t)stream>>>((bf16*)g, (const bf16*)x, coef, bcoef, M, C, cvc);
  DWN_LAUNCH_CHECK();
  return 0;
}

// out[i] = sum_z partial[z][i]   (split-K / per-CTA partial reductions); optional transposed channel layout
__global__ void reduce_rows_kernel(const float* __restrict__ partial, int Z, long n, float* __restrict__ out) {
  // fp64 accumulation: the split-K partials of the Gram matrix / weight gradients are long sums of like-signed terms
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double s4[4] = {0.0, 0.0, 0.0, 0.0};  // four independent rows in flight
    int z = 0;
    for (; z + 3 < Z; z += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] += (double)partial[(long)(z + u) * n + i];
    }
    for (; z < Z; ++z) s4[0] += (double)partial[(long)z * n + i];
    out[i] = (float)((s4[0] + s4[1]) + (s4[2] + s4[3]));
  }
}
extern "C" int dwn_reduce_rows(const float* partial, int Z, long n, float* out, void* stream) {
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  reduce_rows_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(partial, Z, n, out);
  DWN_LAUNCH_CHECK();
  return 0;
}

// depth-wise weight gradient from partial[P][NQ][C] quantities q0..q0+KK-1 -> dw[C][KK]
__global__ void __launch_bounds__(1024) dw_wgrad_finalize_kernel(const float* __restrict__ partial, int P, int NQ, int q0,
                                                                int KK, float* __restrict__ dw, int C) {
  __shared__ float s0[32][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, k = blockIdx.y;
  float a = 0.f;
  if (c < C) {
    float a4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent partial rows in flight
    int p = sl;
    for (; p + 96 < P; p += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[u] += partial[((long)(p + 32 * u) * NQ + q0 + k) * C + c];
    }
    for (; p < P; p += 32) a4[0] += partial[((long)p * NQ + q0 + k) * C + c];
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  s0[sl][cl] = a;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  double s = 0;
  for (int i = 0; i < 32; ++i) s += (double)s0[i][cl];
  dw[(long)c * KK + k] = (float)s;
}
extern "C" int dwn_dw_wgrad_finalize(const float* partial, int P, int NQ, int q0, int KK, float* dw, int C, void* stream) {
  dim3 grid((C + 31) / 32, KK);
  dw_wgrad_finalize_kernel<<<grid, 1024, 0, (cudaStream_t)stream>>>(partial, P, NQ, q0, KK, dw, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// stem backward: one pass over dX0 [M][C0] and the 5-channel input accumulates
//   partial[P][CIN+1][C0] = { G[c][k] = sum_m dy[m][c]*x[m][k] (k<CIN), S[c] = sum_m dy[m][c] }
// the BN backward and the weight gradient then follow analytically from the input moments.
// =================================================================================================
template <int CIN>
__global__ void stem_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                       float* __restrict__ partial, int plane, int M, int C0, int cqc, FastDiv dplane) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  float st[CIN + 1][4] = {};
#pragma unroll 2
  for (int m = blockIdx.x * ln + lane; m < M; m += gridDim.x * ln) {
    const int b = dplane.div(m), pos = m - b * plane;
    float g[4];
    ldq(dy + (long)m * C0 + c, g);
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
      const float xv = __ldg(&x[((long)b * CIN + k) * plane + pos]);
#pragma unroll
      for (int j = 0; j < 4; ++j) st[k][j] = fmaf(g[j], xv, st[k][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) st[CIN][j] += g[j];
  }
  block_reduce_channels<CIN + 1, 4>(st, smem, cqc, ln, partial + (long)blockIdx.x * (CIN + 1) * C0, C0,
                                    blockIdx.y * cqc * 4);
}

// one warp per output channel (lanes stride over the partial rows), double precision
__global__ void stem_bwd_finalize_kernel(const float* __restrict__ partial, int P, int cin, const double* __restrict__ mom,
                                         double count, const float* __restrict__ w, const float* __restrict__ coef,
                                         float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         int C0) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C0) return;
  double G[8], S = 0;
  for (int k = 0; k < cin; ++k) G[k] = 0;
  for (int p = lane; p < P; p += 32) {
    for (int k = 0; k < cin; ++k) G[k] += (double)partial[((long)p * (cin + 1) + k) * C0 + c];
    S += (double)partial[((long)p * (cin + 1) + cin) * C0 + c];
  }
  for (int k = 0; k < cin; ++k) G[k] = warp_sum_d(G[k]);
  S = warp_sum_d(S);
  if (lane != 0) return;
  // second moments X2[j][k] from the packed upper triangle
  double X1[8], X2[8][8];
  int q = cin;
  for (int j = 0; j < cin; ++j) X1[j] = mom[j];
  for (int j = 0; j < cin; ++j)
    for (int k = j; k < cin; ++k) { X2[j][k] = mom[q]; X2[k][j] = mom[q]; ++q; }
  const double mean = coef[2 * C0 + c], rstd = coef[3 * C0 + c];
  const double gr = (double)coef[c];  // gamma*rstd
  double dot = 0;
  for (int k = 0; k < cin; ++k) dot += (double)w[c * cin + k] * G[k];  // sum_m dy*y_raw
  const double dg = rstd * (dot - mean * S);                           // sum dy*yhat
  dbeta[c] = (float)S;
  dgamma[c] = (float)dg;
  for (int k = 0; k < cin; ++k) {
    double wx = 0;  // sum_m y_raw*x_k = sum_j w_cj X2[j][k]
    for (int j = 0; j < cin; ++j) wx += (double)w[c * cin + j] * X2[j][k];
    const double yhx = rstd * (wx - mean * X1[k]);  // sum_m yhat*x_k
    dw[c * cin + k] = (float)(gr * (G[k] - (S / count) * X1[k] - (dg / count) * yhx));
  }
}

extern "C" int dwn_stem_bwd(const float* dy, const float* x, float* partial, int P, const double* mom, const float* w,
                            const float* coef, float* dw, float* dgamma, float* dbeta, int B, int cin, long plane, int C0,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int cqc = dwn_largest_divisor_le(C0 / 4, 64), ln = 256 / cqc;
  dim3 grid(P, (C0 / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * (cin + 1) * 4 * sizeof(float);
  const long M = (long)B * plane;
  switch (cin) {
#define CASE(N) case N: stem_bwd_reduce_kernel<N><<<grid, block, sm, st>>>(dy, x, partial, (int)plane, (int)M, C0, cqc, FastDiv((int)plane)); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: return dwn_fail("dwn_stem_bwd: in_channels=%d unsupported", cin);
  }
  DWN_LAUNCH_CHECK();
  stem_bwd_finalize_kernel<<<(C0 + 7) / 8, 256, 0, st>>>(partial, P, cin, mom, (double)M, w, coef, dw, dgamma, dbeta, C0);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// cortex layer backward (forward: dwn_cortex_out).  j = output channel, c = src(j) = conv channel.
//   partial[J][4][O] = { [c] sum dyh, [c] sum dyh*yhat, [j] sum dOut, [j] sum dOut*xhat_sc }
// =================================================================================================
template <typename T>
__global__ void cortex_bwd_reduce_kernel(const float* __restrict__ dOut, const T* __restrict__ y,
                                         const float* __restrict__ coef, const float* __restrict__ dp,
                                         const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                         float* __restrict__ partial, int M, int Tn, int I, int O, int G) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= O) return;
  const int per = O / G;
  const int c = (j % G) * per + j / G;
  const float sc = coef[c], sh = coef[O + c], mu = coef[2 * O + c], rs = coef[3 * O + c];
  const float ms = coef_sc[2 * O + j], rss = coef_sc[3 * O + j];
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  for (int m = blockIdx.y; m < M; m += gridDim.y) {
    const float g = dOut[(long)m * O + j];
    const float yv = ld1<T>(y + (long)m * O + c);
    const float d = (dp ? dp[m / Tn] : 1.0f) * g * silu_grad_t<T>(fmaf(yv, sc, sh));
    q0 += d;
    q1 += d * ((yv - mu) * rs);
    q2 += g;
    q3 += g * ((xin[(long)m * I + (j % I)] - ms) * rss);
  }
  float* p = partial + (long)blockIdx.y * 4 * O;
  p[c] = q0;
  p[O + c] = q1;
  p[2 * O + j] = q2;
  p[3 * O + j] = q3;
}

template <typename T>
__global__ void cortex_bwd_dy_kernel(const float* __restrict__ dOut, const T* __restrict__ y,
                                     const float* __restrict__ coef, const float* __restrict__ bcoef,
                                     const float* __restrict__ dp, T* __restrict__ dY, int M, int Tn, int O, int G) {
  const int per = O / G;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)M * O; i += (long)gridDim.x * blockDim.x) {
    const int m = (int)(i / O), c = (int)(i % O);
    const int j = (c % per) * G + c / per;  // inverse shuffle
    const float yv = ld1<T>(y + i);
    const float d = (dp ? dp[m / Tn] : 1.0f) * dOut[(long)m * O + j] * silu_grad_t<T>(fmaf(yv, coef[c], coef[O + c]));
    const float yh = (yv - coef[2 * O + c]) * coef[3 * O + c];
    st1<T>(dY + i, coef[c] * (d - bcoef[c] - yh * bcoef[O + c]));
  }
}

__global__ void cortex_in_bwd_kernel(const float* __restrict__ dXc, const float* __restrict__ dOut,
                                     const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                     const float* __restrict__ bcoef_sc, float* __restrict__ dX, int M, int I, int O) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)M * I; i += (long)gridDim.x * blockDim.x) {
    const int m = (int)(i / I), k = (int)(i % I);
    float s = dXc[i];
    const float x = xin[i];
    for (int j = k; j < O; j += I) {
      const float xh = (x - coef_sc[2 * O + j]) * coef_sc[3 * O + j];
      s += coef_sc[j] * (dOut[(long)m * O + j] - bcoef_sc[j] - xh * bcoef_sc[O + j]);
    }
    dX[i] = s;
  }
}

extern "C" int dwn_cortex_bwd_reduce(const float* dOut, const void* y, const float* coef, const float* dp,
                                     const float* xin, const float* coef_sc, float* partial, int J, int M, int Tn, int I,
                                     int O, int G, int dtype, void* stream) {
  dim3 grid((O + 127) / 128, J);
  if (dtype == DWN_DT_F32)
    cortex_bwd_reduce_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(dOut, (const float*)y, coef, dp, xin, coef_sc,
                                                                            partial, M, Tn, I, O, G);
  else
    cortex_bwd_reduce_kernel<bf16><<<grid, 128, 0, (cudaStream_t)stream>>>(dOut, (const bf16*)y, coef, dp, xin, coef_sc,
                                                                           partial, M, Tn, I, O, G);
  DWN_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwn_cortex_bwd_dy(const float* dOut, const void* y, const float* coef, const float* bcoef, const float* dp,
                                 void* dY, int M, int Tn, int O, int G, int dtype, void* stream) {
  long n = (long)M * O;
  int gx = (int)((n + 255) / 256);
  if (gx > 2048) gx = 2048;
  if (dtype == DWN_DT_F32)
    cortex_bwd_dy_kernel<float><<<gx, 256, 0, (cudaStream_t)stream>>>(dOut, (const float*)y, coef, bcoef, dp, (float*)dY,
                                                                      M, Tn, O, G);
  else
    cortex_bwd_dy_kernel<bf16><<<gx, 256, 0, (cudaStream_t)stream>>>(dOut, (const bf16*)y, coef, bcoef, dp, (bf16*)dY, M,
                                                                     Tn, O, G);
  DWN_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwn_cortex_in_bwd(const float* dXc, const float* dOut, const float* xin, const float* coef_sc,
                                 const float* bcoef_sc, float* dX, int M, int I, int O, void* stream) {
  long n = (long)M * I;
  int gx = (int)((n + 255) / 256);
  if (gx > 2048) gx = 2048;
  cortex_in_bwd_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(dXc, dOut, xin, coef_sc, bcoef_sc, dX, M, I, O);
  DWN_LAUNCH_CHECK();
  return 0;
}
