// Forward kernels of the DwiseNeuro core / cortex glue (channels-last, fused BN+SiLU on load).
// Reference semantics: /root/reference/src/models/dwiseneuro.py (line numbers cited per kernel).
#include "dwn_common.cuh"
#include "dwn_reduce.cuh"
#include "dwn_bulk.cuh"
#include "dwn_sdw_v3.cuh"
#include "dwn_sdw_fwd_tma.cuh"
#include <cstdlib>

// =================================================================================================
// input moments: sums and second moments of the 5 input channels (NCDHW fp32 input).
// Used to derive the stem's train-mode BatchNorm statistics analytically (stem conv is linear):
//   mean_c = w_c . mu_x ; var_c = w_c^T Cov_x w_c     (dwiseneuro.py:306-309)
// partial layout: [grid][NM] doubles, NM = CIN + CIN*(CIN+1)/2  (sum_k, then upper-tri sum_jk j<=k)
// =================================================================================================
template <int CIN>
__global__ void __launch_bounds__(256) input_moments_kernel(const float* __restrict__ x, long plane, long total,
                                                           double* __restrict__ partial) {
  constexpr int NM = CIN + CIN * (CIN + 1) / 2;
  float acc[NM];
  double dacc[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) { acc[i] = 0.f; dacc[i] = 0.0; }
  int it = 0;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long b = idx / plane, pos = idx - b * plane;
    float v[CIN];
#pragma unroll
    for (int k = 0; k < CIN; ++k) v[k] = x[(b * CIN + k) * plane + pos];
    int q = CIN;
#pragma unroll
    for (int j = 0; j < CIN; ++j) {
      acc[j] += v[j];
#pragma unroll
      for (int k = j; k < CIN; ++k) acc[q++] += v[j] * v[k];
    }
    if (++it == 32) {  // flush to double to keep fp32 partial sums short
#pragma unroll
      for (int i = 0; i < NM; ++i) { dacc[i] += (double)acc[i]; acc[i] = 0.f; }
      it = 0;
    }
  }
  __shared__ double red[8][NM];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    double s = warp_sum_d(dacc[i] + (double)acc[i]);
    if (lane == 0) red[wid][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NM) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(long)blockIdx.x * NM + threadIdx.x] = s;
  }
}

// moments[NM] = sum over partial rows
__global__ void moments_finalize_kernel(const double* __restrict__ partial, int P, int NM, double* __restrict__ mom) {
  int i = threadIdx.x;
  if (i < NM) {
    double s = 0;
    for (int p = 0; p < P; ++p) s += partial[(long)p * NM + i];
    mom[i] = s;
  }
}

extern "C" int dwn_input_moments(const float* x, int B, int cin, long plane, double* partial, int P, double* mom,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  long total = (long)B * plane;
  int nm = cin + cin * (cin + 1) / 2;
  switch (cin) {
#define CASE(N) case N: input_moments_kernel<N><<<P, 256, 0, st>>>(x, plane, total, partial); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: return dwn_fail("dwn_input_moments: in_channels=%d unsupported (1..8)", cin);
  }
  DWN_LAUNCH_CHECK();
  moments_finalize_kernel<<<1, 64, 0, st>>>(partial, P, nm, mom);
  DWN_LAUNCH_CHECK();
  return 0;
}

// stem BN coefficients from the input moments (train) — one thread per output channel.
// Updates running stats exactly like nn.BatchNorm3d (momentum 0.1, unbiased var) (dwiseneuro.py:9-22).
__global__ void stem_coef_kernel(const double* __restrict__ mom, int cin, double count, const float* __restrict__ w,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ rmean, float* __restrict__ rvar, long long* __restrict__ nbt,
                                 float momentum, float eps, float* __restrict__ coef, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  double mu[8];
  for (int k = 0; k < cin; ++k) mu[k] = mom[k] / count;
  double mean = 0;
  for (int k = 0; k < cin; ++k) mean += (double)w[c * cin + k] * mu[k];
  double var = 0;
  int q = cin;
  for (int j = 0; j < cin; ++j)
    for (int k = j; k < cin; ++k) {
      double cov = mom[q++] / count - mu[j] * mu[k];
      double ww = (double)w[c * cin + j] * (double)w[c * cin + k];
      var += (j == k ? 1.0 : 2.0) * ww * cov;
    }
  if (var < 0) var = 0;
  double rstd = 1.0 / sqrt(var + (double)eps);
  float sc = (float)((double)gamma[c] * rstd);
  coef[c] = sc;
  coef[C + c] = (float)((double)beta[c] - mean * (double)gamma[c] * rstd);
  coef[2 * C + c] = (float)mean;
  coef[3 * C + c] = (float)rstd;
  if (rmean) {
    double unb = count > 1 ? var * count / (count - 1.0) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

extern "C" int dwn_stem_coef(const double* mom, int cin, double count, const float* w, const float* gamma,
                             const float* beta, float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                             float* coef, int C, void* stream) {
  stem_coef_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mom, cin, count, w, gamma, beta, rmean, rvar, nbt,
                                                                      momentum, eps, coef, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// generic BN finalize: partial[P][2][C] (sum, sumsq) -> coef[4][C]; eval mode uses running stats.
// =================================================================================================
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int P, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ rmean, float* __restrict__ rvar,
                                                          long long* __restrict__ nbt, float momentum, float eps,
                                                          int training, float* __restrict__ coef, int C, int Cp,
                                                          int NQ) {
  // block = (8 channels, 128 row slices), see finalize_colsum2; partial has Cp channels, channel c reads column
  // c % Cp (cyclic channel tiling of the shortcut, dwiseneuro.py:130-132)
  __shared__ double s_red[512];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt && training) *nbt += 1;
  double mean, var;
  if (training) {
    double s, q;
    finalize_colsum2(partial, P, NQ, 0, 1, Cp, c < C ? c % Cp : 0, c < C, s_red, s, q);
    if (threadIdx.x >= 8 || c >= C) return;
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0) var = 0;
    if (rmean) {
      double unb = count > 1 ? var * count / (count - 1.0) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  } else {
    if (threadIdx.x >= 8 || c >= C) return;
    mean = rmean[c];
    var = rvar[c];
  }
  double rstd = 1.0 / sqrt(var + (double)eps);
  double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
  coef[c] = (float)(g * rstd);
  coef[C + c] = (float)(b - mean * g * rstd);
  coef[2 * C + c] = (float)mean;
  coef[3 * C + c] = (float)rstd;
}

extern "C" int dwn_bn_finalize(const float* partial, int P, double count, const float* gamma, const float* beta,
                               float* rmean, float* rvar, long long* nbt, float momentum, float eps, int training,
                               float* coef, int C, int Cp, int NQ, void* stream) {
  bn_finalize_kernel<<<(C + 7) / 8, 1024, 0, (cudaStream_t)stream>>>(partial, P, count, gamma, beta, rmean, rvar, nbt,
                                                                      momentum, eps, training, coef, C, Cp > 0 ? Cp : C,
                                                                      NQ > 0 ? NQ : 2);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// stem: Conv3d(cin -> C0, 1x1x1, no bias) + BN + positional encoding of block 0, writes the
// channels-last trunk X0[M][C0] (fp32, + optional bf16 copy) and the partial stats of the
// (strided) shortcut of block 0.       (dwiseneuro.py:306-309, 147-192, 125-134)
// =================================================================================================
template <int CIN>
__global__ void stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ coef,
                                const float* __restrict__ pe_t, const float* __restrict__ pe_h,
                                const float* __restrict__ pe_w, float* __restrict__ out, bf16* __restrict__ out_bf,
                                float* __restrict__ partial, NearestMap nh, NearestMap nw, int B, int Tn, int H, int W,
                                int C0, int cqc, FastDiv dw, FastDiv dh, FastDiv dplane) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  float wr[4][CIN], sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = coef[c + j];
    sh[j] = coef[C0 + c + j];
#pragma unroll
    for (int k = 0; k < CIN; ++k) wr[j][k] = w[(c + j) * CIN + k];
  }
  float st[3][4] = {};
  const int plane = Tn * H * W, M = B * plane;
  // Four positions per trip: all 4*CIN input scalars are loaded before the first use, so a trip costs one load latency
  // instead of four (the kernel is write-bound but was limited by these dependent loads).  A per-thread cp.async
  // prefetch was tried and was 2x slower: the 16 threads of a position share the scalars, which plain loads
  // broadcast but LDGSTS copies once per thread.
  const int m0 = blockIdx.x * ln + lane, mstep = gridDim.x * ln;
  for (int mb = m0; mb < M; mb += 4 * mstep) {
    float xv[4][CIN];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = mb + u * mstep;
      if (m < M) {
        const int b = dplane.div(m), pos = m - b * plane;
#pragma unroll
        for (int k = 0; k < CIN; ++k) xv[u][k] = __ldg(&x[((long)b * CIN + k) * plane + pos]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int m = mb + u * mstep;
      if (m >= M) break;
      const int b = dplane.div(m), pos = m - b * plane;
      const int wq = dw.mod(pos), hq = dh.mod(dw.div(pos)), tq = dh.div(dw.div(pos));
      float pt[4], ph[4], pw[4], o[4];
      ldq(pe_t + tq * C0 + c, pt);
      ldq(pe_h + hq * C0 + c, ph);
      ldq(pe_w + wq * C0 + c, pw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < CIN; ++k) a = fmaf(wr[j][k], xv[u][k], a);
        o[j] = fmaf(a, sc[j], sh[j]) + ((pt[j] + ph[j]) + pw[j]);
      }
      stq(out + (long)m * C0 + c, o);
      if (out_bf) stq(out_bf + (long)m * C0 + c, o);
      // Gram statistics describe E = Xb W^T, so the column sum is taken over the bf16-ROUNDED operand when there is one
      // (inputs with many repeated values round coherently: mean(Xb) - mean(X) does not average out)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[2][j] += out_bf ? __bfloat162float(__float2bfloat16_rn(o[j])) : o[j];
      if (partial && nh.dst(hq) >= 0 && nw.dst(wq) >= 0) {  // positions the next block's shortcut gathers
#pragma unroll
        for (int j = 0; j < 4; ++j) { st[0][j] += o[j]; st[1][j] = fmaf(o[j], o[j], st[1][j]); }
      }
    }
  }
  if (partial)  // [P][3][C0]: strided sum / sumsq (next shortcut BN) and the full column sum (Gram statistics)
    block_reduce_channels<3, 4>(st, smem, cqc, ln, partial + (long)blockIdx.x * 3 * C0, C0, blockIdx.y * cqc * 4);
}

extern "C" int dwn_stem_fwd(const float* x, const float* w, const float* coef, const float* pe_t, const float* pe_h,
                            const float* pe_w, float* out, void* out_bf, float* partial, int P, int next_stride, int B,
                            int cin, int Tn, int H, int W, int C0, void* stream) {
  DWN_REQUIRE(C0 % 4 == 0, "dwn_stem_fwd: C0 %% 4 != 0");
  int cqc = dwn_largest_divisor_le(C0 / 4, 64);
  int ln = 256 / cqc;
  dim3 grid(P, (C0 / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * 3 * 4 * sizeof(float);
  switch (cin) {
#define CASE(N)                                                                                                   \
  case N:                                                                                                         \
    stem_fwd_kernel<N><<<grid, block, sm, (cudaStream_t)stream>>>(x, w, coef, pe_t, pe_h, pe_w, out, (bf16*)out_bf, \
                                                                  partial, NearestMap(H, dwn_ceil_div(H, next_stride)),  \
                                                                  NearestMap(W, dwn_ceil_div(W, next_stride)), B, Tn, H, W, C0, cqc, \
                                                                  FastDiv(W), FastDiv(H), FastDiv(Tn * H * W));        \
    break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: return dwn_fail("dwn_stem_fwd: in_channels=%d unsupported (1..8)", cin);
  }
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// spatial depth-wise (1,3,3) conv, stride S, pad 1, with BN+SiLU of the producer applied on load.
//   in  E_raw [NP][H][W][C]   (pre-BN output of conv_pw)          (dwiseneuro.py:90-102)
//   out S_raw [NP][Ho][Wo][C] (pre-BN), partial[P][2][C] = column sum / sumsq of the stored values
// CTA tile: one (b,t) plane, THO output rows, all W, CC channels; activated halo tile in smem (fp32).
// =================================================================================================
template <typename T, int S, int THO>
__global__ void __launch_bounds__(256, 2)
sdw_fwd_kernel(const T* __restrict__ in, const float* __restrict__ coef, const float* __restrict__ wgt,
               T* __restrict__ out, float* __restrict__ partial, int NP, int H, int W, int C, int CC, int nchunks,
               int wsh, int cvsh) {
  // grid is 1-D with the channel chunk fastest: CTAs that share a (plane, band) tile run together, so every
  // 128-byte line of E_raw is consumed while it is L2 resident.  wsh/cvsh = log2(W), log2(CC/V) or -1.
  constexpr int V = VecT<T>::V;
  constexpr int NR = (THO - 1) * S + 3;
  extern __shared__ float tile[];
  const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S, WP = W + 2;  // conv k=3 pad=1 stride S: ceil(size / S)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int c0 = chunk * CC;
  // ---- per-thread constants for the load phase (fixed channel vector)
  const int cvn = CC / V;
  const int lcv = tid % cvn;
  float lp0[V], lp1[V];
#pragma unroll
  for (int j = 0; j < V; ++j) BnSilu<T>::prep(coef[c0 + lcv * V + j], coef[C + c0 + lcv * V + j], lp0[j], lp1[j]);
  // ---- per-thread constants for the compute phase
  const int cqn = CC / 4;
  const int cq = tid % cqn, wo = tid / cqn;
  float wr[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) wr[k][j] = wgt[(c0 + cq * 4 + j) * 9 + k];
  float st[2][4] = {};
  // zero halo columns once
  for (int i = tid; i < NR * 2 * CC; i += nthr) {
    int r = i / (2 * CC), rem = i % (2 * CC);
    int col = (rem / CC) ? (W + 1) : 0;
    tile[(r * WP + col) * CC + (rem % CC)] = 0.f;
  }
  const int nb = Ho / THO;
  const int ntiles = NP * nb;
  const int nvec = NR * W * cvn;
  for (int t = worker; t < ntiles; t += nworkers) {
    const int p = t / nb, ho0 = (t % nb) * THO;
    const int hi0 = ho0 * S - 1;
    __syncthreads();  // previous compute done before overwriting the tile
    // batches of UB vectors per thread: all global loads of a batch are issued before the first use, so a tile costs
    // nvec / (nthr * UB) exposed load latencies instead of nvec / (2 * nthr) (the fp32 launches ran at 1.5 TB/s)
    constexpr int UB = 32 / V;
    for (int i0 = tid; i0 < nvec; i0 += nthr * UB) {
      float v[UB][V];
      int dsto[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * nthr;
        dsto[u] = -1;
        if (i < nvec) {
          int r, wq;
          if (wsh >= 0) { r = i >> (wsh + cvsh); wq = (i >> cvsh) & (W - 1); }
          else { r = i / (W * cvn); wq = (i / cvn) % W; }
          const int hi = hi0 + r;
          dsto[u] = (r * WP + wq + 1) * CC + lcv * V;
          if (hi >= 0 && hi < H) {
            ldv(in + (((long)p * H + hi) * W + wq) * C + c0 + lcv * V, v[u]);
          } else {
            dsto[u] = -2 - dsto[u];  // outside the image: store zeros (padding applies after the activation)
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (dsto[u] == -1) continue;
        const bool inside = dsto[u] >= 0;
        float* dst = tile + (inside ? dsto[u] : -2 - dsto[u]);
#pragma unroll
        for (int j = 0; j < V; ++j) v[u][j] = inside ? BnSilu<T>::act(v[u][j], lp0[j], lp1[j]) : 0.f;
#pragma unroll
        for (int j = 0; j < V; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(v[u][j], v[u][j + 1], v[u][j + 2], v[u][j + 3]);
      }
    }
    __syncthreads();
    float R[3][3][4];
    auto load_row = [&](int r) {
      const float* src = tile + ((r * WP + wo * S) * CC + cq * 4);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float4 q = *reinterpret_cast<const float4*>(src + kw * CC);
        R[r % 3][kw][0] = q.x; R[r % 3][kw][1] = q.y; R[r % 3][kw][2] = q.z; R[r % 3][kw][3] = q.w;
      }
    };
    if (S == 1) { load_row(0); load_row(1); } else { load_row(0); }
#pragma unroll
    for (int hl = 0; hl < THO; ++hl) {
      if (S == 1) { load_row(hl + 2); } else { load_row(2 * hl + 1); load_row(2 * hl + 2); }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(R[(hl * S + kh) % 3][kw][j], wr[kh * 3 + kw][j], acc[j]);
      stq(out + ((((long)p * Ho + ho0 + hl) * Wo + wo) * C + c0 + cq * 4), acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float r = rnd<T>(acc[j]);
        st[0][j] += r;
        st[1][j] = fmaf(r, r, st[1][j]);
      }
    }
  }
  __syncthreads();
  if (partial) block_reduce_channels<2, 4>(st, tile, cqn, Wo, partial + (long)worker * 2 * C, C, c0);
}

static inline int ilog2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

template <typename T, int S>
static int sdw_fwd_launch(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP,
                          int H, int W, int C, cudaStream_t st) {
  constexpr int V = VecT<T>::V;
  const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
  int CC = 128;  // largest power of two that divides C with (CC/4)*Wo <= 256 threads
  while (CC >= 8 && (C % CC != 0 || (CC / 4) * Wo > 256)) CC /= 2;
  DWN_REQUIRE(CC >= 8 && C % CC == 0 && (CC / 4) * Wo <= 256, "dwn_sdw_fwd: unsupported C=%d W=%d", C, W);
  int THO = (S == 1 && Ho % 8 == 0) ? 8 : (Ho % 4 == 0 ? 4 : (Ho % 2 == 0 ? 2 : 1));
  const int NR = (THO - 1) * S + 3;
  size_t sm = (size_t)NR * (W + 2) * CC * sizeof(float);
  size_t sm_red = (size_t)(CC / 4) * Wo * 2 * 4 * sizeof(float);
  if (sm_red > sm) sm = sm_red;
  const int nchunks = C / CC;
  int wsh = ilog2_exact(W), cvsh = ilog2_exact(CC / V);
  if (wsh < 0 || cvsh < 0) wsh = cvsh = -1;
  dim3 grid(P * nchunks), block((CC / 4) * Wo);
#define LAUNCH(THO_)                                                                                             \
  {                                                                                                              \
    auto k = sdw_fwd_kernel<T, S, THO_>;                                                                         \
    if (sm > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);           \
    k<<<grid, block, sm, st>>>((const T*)in, coef, wgt, (T*)out, partial, NP, H, W, C, CC, nchunks, wsh, cvsh);  \
  }
  switch (THO) {
    case 8: LAUNCH(8) break;
    case 4: LAUNCH(4) break;
    case 2: LAUNCH(2) break;
    default: LAUNCH(1) break;
  }
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}

// pipelined bf16 path (cp.async double buffering + FFMA2); returns 1 if the shape is not eligible
template <int S>
static int sdw_fwd_v3_launch(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP,
                             int H, int W, int C, cudaStream_t st) {
  if (H % S != 0 || W % S != 0) return 1;  // ceil-sized outputs take the generic kernel
  const int Ho = H / S, Wo = W / S;
  int CC = 1024 / Wo;
  if (CC > 128) CC = 128;
  while (CC >= 8 && (C % CC != 0)) CC /= 2;
  if (CC < 8 || C % CC != 0 || (CC / 4) * Wo != 256) return 1;
  const int cvsh = ilog2_exact(CC / 8);
  const int vpr = W * (CC / 8);  // 16-byte vectors per tile row
  if (cvsh < 0 || (vpr != 128 && vpr != 256)) return 1;
  const int rpi = 256 / vpr;
  const int THO = S == 1 ? (Ho % 8 == 0 ? 8 : (Ho % 4 == 0 ? 4 : 0)) : (Ho % 2 == 0 ? 2 : 0);
  if (THO == 0) return 1;
  const int NR = (THO - 1) * S + 3;
  if (NR % rpi != 0) return 1;
  const int nbsh = ilog2_exact(Ho / THO);
  if (nbsh < 0) return 1;
  const size_t nvec = (size_t)NR * vpr;
  size_t sm = 3 * nvec * 16 + ((size_t)NR * (W + 2) + 2) * CC * sizeof(bf16);  // raw ring + bf16 activated tile
  if (sm < (size_t)256 * 2 * 4 * sizeof(float)) sm = (size_t)256 * 2 * 4 * sizeof(float);
  const int nchunks = C / CC;
  dim3 grid(P * nchunks), block(256);
#define LAUNCH(THO_, RPI_, CC_)                                                                                \
  {                                                                                                            \
    auto k = sdw_fwd_v3_kernel<S, THO_, RPI_, CC_>;                                                            \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                             \
    k<<<grid, block, sm, st>>>((const bf16*)in, coef, wgt, (bf16*)out, partial, NP, H, C, nchunks, nbsh);      \
  }
  // vpr = W*CC/8 = S*128 -> rpi = 2 for stride 1, 1 for stride 2
  if constexpr (S == 1) {
    if (rpi != 2) return 1;
    if (THO == 8) {
      if (CC == 32) LAUNCH(8, 2, 32) else if (CC == 64) LAUNCH(8, 2, 64) else if (CC == 128) LAUNCH(8, 2, 128) else return 1;
    } else {
      if (CC == 32) LAUNCH(4, 2, 32) else if (CC == 64) LAUNCH(4, 2, 64) else if (CC == 128) LAUNCH(4, 2, 128) else return 1;
    }
  } else {
    if (rpi != 1) return 1;
    if (CC == 32) LAUNCH(2, 1, 32) else if (CC == 64) LAUNCH(2, 1, 64) else if (CC == 128) LAUNCH(2, 1, 128) else return 1;
  }
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_sdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int NP,
                           int H, int W, int C, int stride, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DWN_REQUIRE(stride == 1 || stride == 2, "dwn_sdw_fwd: stride %d unsupported", stride);
  if (dtype == DWN_DT_F32)
    return stride == 1 ? sdw_fwd_launch<float, 1>(in, coef, wgt, out, partial, P, NP, H, W, C, st)
                       : sdw_fwd_launch<float, 2>(in, coef, wgt, out, partial, P, NP, H, W, C, st);
  // TMA-staged kernels first (DWN_SDW_FWD_TMA=0 selects the cp.async kernels, DWN_SDW_FWD_THO the rows per item: A/B runs)
  const char* env_tma = getenv("DWN_SDW_FWD_TMA");
  const char* env_tho = getenv("DWN_SDW_FWD_THO");
  int rc = 1;
  if (!env_tma || atoi(env_tma) != 0) {
    const int tho = env_tho ? atoi(env_tho) : 0;
    rc = stride == 1 ? sdw_fwd_v6_launch<1>(in, coef, wgt, out, partial, P, NP, H, W, C, tho, st)
                     : sdw_fwd_v6_launch<2>(in, coef, wgt, out, partial, P, NP, H, W, C, tho, st);
    if (rc <= 0) return rc;
  }
  rc = stride == 1 ? sdw_fwd_v3_launch<1>(in, coef, wgt, out, partial, P, NP, H, W, C, st)
                   : sdw_fwd_v3_launch<2>(in, coef, wgt, out, partial, P, NP, H, W, C, st);
  if (rc <= 0) return rc;
  return stride == 1 ? sdw_fwd_launch<bf16, 1>(in, coef, wgt, out, partial, P, NP, H, W, C, st)
                     : sdw_fwd_launch<bf16, 2>(in, coef, wgt, out, partial, P, NP, H, W, C, st);
}

// =================================================================================================
// temporal depth-wise (5,1,1) conv, pad 2, BN+SiLU of the producer applied on load.
//   in S_raw [B][T][HW][C] -> out Tm_raw (same shape), partial[P][2][C]     (dwiseneuro.py:105-111)
// thread = (channel quad, position); the whole T column lives in registers (TT = compile-time T).
// =================================================================================================
template <typename T, int TT>
__global__ void __launch_bounds__(256, 2)
tdw_fwd_kernel(const T* __restrict__ in, const float* __restrict__ coef, const float* __restrict__ wgt,
               T* __restrict__ out, float* __restrict__ partial, int B, int Tn, int HW, int C, int cqc, FastDiv dhw) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  f32x2 p0[2], p1[2], w2[5][2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float a0, a1, b0, b1;
    BnSilu<T>::prep(coef[c + 2 * h], coef[C + c + 2 * h], a0, b0);
    BnSilu<T>::prep(coef[c + 2 * h + 1], coef[C + c + 2 * h + 1], a1, b1);
    p0[h] = pk2(a0, a1);
    p1[h] = pk2(b0, b1);
#pragma unroll
    for (int k = 0; k < 5; ++k) w2[k][h] = pk2(wgt[(c + 2 * h) * 5 + k], wgt[(c + 2 * h + 1) * 5 + k]);
  }
  f32x2 st2[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
  const int npos = B * HW;
  const long tstride = (long)HW * C;
  const long bstride = (long)Tn * tstride;
  for (int pos = blockIdx.x * ln + lane; pos < npos; pos += gridDim.x * ln) {
    const int b = dhw.div(pos), hw = pos - b * HW;
    const T* ip = in + b * bstride + (long)hw * C + c;
    T* op = out + b * bstride + (long)hw * C + c;
    if (TT > 0) {
      f32x2 a[TT > 0 ? TT : 1][2];
#pragma unroll
      for (int t = 0; t < TT; ++t) ldq2(ip + t * tstride, a[t]);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        a[t][0] = bnsilu2<T>(a[t][0], p0[0], p1[0]);
        a[t][1] = bnsilu2<T>(a[t][1], p0[1], p1[1]);
      }
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        f32x2 o[2] = {0ull, 0ull};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int ts = t + k - 2;
          if (ts >= 0 && ts < TT) {
            ffma2(o[0], a[ts][0], w2[k][0]);
            ffma2(o[1], a[ts][1], w2[k][1]);
          }
        }
        stq2(op + t * tstride, o);
        fadd2(st2[0][0], o[0]);
        fadd2(st2[0][1], o[1]);
        ffma2(st2[1][0], o[0], o[0]);
        ffma2(st2[1][1], o[1], o[1]);
      }
    } else {
      f32x2 win[5][2];
#pragma unroll
      for (int k = 0; k < 5; ++k) { win[k][0] = 0ull; win[k][1] = 0ull; }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t < Tn) {
          f32x2 v[2];
          ldq2(ip + t * tstride, v);
          win[3 + t][0] = bnsilu2<T>(v[0], p0[0], p1[0]);
          win[3 + t][1] = bnsilu2<T>(v[1], p0[1], p1[1]);
        }
      }
      for (int t = 0; t < Tn; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { win[k][0] = win[k + 1][0]; win[k][1] = win[k + 1][1]; }
        if (t + 2 < Tn) {
          f32x2 v[2];
          ldq2(ip + (t + 2) * tstride, v);
          win[4][0] = bnsilu2<T>(v[0], p0[0], p1[0]);
          win[4][1] = bnsilu2<T>(v[1], p0[1], p1[1]);
        } else {
          win[4][0] = 0ull; win[4][1] = 0ull;
        }
        f32x2 o[2] = {0ull, 0ull};
#pragma unroll
        for (int k = 0; k < 5; ++k) { ffma2(o[0], win[k][0], w2[k][0]); ffma2(o[1], win[k][1], w2[k][1]); }
        stq2(op + t * tstride, o);
        fadd2(st2[0][0], o[0]);
        fadd2(st2[0][1], o[1]);
        ffma2(st2[1][0], o[0], o[0]);
        ffma2(st2[1][1], o[1], o[1]);
      }
    }
  }
  if (partial) {
    float st[2][4];
    upk2(st2[0][0], st[0][0], st[0][1]); upk2(st2[0][1], st[0][2], st[0][3]);
    upk2(st2[1][0], st[1][0], st[1][1]); upk2(st2[1][1], st[1][2], st[1][3]);
    block_reduce_channels<2, 4>(st, smem, cqc, ln, partial + (long)blockIdx.x * 2 * C, C, blockIdx.y * cqc * 4);
  }
}

extern "C" int dwn_tdw_fwd(const void* in, const float* coef, const float* wgt, void* out, float* partial, int P, int B,
                           int Tn, int HW, int C, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DWN_REQUIRE(C % 4 == 0, "dwn_tdw_fwd: C %% 4 != 0");
  int cqc = dwn_largest_divisor_le(C / 4, 128);
  int ln = 256 / cqc;
  if (ln < 1) ln = 1;
  dim3 grid(P, (C / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * 2 * 4 * sizeof(float);
#define GO(TY, TTV) tdw_fwd_kernel<TY, TTV><<<grid, block, sm, st>>>((const TY*)in, coef, wgt, (TY*)out, partial, B, Tn, HW, C, cqc, FastDiv(HW))
  if (dtype == DWN_DT_F32) {
    if (Tn == 16) GO(float, 16); else GO(float, 0);
  } else {
    if (Tn == 16) GO(bf16, 16); else GO(bf16, 0);
  }
#undef GO
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// SE squeeze: a = SiLU(BN3(Tm_raw)) is materialised (the A operand of the projection GEMM) and
// pooled per (sample, channel).  partial[B][J][C]                        (dwiseneuro.py:38-39)
// =================================================================================================
template <typename T>
__global__ void se_pool_kernel(const T* __restrict__ in, const float* __restrict__ coef, T* __restrict__ act,
                               float* __restrict__ partial, int Nsp, int C, int cvc, long pstride) {
  // pstride > 0 (fp32 only): `act` receives the three bf16 planes of a (hi, mid, lo; dwn_split3 layout, planes pstride
  // elements apart) instead of fp32 values — the A operand of the fp32-accurate tensor-core projection GEMM
  constexpr int V = VecT<T>::V;
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cv = tid % cvc, lane = tid / cvc, ln = blockDim.x / cvc;
  const int c = (blockIdx.y * cvc + cv) * V;
  const int b = blockIdx.z;
  float sc[V], sh[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { sc[j] = coef[c + j]; sh[j] = coef[C + c + j]; }
  float st[1][V] = {};
  const long base = (long)b * Nsp * C + c;
  float q0[V], q1[V];
#pragma unroll
  for (int j = 0; j < V; ++j) BnSilu<T>::prep(sc[j], sh[j], q0[j], q1[j]);
#pragma unroll 4
  for (int r = blockIdx.x * ln + lane; r < Nsp; r += gridDim.x * ln) {
    float v[V];
    ldv(in + base + (long)r * C, v);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = BnSilu<T>::act(v[j], q0[j], q1[j]);
    if (sizeof(T) == 4 && pstride > 0) {
      bf16* pl = reinterpret_cast<bf16*>(act) + base + (long)r * C;
      float h[4], m[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
        const float r1 = v[j] - h[j];
        m[j] = __bfloat162float(__float2bfloat16_rn(r1));
        l[j] = r1 - m[j];
      }
      *reinterpret_cast<uint2*>(pl) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      *reinterpret_cast<uint2*>(pl + pstride) = make_uint2(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]));
      *reinterpret_cast<uint2*>(pl + 2 * pstride) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    } else {
      stv(act + base + (long)r * C, v);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) st[0][j] += v[j];
  }
  block_reduce_channels<1, V>(st, smem, cvc, ln, partial + ((long)b * gridDim.x + blockIdx.x) * C, C,
                              blockIdx.y * cvc * V);
}

// Bulk-staged variant (bf16, C/8 <= 256): CTA (j, b) streams contiguous chunks of R rows x C channels of Tm_raw through
// a 3-stage cp.async.bulk ring (see dwn_bulk.cuh), activates them from shared memory and writes A with 16-byte stores.
__global__ void __launch_bounds__(256, 3)
se_pool_bulk_kernel(const bf16* __restrict__ in, const float* __restrict__ coef, bf16* __restrict__ act,
                    float* __restrict__ partial, int Nsp, int C, int cvc, int R) {
  constexpr int NS = 3;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int tid = threadIdx.x;
  const size_t chunk_elems = (size_t)R * C;
  bf16* buf = reinterpret_cast<bf16*>(smraw);  // [NS][R*C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + (size_t)NS * chunk_elems * sizeof(bf16));
  const int cv = tid % cvc, lane = tid / cvc, ln = blockDim.x / cvc;
  const int c = cv * 8;
  const int b = blockIdx.y, j = blockIdx.x, J = gridDim.x;
  const int nch = (Nsp + R - 1) / R;
  const bf16* xbase = in + (long)b * Nsp * C;
  bf16* obase = act + (long)b * Nsp * C + c;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) bk_mbar_init(&bars[s], 1);
    bk_mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int ch, int stage) {  // one thread
    const int rows = min(R, Nsp - ch * R);
    const uint32_t bytes = (uint32_t)((size_t)rows * C * sizeof(bf16));
    bk_mbar_expect_tx(&bars[stage], bytes);
    bk_bulk_g2s(buf + (size_t)stage * chunk_elems, xbase + (long)ch * R * C, bytes, &bars[stage]);
  };
  if (tid == 0)
    for (int s = 0; s < NS; ++s)
      if (j + s * J < nch) issue(j + s * J, s);
  float q0[8], q1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) BnSilu<bf16>::prep(coef[c + e], coef[C + c + e], q0[e], q1[e]);
  float st[1][8] = {};
  int it = 0;
  for (int ch = j; ch < nch; ch += J, ++it) {
    const int stage = it % NS;
    bk_mbar_wait(&bars[stage], (uint32_t)((it / NS) & 1));
    const int rows = min(R, Nsp - ch * R);
    const bf16* xs = buf + (size_t)stage * chunk_elems + c;
    bf16* op = obase + (long)ch * R * C;
#pragma unroll 2
    for (int r = lane; r < rows; r += ln) {
      float v[8];
      ldv(xs + (size_t)r * C, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = BnSilu<bf16>::act(v[e], q0[e], q1[e]);
      stv(op + (long)r * C, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) st[0][e] += v[e];
    }
    __syncthreads();  // everybody is done with this stage
    if (tid == 0 && ch + NS * J < nch) issue(ch + NS * J, stage);
  }
  block_reduce_channels<1, 8>(st, reinterpret_cast<float*>(smraw), cvc, ln, partial + ((long)b * J + j) * C, C, 0);
}

extern "C" int dwn_se_pool(const void* in, const float* coef, void* act, float* partial, int J, int B, int Nsp, int C,
                           int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DWN_DT_BF16 && C % 8 == 0 && C / 8 <= 256) {
    const int cvc = C / 8, ln = 256 / cvc;
    int R = (int)(14336 / ((size_t)C * sizeof(bf16)));
    if (R < 1) R = 1;
    if (R > Nsp) R = Nsp;
    size_t sm = (size_t)3 * R * C * sizeof(bf16) + 64;
    const size_t sm_red = (size_t)cvc * ln * 8 * sizeof(float);
    if (sm_red > sm) sm = sm_red;
    cudaFuncSetAttribute(se_pool_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    se_pool_bulk_kernel<<<dim3(J, B), cvc * ln, sm, st>>>((const bf16*)in, coef, (bf16*)act, partial, Nsp, C, cvc, R);
    DWN_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == DWN_DT_F32 || dtype == 2) {  // 2: fp32 input, output = three bf16 planes (see se_pool_kernel)
    int cvc = dwn_largest_divisor_le(C / 4, 64), ln = 256 / cvc;
    dim3 grid(J, (C / 4) / cvc, B), block(cvc * ln);
    const long n = (long)B * Nsp * C;
    se_pool_kernel<float><<<grid, block, block.x * 4 * sizeof(float), st>>>((const float*)in, coef, (float*)act, partial,
                                                                           Nsp, C, cvc, dtype == 2 ? (n + 7) / 8 * 8 : 0);
  } else {
    DWN_REQUIRE(C % 8 == 0, "dwn_se_pool: C %% 8 != 0");
    int cvc = dwn_largest_divisor_le(C / 8, 64), ln = 256 / cvc;
    dim3 grid(J, (C / 8) / cvc, B), block(cvc * ln);
    se_pool_kernel<bf16><<<grid, block, block.x * 8 * sizeof(float), st>>>((const bf16*)in, coef, (bf16*)act, partial,
                                                                          Nsp, C, cvc, 0);
  }
  DWN_LAUNCH_CHECK();
  return 0;
}

// SE excitation MLP, one CTA (1024 threads) per sample (dwiseneuro.py:40-43):
// mean -> reduce(+b) -> SiLU -> expand(+b) -> sigmoid.  Loops are unrolled so that several loads are in flight.
__global__ void __launch_bounds__(1024) se_mlp_kernel(const float* __restrict__ partial, int J, float inv_n,
                                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2, const float* __restrict__ b2,
                                                     float* __restrict__ mean_out, float* __restrict__ hpre_out,
                                                     float* __restrict__ gate_out, int C, int RD) {
  extern __shared__ float sm[];  // mean[C], h[RD]
  float* s_mean = sm;
  float* s_h = sm + C;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j0 = 0; j0 < J; j0 += 16) {  // 16 independent loads in flight, same summation order
      float pv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pv[j] = (j0 + j < J) ? partial[((long)b * J + min(j0 + j, J - 1)) * C + c] : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) s += pv[j];
    }
    s *= inv_n;
    s_mean[c] = s;
    mean_out[(long)b * C + c] = s;
  }
  __syncthreads();
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  // global loads in explicit batches of 8 with select-predication: the plain unrolled loops waited for every load before
  // issuing the next one (35 us at C = 1792, pure latency)
  for (int r = wid; r < RD; r += nw) {
    float s = 0.f;
    for (int c0 = 0; c0 < C; c0 += 512) {
      float wv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + lane + 32 * j;
        wv[j] = c < C ? __ldg(&w1[(long)r * C + c]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s = fmaf(wv[j], s_mean[min(c0 + lane + 32 * j, C - 1)], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      s += b1[r];
      hpre_out[(long)b * RD + r] = s;
      s_h[r] = s / (1.0f + expf(-s));
    }
  }
  __syncthreads();
  // gate[c] = sigmoid(b2[c] + sum_r w2[c][r] h[r]): warp = 8 consecutive channels at a time (one contiguous 8 x RD block of
  // w2), lanes = r
  for (int cb = wid * 8; cb < C; cb += nw * 8) {
    float s8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s8[j] = 0.f;
    for (int r0 = 0; r0 < RD; r0 += 64) {
      float wa[8], wb[8];
      const int ra = r0 + lane, rb = r0 + 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool okc = cb + j < C;
        wa[j] = (okc && ra < RD) ? __ldg(&w2[(long)(cb + j) * RD + ra]) : 0.f;
        wb[j] = (okc && rb < RD) ? __ldg(&w2[(long)(cb + j) * RD + rb]) : 0.f;
      }
      const float ha = s_h[min(ra, RD - 1)], hb = s_h[min(rb, RD - 1)];
#pragma unroll
      for (int j = 0; j < 8; ++j) s8[j] = fmaf(wb[j], hb, fmaf(wa[j], ha, s8[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s8[j] = warp_sum(s8[j]);
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine = lane == j ? s8[j] : mine;
    if (lane < 8 && cb + lane < C) gate_out[(long)b * C + cb + lane] = 1.0f / (1.0f + expf(-(mine + b2[cb + lane])));
  }
}

extern "C" int dwn_se_mlp(const float* partial, int J, int Nsp, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* mean_out, float* hpre_out, float* gate_out, int B, int C, int RD,
                          void* stream) {
  se_mlp_kernel<<<B, 1024, (C + RD) * sizeof(float), (cudaStream_t)stream>>>(partial, J, 1.0f / (float)Nsp, w1, b1, w2, b2,
                                                                             mean_out, hpre_out, gate_out, C, RD);
  DWN_LAUNCH_CHECK();
  return 0;
}

// fold the SE gate into per-sample projection weights: Wb[b][n][k] = W[n][k] * gate[b][k]
template <typename T>
__global__ void fold_gate_kernel(const float* __restrict__ w, const float* __restrict__ gate, T* __restrict__ out, long NK,
                                 int K) {
  const int b = blockIdx.y;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < NK; i += (long)gridDim.x * blockDim.x * 4) {
    int k = (int)(i % K);
    float wv[4], g[4], o[4];
    ldq(w + i, wv);
    ldq(gate + (long)b * K + k, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = wv[j] * g[j];
    stq(out + (long)b * NK + i, o);
  }
}

extern "C" int dwn_fold_gate(const float* w, const float* gate, void* out, int B, int N, int K, int dtype, void* stream) {
  DWN_REQUIRE(K % 4 == 0, "dwn_fold_gate: K %% 4 != 0");
  long NK = (long)N * K;
  int gx = (int)((NK / 4 + 255) / 256);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, B);
  if (dtype == DWN_DT_F32)
    fold_gate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(w, gate, (float*)out, NK, K);
  else
    fold_gate_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(w, gate, (bf16*)out, NK, K);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// residual epilogue of an inverted-residual block (dwiseneuro.py:136-144, 125-134, 46-54):
//   out = dp[b]*BN4(Y_raw) + BN_sc(shortcut[nearest-strided, channel-tiled]) (+ PE of the next block)
// writes the fp32 trunk (+bf16 copy) and the partial stats of the next block's (strided) shortcut.
// =================================================================================================
template <typename T>
__global__ void block_out_kernel(const T* __restrict__ y_raw, const float* __restrict__ coef4, const float* __restrict__ dp,
                                 const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                 const float* __restrict__ pe_t, const float* __restrict__ pe_h,
                                 const float* __restrict__ pe_w, float* __restrict__ out, bf16* __restrict__ out_bf,
                                 float* __restrict__ partial, NearestMap nh, NearestMap nw, int B, int Tn, int Ho, int Wo,
                                 int Ci, int Co, NearestMap mh, NearestMap mw, int cqc, FastDiv dw, FastDiv dh, FastDiv dt) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cq = tid % cqc, lane = tid / cqc, ln = blockDim.x / cqc;
  const int c = (blockIdx.y * cqc + cq) * 4;
  const int ci = c % Ci;  // cyclic channel tile (Ci % 4 == 0 so quads never straddle)
  float s4[4], h4[4], ss[4], hs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s4[j] = coef4[c + j];
    h4[j] = coef4[Co + c + j];
    ss[j] = coef_sc[c + j];
    hs[j] = coef_sc[Co + c + j];
  }
  float st[3][4] = {};
  const int Hi = mh.in, Wi = mw.in;
  const int Mo = B * Tn * Ho * Wo;
  // the two streamed operands (Y_raw quad, shortcut quad) go through a per-thread cp.async pipeline (see ThreadPipe):
  // with ~90 registers only 2 CTAs fit per SM and plain loads kept < 25 KB in flight per SM
  constexpr int DEPTH = 8;
  ThreadPipe<DEPTH, 2> pipe(smem, blockDim.x, tid);
  const int m0 = blockIdx.x * ln + lane, mstep = gridDim.x * ln;
  auto issue = [&](int k) {
    const int m = m0 + k * mstep;
    if (m < Mo) {
      const int wq = dw.mod(m), r1 = dw.div(m);
      const int hq = dh.mod(r1), bt = dh.div(r1);
      pipe_issue_quad<T>(pipe.slot(k, 0), y_raw + (long)m * Co + c);
      cp_async16_ok(pipe.slot(k, 1), xin + (((long)bt * Hi + mh.src(hq)) * Wi + mw.src(wq)) * Ci + ci);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) issue(k);
  int kk = 0;
  for (int m = m0; m < Mo; m += mstep, ++kk) {
    const int wq = dw.mod(m), r1 = dw.div(m);
    const int hq = dh.mod(r1), bt = dh.div(r1);
    const int tq = dt.mod(bt), b = dt.div(bt);
    float y[4], x[4], o[4];
    // the small cached loads (drop-path scale, position-encoding rows) are issued before the pipeline wait: the asm
    // barrier of cp.async.wait_group keeps the compiler from hoisting them, and after it their latency is exposed
    const float d = dp ? dp[b] : 1.0f;
    float pt[4] = {0.f, 0.f, 0.f, 0.f}, ph[4] = {0.f, 0.f, 0.f, 0.f}, pw[4] = {0.f, 0.f, 0.f, 0.f};
    if (pe_t) {
      ldq(pe_t + tq * Co + c, pt);
      ldq(pe_h + hq * Co + c, ph);
      ldq(pe_w + wq * Co + c, pw);
    }
    cp_async_wait<DEPTH - 1>();
    pipe_read_quad<T>(pipe.slot(kk, 0), y);
    quad_from(*pipe.slot(kk, 1), x);
    issue(kk + DEPTH);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = d * fmaf(y[j], s4[j], h4[j]) + fmaf(x[j], ss[j], hs[j]) + ((pt[j] + ph[j]) + pw[j]);
    stq(out + (long)m * Co + c, o);
    if (out_bf) stq(out_bf + (long)m * Co + c, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) st[2][j] += out_bf ? __bfloat162float(__float2bfloat16_rn(o[j])) : o[j];  // see stem_fwd_kernel
    if (partial && nh.dst(hq) >= 0 && nw.dst(wq) >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { st[0][j] += o[j]; st[1][j] = fmaf(o[j], o[j], st[1][j]); }
    }
  }
  cp_async_wait<0>();
  if (partial)  // [P][3][Co], see stem_fwd_kernel
    block_reduce_channels<3, 4>(st, smem, cqc, ln, partial + (long)blockIdx.x * 3 * Co, Co, blockIdx.y * cqc * 4);
}

extern "C" int dwn_block_out(const void* y_raw, const float* coef4, const float* dp, const float* xin,
                             const float* coef_sc, const float* pe_t, const float* pe_h, const float* pe_w, float* out,
                             void* out_bf, float* partial, int P, int next_stride, int B, int Tn, int Ho, int Wo, int Ci,
                             int Co, int stride, int Hi, int Wi, int dtype, void* stream) {
  DWN_REQUIRE(Ci % 4 == 0 && Co % 4 == 0, "dwn_block_out: channels must be multiples of 4");
  DWN_REQUIRE(Ho == dwn_ceil_div(Hi, stride) && Wo == dwn_ceil_div(Wi, stride),
              "dwn_block_out: output %dx%d is not ceil(%dx%d / %d)", Ho, Wo, Hi, Wi, stride);
  const NearestMap mh(Hi, Ho), mw(Wi, Wo), nh(Ho, dwn_ceil_div(Ho, next_stride)), nw(Wo, dwn_ceil_div(Wo, next_stride));
  int cqc = dwn_largest_divisor_le(Co / 4, 64), ln = 256 / cqc;
  dim3 grid(P, (Co / 4) / cqc), block(cqc * ln);
  size_t sm = (size_t)block.x * 12 * sizeof(float);
  if (ThreadPipe<8, 2>::bytes(block.x) > sm) sm = ThreadPipe<8, 2>::bytes(block.x);
  cudaFuncSetAttribute(block_out_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaFuncSetAttribute(block_out_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (dtype == DWN_DT_F32)
    block_out_kernel<float><<<grid, block, sm, (cudaStream_t)stream>>>((const float*)y_raw, coef4, dp, xin, coef_sc, pe_t,
                                                                       pe_h, pe_w, out, (bf16*)out_bf, partial,
                                                                       nh, nw, B, Tn, Ho, Wo, Ci, Co, mh, mw, cqc,
                                                                       FastDiv(Wo), FastDiv(Ho), FastDiv(Tn));
  else
    block_out_kernel<bf16><<<grid, block, sm, (cudaStream_t)stream>>>((const bf16*)y_raw, coef4, dp, xin, coef_sc, pe_t,
                                                                      pe_h, pe_w, out, (bf16*)out_bf, partial,
                                                                      nh, nw, B, Tn, Ho, Wo, Ci, Co, mh, mw, cqc,
                                                                      FastDiv(Wo), FastDiv(Ho), FastDiv(Tn));
  DWN_LAUNCH_CHECK();
  return 0;
}

// spatial average pool over (H,W) (dwiseneuro.py:374,400): [BT][HW][C] fp32 -> [BT][C] fp32 (+bf16)
__global__ void pool_hw_kernel(const float* __restrict__ in, float* __restrict__ out, bf16* __restrict__ out_bf, int HW,
                               int C) {
  const long bt = blockIdx.x;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < HW; ++i) {
      float v[4];
      ldq(in + (bt * HW + i) * C + c, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] *= inv;
    stq(out + bt * C + c, a);
    if (out_bf) stq(out_bf + bt * C + c, a);
  }
}

extern "C" int dwn_pool_hw(const float* in, float* out, void* out_bf, int BT, int HW, int C, void* stream) {
  pool_hw_kernel<<<BT, 64, 0, (cudaStream_t)stream>>>(in, out, (bf16*)out_bf, HW, C);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// column statistics of a row-major [M][ld] matrix (first C columns): partial[J][2][C]
// =================================================================================================
template <typename T>
__global__ void colstats_kernel(const T* __restrict__ x, long M, int ld, int C, float* __restrict__ partial, int cvc) {
  constexpr int V = VecT<T>::V;
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int cv = tid % cvc, lane = tid / cvc, ln = blockDim.x / cvc;
  const int c = (blockIdx.y * cvc + cv) * V;
  float st[2][V] = {};
#pragma unroll 4
  for (long r = (long)blockIdx.x * ln + lane; r < M; r += (long)gridDim.x * ln) {
    float v[V];
    ldv(x + r * ld + c, v);
#pragma unroll
    for (int j = 0; j < V; ++j) { st[0][j] += v[j]; st[1][j] += v[j] * v[j]; }
  }
  block_reduce_channels<2, V>(st, smem, cvc, ln, partial + (long)blockIdx.x * 2 * C, C, blockIdx.y * cvc * V);
}

extern "C" int dwn_colstats(const void* x, long M, int ld, int C, float* partial, int J, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DWN_DT_F32) {
    DWN_REQUIRE(C % 4 == 0 && ld % 4 == 0, "dwn_colstats: C/ld %% 4 != 0");
    int cvc = dwn_largest_divisor_le(C / 4, 64), ln = 256 / cvc;
    dim3 grid(J, (C / 4) / cvc), block(cvc * ln);
    colstats_kernel<float><<<grid, block, block.x * 8 * sizeof(float), st>>>((const float*)x, M, ld, C, partial, cvc);
  } else {
    DWN_REQUIRE(C % 8 == 0 && ld % 8 == 0, "dwn_colstats: C/ld %% 8 != 0");
    int cvc = dwn_largest_divisor_le(C / 8, 64), ln = 256 / cvc;
    dim3 grid(J, (C / 8) / cvc), block(cvc * ln);
    colstats_kernel<bf16><<<grid, block, block.x * 16 * sizeof(float), st>>>((const bf16*)x, M, ld, C, partial, cvc);
  }
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// cortex layer epilogue (dwiseneuro.py:195-234): BN+SiLU of the grouped conv output, channel
// shuffle folded into the read index, drop-path, + BN_sc(cyclically tiled shortcut).
//   out[m][j] = dp[b]*SiLU(BN(Y[m][src(j)])) + BN_sc(x[m][j mod I]),  src(j) = (j%g)*(O/g) + j/g
// =================================================================================================
template <typename T>
__global__ void cortex_out_kernel(const T* __restrict__ y, const float* __restrict__ coef, const float* __restrict__ dp,
                                  const float* __restrict__ xin, const float* __restrict__ coef_sc,
                                  float* __restrict__ out, bf16* __restrict__ out_bf, int M, int Tn, int I, int O, int G) {
  const int per = O / G;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)M * O; i += (long)gridDim.x * blockDim.x) {
    int m = (int)(i / O), j = (int)(i % O);
    int src = (j % G) * per + j / G;
    float v = ld1<T>(y + (long)m * O + src);
    v = silu_t<T>(fmaf(v, coef[src], coef[O + src]));
    v = rnd<T>(v);
    float d = dp ? dp[m / Tn] : 1.0f;
    float s = fmaf(xin[(long)m * I + (j % I)], coef_sc[j], coef_sc[O + j]);
    float o = d * v + s;
    out[i] = o;
    if (out_bf) out_bf[i] = __float2bfloat16_rn(o);
  }
}

extern "C" int dwn_cortex_out(const void* y, const float* coef, const float* dp, const float* xin, const float* coef_sc,
                              float* out, void* out_bf, int M, int Tn, int I, int O, int G, int dtype, void* stream) {
  long n = (long)M * O;
  int gx = (int)((n + 255) / 256);
  if (gx > 2048) gx = 2048;
  if (dtype == DWN_DT_F32)
    cortex_out_kernel<float><<<gx, 256, 0, (cudaStream_t)stream>>>((const float*)y, coef, dp, xin, coef_sc, out,
                                                                   (bf16*)out_bf, M, Tn, I, O, G);
  else
    cortex_out_kernel<bf16><<<gx, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, coef, dp, xin, coef_sc, out,
                                                                  (bf16*)out_bf, M, Tn, I, O, G);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// readout input: Dropout1d mask (per sample, per channel) applied to the cortex output and cast
// to the GEMM operand type; optional transposed copy for the weight-gradient GEMM.
//   xm[m][k] = x[m][k] * mask[b][k]      xt[k][m] = xm[m][k]              (dwiseneuro.py:276)
// =================================================================================================
template <typename T>
__global__ void readout_prep_kernel(const float* __restrict__ x, const float* __restrict__ mask, T* __restrict__ xm,
                                    T* __restrict__ xt, int M, int K, int Tn) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int m = m0 + r, k = k0 + tx;
    float v = 0.f;
    if (m < M && k < K) {
      v = x[(long)m * K + k];
      if (mask) v *= mask[(long)(m / Tn) * K + k];
      if (xm) st1<T>(xm + (long)m * K + k, v);
    }
    tile[r][tx] = v;
  }
  if (!xt) return;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int k = k0 + r, m = m0 + tx;
    if (m < M && k < K) st1<T>(xt + (long)k * M + m, tile[tx][r]);
  }
}

extern "C" int dwn_readout_prep(const float* x, const float* mask, void* xm, void* xt, int M, int K, int Tn, int dtype,
                                void* stream) {
  dim3 grid((K + 31) / 32, (M + 31) / 32), block(32, 8);
  if (dtype == DWN_DT_F32)
    readout_prep_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(x, mask, (float*)xm, (float*)xt, M, K, Tn);
  else
    readout_prep_kernel<bf16><<<grid, block, 0, (cudaStream_t)stream>>>(x, mask, (bf16*)xm, (bf16*)xt, M, K, Tn);
  DWN_LAUNCH_CHECK();
  return 0;
}

// fp32 -> storage type cast (weight shadows), n % 4 == 0 not required
template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    st1<T>(out + i, in[i]);
}
extern "C" int dwn_cast_bf16(const float* in, void* out, long n, void* stream) {
  int gx = (int)((n + 255) / 256);
  if (gx > 4096) gx = 4096;
  if (gx < 1) gx = 1;
  cast_kernel<bf16><<<gx, 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out, n);
  DWN_LAUNCH_CHECK();
  return 0;
}

// fp32 -> three bf16 planes (hi, mid, lo): x = hi + mid + lo up to 2^-25 |x|.  Operands of the fp32-accurate tensor-core
// GEMM (dwn_gemm split = 3).  The residuals x - hi and x - hi - mid are exact in fp32.
__global__ void __launch_bounds__(256) split3_kernel(const float4* __restrict__ in, uint2* __restrict__ p0,
                                                     uint2* __restrict__ p1, uint2* __restrict__ p2, long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(in + i);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float h[4], m[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
      const float r1 = x[j] - h[j];
      m[j] = __bfloat162float(__float2bfloat16_rn(r1));
      l[j] = r1 - m[j];
    }
    p0[i] = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
    p1[i] = make_uint2(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]));
    p2[i] = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
  }
}
extern "C" int dwn_split3(const float* src, void* dst, long n, void* stream) {
  DWN_REQUIRE(n % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "dwn_split3: n %% 4 != 0 or misaligned");
  const long n4 = n / 4;
  long gx = (n4 + 255) / 256;
  if (gx > 148 * 16) gx = 148 * 16;
  if (gx < 1) gx = 1;
  uint2* d = (uint2*)dst;
  split3_kernel<<<(int)gx, 256, 0, (cudaStream_t)stream>>>((const float4*)src, d, d + n4, d + 2 * n4, n4);
  DWN_LAUNCH_CHECK();
  return 0;
}
