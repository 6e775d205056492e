// Algebraic shortcuts around the point-wise expansion conv_pw (E = X W^T, dwiseneuro.py:90-93), bf16 pipeline.
// Both use that E is linear in the block input X, whose Gram matrix Gx = X^T X (ci x ci) and column sums sx are
// tiny compared with E (M x 7ci):
//   forward : BatchNorm statistics of E without reading E:  mean_c = w_c.sx/M,  var_c = w_c^T (Gx/M - mu mu^T) w_c
//   backward: BN1 backward is affine, dE_raw = a*G - d*E - b (G = dE_pre), so
//        dX = G (diag(a) W) - X (W^T diag(d) W) - (b^T W)          -> one dual-K GEMM, no pass over E
//        dW = diag(a) (G^T X) - b (x) sx - diag(d) W Gx            -> the split-K wgrad GEMM on G + a tiny finalize
// This removes colstats(E) and bn_bwd_apply (4 passes over the largest tensor of the network).
#include "dwn_common.cuh"
#include <cstdlib>

// out[c] = sum_p partial[p][q][c]      (block = 32 channels x 32 slices of P)
__global__ void __launch_bounds__(1024) partial_colsum_kernel(const float* __restrict__ partial, int P, int NQ, int q,
                                                             int C, float* __restrict__ out) {
  __shared__ float s0[32][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a = 0.f;
  if (c < C)
    for (int p = sl; p < P; p += 32) a += partial[((long)p * NQ + q) * C + c];
  s0[sl][cl] = a;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  double s = 0;
  for (int i = 0; i < 32; ++i) s += (double)s0[i][cl];
  out[c] = (float)s;
}
extern "C" int dwn_partial_colsum(const float* partial, int P, int NQ, int q, int C, float* out, void* stream) {
  partial_colsum_kernel<<<(C + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partial, P, NQ, q, C, out);
  DWN_LAUNCH_CHECK();
  return 0;
}

// Gram matrix from its split-K partials: gram[j][k] = sum_z part[z][j][k] (fp64 accumulation) and the CENTRED second
// moment cgram[j][k] = gram[j][k]/M - mu_j mu_k formed in fp64 entry by entry (mu = sx/M), stored in fp32: the
// cancellation of E[x x^T] - mu mu^T happens here, in double, so the quadratic forms of pw_stats_kernel are benign.
// block = 32 entries x 32 z-slices
__global__ void __launch_bounds__(1024) gram_finalize_kernel(const float* __restrict__ part, int Z, int ci,
                                                            const float* __restrict__ sx, double count,
                                                            float* __restrict__ gram, float* __restrict__ cgram) {
  __shared__ double sred[32][33];
  const int el = threadIdx.x & 31, zl = threadIdx.x >> 5;
  const long n = (long)ci * ci;
  const long i = (long)blockIdx.x * 32 + el;
  double a = 0.0;
  if (i < n) {
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    int z = zl;
    for (; z + 96 < Z; z += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[u] += (double)part[(long)(z + 32 * u) * n + i];
    }
    for (; z < Z; z += 32) a4[0] += (double)part[(long)z * n + i];
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  sred[zl][el] = a;
  __syncthreads();
  if (zl != 0 || i >= n) return;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 32; ++q) s += sred[q][el];
  const int j = (int)(i / ci), k = (int)(i % ci);
  const double inv_m = 1.0 / count;
  gram[i] = (float)s;
  cgram[i] = (float)(s * inv_m - ((double)sx[j] * inv_m) * ((double)sx[k] * inv_m));
}
extern "C" int dwn_gram_finalize(const float* part, int Z, int ci, const float* sx, double count, float* gram,
                                 float* cgram, void* stream) {
  const long n = (long)ci * ci;
  gram_finalize_kernel<<<(int)((n + 31) / 32), 1024, 0, (cudaStream_t)stream>>>(part, Z, ci, sx, count, gram, cgram);
  DWN_LAUNCH_CHECK();
  return 0;
}

// one warp per output channel c of conv_pw.  var = w^T C w on the centred second moment C (gram_finalize_kernel): lane l
// owns the columns k = l + 32 i of t = w^T C, rows of C are read coalesced in batches of 4 rows x 8 columns of independent
// loads (the former row-slice loop issued one dependent L2 round trip per row group: 100 us at ci = 256); fp32 partial sums
// over 16 rows are flushed into fp64.  A variant that staged C through shared memory measured slower inside the replayed
// step (C is L2-resident there): tests/gpu_checks/run_ab.sh
__global__ void __launch_bounds__(256) pw_stats_kernel(const float* __restrict__ cgram, const float* __restrict__ sx,
                                                             const bf16* __restrict__ w, double count,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ rmean, float* __restrict__ rvar,
                                                             long long* __restrict__ nbt, float momentum, float eps,
                                                             float* __restrict__ coef, int mid, int ci) {
  extern __shared__ __align__(16) float sw[];  // [8][ci]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + wid;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
  if (c >= mid) return;
  float* wc = sw + wid * ci;
  for (int k = lane; k < ci; k += 32) wc[k] = __bfloat162float(w[(long)c * ci + k]);
  __syncwarp();
  const double inv_m = 1.0 / count;
  double m1 = 0;
  for (int k = lane; k < ci; k += 32) m1 += (double)wc[k] * (double)sx[k];
  m1 = warp_sum_d(m1);
  double q = 0;
  for (int kb = 0; kb < ci; kb += 256) {
    double td[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) td[i] = 0.0;
    for (int j0 = 0; j0 < ci; j0 += 16) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = 0.f;
#pragma unroll
      for (int jb = 0; jb < 16; jb += 4) {
        float cv[4][8];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = j0 + jb + jj, k = kb + lane + 32 * i;
            cv[jj][i] = (j < ci && k < ci) ? cgram[(long)j * ci + k] : 0.f;
          }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float wj = wc[min(j0 + jb + jj, ci - 1)];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = fmaf(cv[jj][i], wj, t[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) td[i] += (double)t[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kb + lane + 32 * i;
      if (k < ci) q += td[i] * (double)wc[k];
    }
  }
  q = warp_sum_d(q);
  if (lane != 0) return;
  const double mean = m1 * inv_m;
  double var = q;
  if (var < 0) var = 0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma[c], b = beta[c];
  coef[c] = (float)(g * rstd);
  coef[mid + c] = (float)(b - mean * g * rstd);
  coef[2 * mid + c] = (float)mean;
  coef[3 * mid + c] = (float)rstd;
  if (rmean) {
    const double unb = count > 1 ? var * count / (count - 1.0) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}
extern "C" int dwn_pw_stats(const float* cgram, const float* sx, const void* w_bf16, double count, const float* gamma,
                            const float* beta, float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                            float* coef, int mid, int ci, void* stream) {
  pw_stats_kernel<<<(mid + 7) / 8, 256, 8 * ci * sizeof(float), (cudaStream_t)stream>>>(
      cgram, sx, (const bf16*)w_bf16, count, gamma, beta, rmean, rvar, nbt, momentum, eps, coef, mid, ci);
  DWN_LAUNCH_CHECK();
  return 0;
}

// BN1-backward coefficients per channel: a = scale, d = scale*rstd*c2, b = scale*(c1 - mean*rstd*c2)
__device__ __forceinline__ void bn_bwd_abd(const float* coef, const float* bcoef, int C, int c, float& a, float& b,
                                           float& d) {
  const float sc = coef[c], mu = coef[2 * C + c], rs = coef[3 * C + c];
  const float c1 = bcoef[c], c2 = bcoef[C + c];
  a = sc;
  d = sc * rs * c2;
  b = sc * (c1 - mu * rs * c2);
}

// step A: per-channel BN1-backward coefficients abd[3][mid] and W' = diag(a) W (bf16)
__global__ void pw_abd_kernel(const float* __restrict__ coef, const float* __restrict__ bcoef, const bf16* __restrict__ w,
                              bf16* __restrict__ wprime, float* __restrict__ abd, int mid, int ci) {
  const int c = blockIdx.x;
  float a, b, d;
  bn_bwd_abd(coef, bcoef, mid, c, a, b, d);
  if (threadIdx.x == 0) { abd[c] = a; abd[mid + c] = b; abd[2 * mid + c] = d; }
  for (int k = threadIdx.x; k < ci; k += blockDim.x)
    wprime[(long)c * ci + k] = __float2bfloat16_rn(a * __bfloat162float(w[(long)c * ci + k]));
}

// step B: partial[s][j][k] = sum_{c in split s} coeff(c,j) * W[c][k],  coeff = d_c*W[c][j] (j < ci) or b_c (j == ci).
// The split's W rows and the 16 coefficients per row are staged in shared memory once (one coalesced load phase),
// so the accumulation loop has no dependent global loads.
constexpr int PWQ_JT = 16, PWQ_SPLIT = 32;
__global__ void pw_q_partial_kernel(const float* __restrict__ abd, const bf16* __restrict__ w,
                                    float* __restrict__ partial, int mid, int ci) {
  extern __shared__ __align__(16) unsigned char pwq_sm[];
  const int k = threadIdx.x;
  const int j0 = blockIdx.x * PWQ_JT, sidx = blockIdx.y;
  const int cper = (mid + PWQ_SPLIT - 1) / PWQ_SPLIT;
  const int cbeg = sidx * cper, cend = min(mid, cbeg + cper);
  const int cnt = max(cend - cbeg, 0);
  float* scf = reinterpret_cast<float*>(pwq_sm);                       // [cper][PWQ_JT]
  bf16* sw = reinterpret_cast<bf16*>(pwq_sm + (size_t)cper * PWQ_JT * sizeof(float));  // [cper][ci]
  if ((ci & 7) == 0) {  // 16-byte copies, all independent (the scalar loop was 56 dependent-issue rounds at ci = 256)
    const uint4* src = reinterpret_cast<const uint4*>(w + (long)cbeg * ci);
    uint4* dst = reinterpret_cast<uint4*>(sw);
    for (int i = k; i < cnt * ci / 8; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int i = k; i < cnt * ci; i += blockDim.x) sw[i] = w[(long)cbeg * ci + i];
  }
  for (int i = k; i < cnt * PWQ_JT; i += blockDim.x) {
    const int cl = i / PWQ_JT, j = j0 + i % PWQ_JT, c = cbeg + cl;
    scf[i] = j < ci ? abd[2 * mid + c] * __bfloat162float(w[(long)c * ci + j]) : (j == ci ? abd[mid + c] : 0.f);
  }
  __syncthreads();
  if (k >= ci) return;
  float acc[PWQ_JT];
#pragma unroll
  for (int jj = 0; jj < PWQ_JT; ++jj) acc[jj] = 0.f;
#pragma unroll 2
  for (int cl = 0; cl < cnt; ++cl) {
    const float wk = __bfloat162float(sw[cl * ci + k]);
    const float4* cf = reinterpret_cast<const float4*>(scf + cl * PWQ_JT);
#pragma unroll
    for (int q = 0; q < PWQ_JT / 4; ++q) {
      const float4 f = cf[q];
      acc[4 * q + 0] = fmaf(f.x, wk, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(f.y, wk, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(f.z, wk, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(f.w, wk, acc[4 * q + 3]);
    }
  }
#pragma unroll
  for (int jj = 0; jj < PWQ_JT; ++jj)
    if (j0 + jj <= ci) partial[((long)sidx * (ci + 1) + j0 + jj) * ci + k] = acc[jj];
}

// step C: -Q (bf16, ci x ci) and r (fp32, ci) from the split partials
__global__ void pw_q_finalize_kernel(const float* __restrict__ partial, bf16* __restrict__ negq, float* __restrict__ r,
                                     int ci) {
  const int j = blockIdx.x, k = threadIdx.x;
  if (k >= ci) return;
  float s = 0.f;
  for (int sp = 0; sp < PWQ_SPLIT; ++sp) s += partial[((long)sp * (ci + 1) + j) * ci + k];
  if (j < ci) negq[(long)j * ci + k] = __float2bfloat16_rn(-s);
  else r[k] = s;
}

// scratch: (3*mid + PWQ_SPLIT*(ci+1)*ci) floats
extern "C" int dwn_pw_bwd_prep(const float* coef, const float* bcoef, const void* w_bf16, void* wprime, void* negq,
                               float* r, float* scratch, int mid, int ci, void* stream) {
  DWN_REQUIRE(ci <= 1024, "dwn_pw_bwd_prep: ci > 1024");
  cudaStream_t st = (cudaStream_t)stream;
  const int nt = ci < 32 ? 32 : ci;
  float* abd = scratch;
  float* partial = scratch + 3 * (size_t)mid;
  pw_abd_kernel<<<mid, nt > 256 ? 256 : nt, 0, st>>>(coef, bcoef, (const bf16*)w_bf16, (bf16*)wprime, abd, mid, ci);
  DWN_LAUNCH_CHECK();
  dim3 gq((ci + 1 + PWQ_JT - 1) / PWQ_JT, PWQ_SPLIT);
  const int cper = (mid + PWQ_SPLIT - 1) / PWQ_SPLIT;
  const size_t qsm = (size_t)cper * PWQ_JT * sizeof(float) + (size_t)cper * ci * sizeof(bf16);
  DWN_REQUIRE(qsm <= 48 * 1024, "dwn_pw_bwd_prep: shared memory");
  pw_q_partial_kernel<<<gq, nt, qsm, st>>>(abd, (const bf16*)w_bf16, partial, mid, ci);
  DWN_LAUNCH_CHECK();
  pw_q_finalize_kernel<<<ci + 1, nt, 0, st>>>(partial, (bf16*)negq, r, ci);
  DWN_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwn_pw_bwd_prep_scratch(int mid, int ci) { return 3 * mid + PWQ_SPLIT * (ci + 1) * ci; }

// dW[c][k] = a_c*P[c][k] - b_c*sx[k] - d_c * sum_j W[c][j]*Gx[j][k]     (block = channel c, thread = k)
__global__ void pw_wgrad_finalize_kernel(const float* __restrict__ Psum, const float* __restrict__ coef,
                                         const float* __restrict__ bcoef, const bf16* __restrict__ w,
                                         const float* __restrict__ gram, const float* __restrict__ sx,
                                         float* __restrict__ dw, int mid, int ci) {
  extern __shared__ float swc[];
  const int c = blockIdx.x, k = threadIdx.x;
  for (int j = k; j < ci; j += blockDim.x) swc[j] = __bfloat162float(w[(long)c * ci + j]);
  __syncthreads();
  if (k >= ci) return;
  float a, b, d;
  bn_bwd_abd(coef, bcoef, mid, c, a, b, d);
  float s = 0.f;
  for (int j = 0; j < ci; ++j) s = fmaf(swc[j], gram[(long)j * ci + k], s);
  dw[(long)c * ci + k] = a * Psum[(long)c * ci + k] - b * sx[k] - d * s;
}
extern "C" int dwn_pw_wgrad_finalize(const float* Psum, const float* coef, const float* bcoef, const void* w_bf16,
                                     const float* gram, const float* sx, float* dw, int mid, int ci, void* stream) {
  DWN_REQUIRE(ci <= 1024, "dwn_pw_wgrad_finalize: ci > 1024");
  pw_wgrad_finalize_kernel<<<mid, ci < 32 ? 32 : ci, ci * sizeof(float), (cudaStream_t)stream>>>(
      Psum, coef, bcoef, (const bf16*)w_bf16, gram, sx, dw, mid, ci);
  DWN_LAUNCH_CHECK();
  return 0;
}
