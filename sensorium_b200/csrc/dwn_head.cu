// Readout backward preparation, Poisson loss (forward / backward), distillation target fill and the
// sliding-window overlap-add of the predictor.
#include "dwn_common.cuh"

// =================================================================================================
// MicePoissonLoss for one mouse (losses.py:10-21): sum_{b: w[b]!=0} w[b] * sum_{n,t} (p - y*log(p+eps))
//   wn = normalised weights column of this mouse (stride wstride); partial[J] doubles
// =================================================================================================
__global__ void __launch_bounds__(256) poisson_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                         const float* __restrict__ wn, int wstride, int per_b4, int B,
                                                         float eps, double* __restrict__ partial) {
  // grid (J, B): one sample per blockIdx.y (masked samples exit at once), float4 streaming, fp32 per-thread
  // accumulation over <= a few hundred elements, double across threads
  const int b = blockIdx.y;
  const float w = wn[(long)b * wstride];
  float acc = 0.f;
  if (w != 0.0f) {
    const float4* p4 = reinterpret_cast<const float4*>(pred) + (long)b * per_b4;
    const float4* t4 = reinterpret_cast<const float4*>(tgt) + (long)b * per_b4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_b4; i += gridDim.x * blockDim.x) {
      const float4 p = p4[i], y = t4[i];
      acc += (p.x - y.x * logf(p.x + eps)) + (p.y - y.y * logf(p.y + eps)) + (p.z - y.z * logf(p.z + eps)) +
             (p.w - y.w * logf(p.w + eps));
    }
    acc *= w;
  }
  __shared__ double red[8];
  double d = warp_sum_d((double)acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[(long)b * gridDim.x + blockIdx.x] = s;
  }
}

// partial must hold B*J doubles
extern "C" int dwn_poisson_fwd(const float* pred, const float* tgt, const float* wn, int wstride, int B, long per_b,
                               float eps, double* partial, int J, void* stream) {
  DWN_REQUIRE(per_b % 4 == 0, "dwn_poisson_fwd: n*T must be a multiple of 4");
  dim3 grid(J, B);
  poisson_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, tgt, wn, wstride, (int)(per_b / 4), B, eps, partial);
  DWN_LAUNCH_CHECK();
  return 0;
}

// d loss / d pred = gout * w[b] * (1 - y/(p+eps)), exactly 0 for masked samples
__global__ void poisson_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                   const float* __restrict__ wn, int wstride, const float* __restrict__ gout, int per_b4,
                                   float eps, float* __restrict__ dpred) {
  const int b = blockIdx.y;
  const float w = wn[(long)b * wstride];
  const float gw = *gout * w;
  const float4* p4 = reinterpret_cast<const float4*>(pred) + (long)b * per_b4;
  const float4* t4 = reinterpret_cast<const float4*>(tgt) + (long)b * per_b4;
  float4* d4 = reinterpret_cast<float4*>(dpred) + (long)b * per_b4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_b4; i += gridDim.x * blockDim.x) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w != 0.0f) {
      const float4 p = p4[i], y = t4[i];
      o.x = gw * (1.0f - y.x / (p.x + eps));
      o.y = gw * (1.0f - y.y / (p.y + eps));
      o.z = gw * (1.0f - y.z / (p.z + eps));
      o.w = gw * (1.0f - y.w / (p.w + eps));
    }
    d4[i] = o;
  }
}

extern "C" int dwn_poisson_bwd(const float* pred, const float* tgt, const float* wn, int wstride, const float* gout, int B,
                               long per_b, float eps, float* dpred, void* stream) {
  DWN_REQUIRE(per_b % 4 == 0, "dwn_poisson_bwd: n*T must be a multiple of 4");
  dim3 grid(32, B);
  poisson_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, tgt, wn, wstride, gout, (int)(per_b / 4), eps, dpred);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// readout backward preparation (forward: GEMM epilogue 1 = bias + softplus):
//   dz = dpred * softplus'(z), softplus' = 1 - exp(-beta*p)  (1 in the linear branch beta*p > 20)
//   dz_nm [G*half][M]            (K-major A of the weight-gradient GEMM, rows >= n_out are zero)
//   dz_mn [M][G][half_pad]       (K-major A of the data-gradient GEMM, zero padded)
//   dbias [G*half]
// grid (ceil(half_pad/32), G), block 256 = 32 neurons x 8
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) readout_bwd_prep_kernel(const float* __restrict__ pred,
                                                               const float* __restrict__ dpred, float beta,
                                                               T* __restrict__ dz_nm, T* __restrict__ dz_mn,
                                                               float* __restrict__ dbias, int B, int Tn, int n_out,
                                                               int half, int half_pad, int G) {
  // BG samples per round: all their loads are issued before the first use, so a round costs one DRAM latency
  // instead of BG (the serial per-sample loop made this kernel pure latency: ~44 us for 48 MB)
  constexpr int BG = 4;
  __shared__ float tile[BG][32][33];
  const int g = blockIdx.y;
  const int n0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const int r = tid >> 3, tq = tid & 7;    // phase 1: 8 consecutive lanes share a neuron
  const int r2 = tid & 31, tt = tid >> 5;  // phase 2: neuron fastest
  const int nl = n0 + r;
  const int n = g * half + nl;
  const bool valid = nl < half && n < n_out;
  const int M = B * Tn;
  float db = 0.f;
  for (int b0 = 0; b0 < B; b0 += BG) {
    for (int t0 = 0; t0 < Tn; t0 += 32) {
      float p[BG][4], d[BG][4];
#pragma unroll
      for (int bb = 0; bb < BG; ++bb)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t0 + tq + 8 * u;
          const bool ok = valid && (b0 + bb < B) && t < Tn;
          const long idx = ok ? ((long)(b0 + bb) * n_out + n) * Tn + t : 0;
          p[bb][u] = ok ? pred[idx] : 0.f;
          d[bb][u] = ok ? dpred[idx] : 0.f;
        }
      __syncthreads();
#pragma unroll
      for (int bb = 0; bb < BG; ++bb)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int t = t0 + tq + 8 * u;
          if (b0 + bb < B && t < Tn) {
            const float bp = beta * p[bb][u];
            const float dz = valid ? d[bb][u] * (bp > 20.0f ? 1.0f : 1.0f - expf(-bp)) : 0.f;
            if (nl < half) st1<T>(dz_nm + (long)(g * half + nl) * M + (long)(b0 + bb) * Tn + t, dz);
            db += rnd<T>(dz);
            tile[bb][t - t0][r] = dz;
          }
        }
      __syncthreads();
      if (n0 + r2 < half_pad)
#pragma unroll
        for (int bb = 0; bb < BG; ++bb)
          if (b0 + bb < B)
            for (int t = t0 + tt; t < Tn && t < t0 + 32; t += 8)
              st1<T>(dz_mn + ((long)(b0 + bb) * Tn + t) * ((long)G * half_pad) + (long)g * half_pad + n0 + r2,
                     tile[bb][t - t0][r2]);
    }
  }
  db += __shfl_xor_sync(0xffffffffu, db, 4);
  db += __shfl_xor_sync(0xffffffffu, db, 2);
  db += __shfl_xor_sync(0xffffffffu, db, 1);
  if (tq == 0 && nl < half) dbias[g * half + nl] = db;
}

extern "C" int dwn_readout_bwd_prep(const float* pred, const float* dpred, float beta, void* dz_nm, void* dz_mn,
                                    float* dbias, int B, int Tn, int n_out, int half, int half_pad, int G, int dtype,
                                    void* stream) {
  dim3 grid((half_pad + 31) / 32, G);
  if (dtype == DWN_DT_F32)
    readout_bwd_prep_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(pred, dpred, beta, (float*)dz_nm,
                                                                           (float*)dz_mn, dbias, B, Tn, n_out, half,
                                                                           half_pad, G);
  else
    readout_bwd_prep_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(pred, dpred, beta, (bf16*)dz_nm, (bf16*)dz_mn,
                                                                          dbias, B, Tn, n_out, half, half_pad, G);
  DWN_LAUNCH_CHECK();
  return 0;
}

// dX[m][k] = sum_i mask[i][b][k] * dXm[i][m][k]    (Dropout1d backward summed over the live readouts)
__global__ void readout_dx_combine_kernel(const float* __restrict__ dxm, const float* __restrict__ masks, int nlive,
                                          float* __restrict__ dX, long MK, long BK, int K, int Tn) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < MK; i += (long)gridDim.x * blockDim.x) {
    const long m = i / K;
    const int k = (int)(i % K);
    float s = 0.f;
    for (int j = 0; j < nlive; ++j) {
      const float mk = masks ? masks[(long)j * BK + (m / Tn) * K + k] : 1.0f;
      s += mk * dxm[(long)j * MK + i];
    }
    dX[i] = s;
  }
}

extern "C" int dwn_readout_dx_combine(const float* dxm, const float* masks, int nlive, float* dX, int M, int K, int Tn,
                                      void* stream) {
  long n = (long)M * K;
  int gx = (int)((n + 255) / 256);
  if (gx > 2048) gx = 2048;
  readout_dx_combine_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(dxm, masks, nlive, dX, n, (long)(M / Tn) * K, K, Tn);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// distillation target fill (argus_models.py:31-41): for every (b, mouse) with weight 0 the target row is
// replaced by the teacher prediction and the weight by r/(1-r) * sum(w) / count(w == 0).
// =================================================================================================
__global__ void distill_prepare_kernel(const float* __restrict__ w, int n, float ratio, unsigned char* __restrict__ mask,
                                       float* __restrict__ dweight) {
  __shared__ float s_sum[32];
  __shared__ float s_cnt[32];
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = w[i];
    s += v;
    const bool z = v == 0.0f;
    mask[i] = z ? 1 : 0;
    c += z ? 1.f : 0.f;
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = s; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ts += s_sum[i]; tc += s_cnt[i]; }
    *dweight = ratio / (1.0f - ratio) * ts / tc;
  }
}

__global__ void distill_fill_kernel(float* __restrict__ tgt, const float* __restrict__ teacher,
                                    const unsigned char* __restrict__ mask, int nmice, int mouse, long per_b, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    if (mask[(i / per_b) * nmice + mouse]) tgt[i] = teacher[i];
}

__global__ void distill_weights_kernel(float* __restrict__ w, const unsigned char* __restrict__ mask,
                                       const float* __restrict__ dweight, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mask[i]) w[i] = *dweight;
}

extern "C" int dwn_distill_prepare(const float* w, int n, float ratio, void* mask, float* dweight, void* stream) {
  distill_prepare_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w, n, ratio, (unsigned char*)mask, dweight);
  DWN_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwn_distill_fill(float* tgt, const float* teacher, const void* mask, int nmice, int mouse, int B,
                                long per_b, void* stream) {
  long total = (long)B * per_b;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  distill_fill_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(tgt, teacher, (const unsigned char*)mask, nmice, mouse, per_b,
                                                            total);
  DWN_LAUNCH_CHECK();
  return 0;
}
extern "C" int dwn_distill_weights(float* w, const void* mask, const float* dweight, int n, void* stream) {
  distill_weights_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (const unsigned char*)mask, dweight, n);
  DWN_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// sliding-window overlap-add of Predictor.predict_trial (predictors.py:46-54), gather form:
//   out[n][f] = sum_{windows i covering f} pred[i-behind][n][pos] / max(sum blend[pos], 1)
//   window i covers frames i-behind, i-behind+step, ..., i   (position "last")
// =================================================================================================
__global__ void window_blend_kernel(const float* __restrict__ pred, const float* __restrict__ blend,
                                    float* __restrict__ out, int n_out, int L, int size, int step, int win0, int nwin,
                                    long pred_wstride) {
  const int behind = (size - 1) * step;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)n_out * L; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / L), f = (int)(i % L);
    float s = 0.f, c = 0.f;
    for (int j = size - 1; j >= 0; --j) {      // ascending window index
      const int pos = j;                       // position of frame f inside the window
      const int idx = f + (size - 1 - j) * step;  // window end index
      if (idx < behind || idx >= L) continue;
      const int w = idx - behind - win0;
      if (w < 0 || w >= nwin) continue;
      s += pred[(long)w * pred_wstride + (long)n * size + pos];
      c += blend[pos];
    }
    out[i] = s / fmaxf(c, 1.0f);
  }
}

extern "C" int dwn_window_blend(const float* pred, const float* blend, float* out, int n_out, int L, int size, int step,
                                int win0, int nwin, long pred_wstride, void* stream) {
  long n = (long)n_out * L;
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  window_blend_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(pred, blend, out, n_out, L, size, step, win0, nwin,
                                                            pred_wstride);
  DWN_LAUNCH_CHECK();
  return 0;
}

// gather sliding windows from a processed trial (5, L, H, W) into clips (NW, 5, size, H, W) (predictors.py:51)
__global__ void window_gather_kernel(const float* __restrict__ inp, float* __restrict__ clips, int Cn, int L, long HW,
                                     int size, int step, int first_index, int nwin) {
  const long per_clip = (long)Cn * size * HW;
  const int behind = (size - 1) * step;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)nwin * per_clip;
       i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i / per_clip);
    long r = i % per_clip;
    const int c = (int)(r / (size * HW));
    r %= (size * HW);
    const int j = (int)(r / HW);
    const long hw = r % HW;
    const int frame = first_index + w - behind + j * step;
    clips[i] = inp[((long)c * L + frame) * HW + hw];
  }
}

extern "C" int dwn_window_gather(const float* inp, float* clips, int Cn, int L, long HW, int size, int step,
                                 int first_index, int nwin, void* stream) {
  long n = (long)nwin * Cn * size * HW;
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  window_gather_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(inp, clips, Cn, L, HW, size, step, first_index, nwin);
  DWN_LAUNCH_CHECK();
  return 0;
}
