// Steps either side of the network:
//   f1  input assembly on the device: raw (Hv, Wv, L) video + behaviour + pupil centre -> the (5, T, H, W) clips of a
//       batch of prediction windows, i.e. StackInputsProcessor (inputs.py:22-36) fused with the window gather of
//       Predictor.predict_trial (predictors.py:42-51).  The padded fp32 (5, L, 64, 64) tensor is never built and the
//       host uploads ~36x fewer bytes per trial.
//   f2  validation metric on the device: streaming per-neuron accumulators for CorrelationMetric (metrics.py:11-31,
//       49-74) instead of copying every masked prediction to the host.
#include "dwn_common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// f1
// ---------------------------------------------------------------------------------------------------------------
template <typename VT>
__global__ void assemble_clips_kernel(const VT* __restrict__ video, const float* __restrict__ behavior,
                                      const float* __restrict__ pupil, float* __restrict__ clips, int L, int Hv, int Wv,
                                      int H, int W, int top, int left, float fill, int size, int step, int last0,
                                      long total) {
  const int behind = (size - 1) * step;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long r = i / W;
    const int h = (int)(r % H); r /= H;
    const int k = (int)(r % size); r /= size;
    const int c = (int)(r % 5);
    const int win = (int)(r / 5);
    const int f = last0 + win - behind + k * step;  // indexes.py:21-33, position "last"
    float v;
    if (c == 0) {
      const int hv = h - top, wv = w - left;
      v = (hv >= 0 && hv < Hv && wv >= 0 && wv < Wv) ? (float)video[((long)hv * Wv + wv) * L + f] : fill;
    } else if (c < 3) {
      v = behavior[(long)(c - 1) * L + f];
    } else {
      v = pupil[(long)(c - 3) * L + f];
    }
    clips[i] = v;
  }
}

// video: (Hv, Wv, L) row-major, fp32 (video_dtype 0) or uint8 (video_dtype 2); behavior, pupil: (2, L) fp32.
// clips: (nw, 5, size, H, W) fp32; window i ends at frame last0 + i and holds frames last0+i-(size-1-k)*step.
extern "C" int dwn_assemble_clips(const void* video, int video_dtype, const float* behavior, const float* pupil,
                                  float* clips, int L, int Hv, int Wv, int H, int W, float fill, int size, int step,
                                  int last0, int nw, void* stream) {
  DWN_REQUIRE(Hv <= H && Wv <= W, "dwn_assemble_clips: video %dx%d larger than the %dx%d canvas", Hv, Wv, H, W);
  DWN_REQUIRE(last0 - (size - 1) * step >= 0 && last0 + nw - 1 < L, "dwn_assemble_clips: window range outside the trial");
  if (nw <= 0) return 0;
  const int top = (H - Hv) / 2, left = (W - Wv) / 2;
  const long total = (long)nw * 5 * size * H * W;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (video_dtype == 0)
    assemble_clips_kernel<float><<<gx, 256, 0, st>>>((const float*)video, behavior, pupil, clips, L, Hv, Wv, H, W, top, left,
                                                     fill, size, step, last0, total);
  else if (video_dtype == 2)
    assemble_clips_kernel<unsigned char><<<gx, 256, 0, st>>>((const unsigned char*)video, behavior, pupil, clips, L, Hv, Wv,
                                                             H, W, top, left, fill, size, step, last0, total);
  else
    return dwn_fail("dwn_assemble_clips: video_dtype %d unsupported (0 = fp32, 2 = uint8)", video_dtype);
  DWN_LAUNCH_CHECK();
  return 0;
}

// Training-batch form of f1: B independent clips.  video (B, T, Hv, Wv) fp32 / uint8 (frames already gathered by the
// loader, unpadded), scalars (B, 4, T) fp32 = behaviour (2) + pupil centre (2) per frame -> clips (B, 5, T, H, W) fp32:
// StackInputsProcessor (inputs.py:22-36) per sample, on the device.  42 MB of H2D per batch of 32 become 1.2 MB.
template <typename VT>
__global__ void assemble_batch_kernel(const VT* __restrict__ video, const float* __restrict__ scalars,
                                      float* __restrict__ clips, int T, int Hv, int Wv, int H, int W, int top, int left,
                                      float fill, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long r = i / W;
    const int h = (int)(r % H); r /= H;
    const int t = (int)(r % T); r /= T;
    const int c = (int)(r % 5);
    const long b = r / 5;
    float v;
    if (c == 0) {
      const int hv = h - top, wv = w - left;
      v = (hv >= 0 && hv < Hv && wv >= 0 && wv < Wv) ? (float)video[((b * T + t) * Hv + hv) * Wv + wv] : fill;
    } else {
      v = scalars[(b * 4 + (c - 1)) * T + t];
    }
    clips[i] = v;
  }
}
extern "C" int dwn_assemble_batch(const void* video, int video_dtype, const float* scalars, float* clips, int B, int T,
                                  int Hv, int Wv, int H, int W, float fill, void* stream) {
  DWN_REQUIRE(Hv <= H && Wv <= W, "dwn_assemble_batch: video %dx%d larger than the %dx%d canvas", Hv, Wv, H, W);
  if (B <= 0) return 0;
  const int top = (H - Hv) / 2, left = (W - Wv) / 2;
  const long total = (long)B * 5 * T * H * W;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (video_dtype == 0)
    assemble_batch_kernel<float><<<gx, 256, 0, st>>>((const float*)video, scalars, clips, T, Hv, Wv, H, W, top, left, fill,
                                                     total);
  else if (video_dtype == 2)
    assemble_batch_kernel<unsigned char><<<gx, 256, 0, st>>>((const unsigned char*)video, scalars, clips, T, Hv, Wv, H, W,
                                                             top, left, fill, total);
  else
    return dwn_fail("dwn_assemble_batch: video_dtype %d unsupported (0 = fp32, 2 = uint8)", video_dtype);
  DWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// f2: acc[n][5] (double) += { sum x, sum y, sum xy, sum x^2, sum y^2 } over the samples with weight != 0 and all T frames
//     (x = prediction, y = target); cnt[0] += T * #samples.  One warp per neuron, fixed reduction order: deterministic.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) corr_update_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                         const float* __restrict__ w, int wstride, int B, int n, int T,
                                                         double* __restrict__ acc, double* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int c = 0;
    for (int b = 0; b < B; ++b) c += (w[(long)b * wstride] != 0.0f) ? 1 : 0;
    cnt[0] += (double)c * (double)T;
  }
  if (j >= n) return;
  double s[5] = {0, 0, 0, 0, 0};
  const int BT = B * T;
  for (int idx = lane; idx < BT; idx += 32) {
    const int b = idx / T, t = idx - b * T;
    if (w[(long)b * wstride] != 0.0f) {
      const long o = ((long)b * n + j) * T + t;
      const double x = (double)pred[o], y = (double)target[o];
      s[0] += x; s[1] += y; s[2] += x * y; s[3] += x * x; s[4] += y * y;
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) s[q] = warp_sum_d(s[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[(long)j * 5 + q] += s[q];
  }
}

extern "C" int dwn_corr_update(const float* pred, const float* target, const float* weights, int wstride, int B, int n,
                               int T, double* acc, double* cnt, void* stream) {
  if (B <= 0 || n <= 0 || T <= 0) return 0;
  corr_update_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(pred, target, weights, wstride, B, n, T, acc, cnt);
  DWN_LAUNCH_CHECK();
  return 0;
}

// corr_j = mean((x-mx)/(sx+eps) * (y-my)/(sy+eps)) with biased std (metrics.py:29-31), out_mean = mean_j corr_j
__global__ void __launch_bounds__(1024) corr_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ cnt,
                                                            int n, double eps, float* __restrict__ out,
                                                            float* __restrict__ out_mean) {
  __shared__ double red[1024];
  const double N = cnt[0];
  double part = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double c = 0;
    if (N > 0) {
      const double mx = acc[(long)j * 5] / N, my = acc[(long)j * 5 + 1] / N;
      double vx = acc[(long)j * 5 + 3] / N - mx * mx, vy = acc[(long)j * 5 + 4] / N - my * my;
      vx = vx > 0 ? vx : 0;
      vy = vy > 0 ? vy : 0;
      c = (acc[(long)j * 5 + 2] / N - mx * my) / ((sqrt(vx) + eps) * (sqrt(vy) + eps));
    }
    if (out) out[j] = (float)c;
    part += c;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out_mean[0] = (float)(red[0] / (double)n);
}

extern "C" int dwn_corr_finalize(const double* acc, const double* cnt, int n, double eps, float* out, float* out_mean,
                                 void* stream) {
  DWN_REQUIRE(n > 0, "dwn_corr_finalize: n == 0");
  corr_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(acc, cnt, n, eps, out, out_mean);
  DWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// f3: CutMix and batch collation on the device (mixers.py:52-67, datasets.py:124-129, 172-187).
//   dwn_cutmix     : out[b] = x1[b] with the box rows [bbx1, bbx2) x columns [bby1, bby2) of the last two dims taken from
//                    x2[b] (the reference indexes `inputs[..., bbx1:bbx2, bby1:bby2]`, i.e. "x" runs over H); a sample
//                    whose box is empty (or that the host decided not to mix) is a plain copy of x1.
//   dwn_lerp_rows  : out[b] = (1 - lam[b]) * t1[b] + lam[b] * t2[b]          (target mixing, one row per sample)
//   dwn_scatter_mouse_targets : out_m[b][j][t] = ids[b] == m ? compact[b][j][t] : 0   (per-mouse target tensors from
//                    the compact per-sample targets, so the host never uploads the ~90 % zeros of the collated batch)
// ---------------------------------------------------------------------------------------------------------------
__global__ void cutmix_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const int* __restrict__ boxes,
                              float* __restrict__ out, long planes, int H, int W, long total4) {
  const int W4 = W >> 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int w0 = (int)(i % W4) * 4;
    long r = i / W4;
    const int h = (int)(r % H);
    const long b = r / H / planes;
    const int hx1 = boxes[b * 4], wy1 = boxes[b * 4 + 1], hx2 = boxes[b * 4 + 2], wy2 = boxes[b * 4 + 3];
    float4 a = *reinterpret_cast<const float4*>(x1 + i * 4);
    if (h >= hx1 && h < hx2 && w0 + 3 >= wy1 && w0 < wy2) {
      const float4 c = *reinterpret_cast<const float4*>(x2 + i * 4);
      if (w0 >= wy1 && w0 < wy2) a.x = c.x;
      if (w0 + 1 >= wy1 && w0 + 1 < wy2) a.y = c.y;
      if (w0 + 2 >= wy1 && w0 + 2 < wy2) a.z = c.z;
      if (w0 + 3 >= wy1 && w0 + 3 < wy2) a.w = c.w;
    }
    *reinterpret_cast<float4*>(out + i * 4) = a;
  }
}

// x1, x2, out: (B, planes, H, W) fp32 (planes = C*T); boxes: (B, 4) int32 {bbx1, bby1, bbx2, bby2} as rand_bbox returns
extern "C" int dwn_cutmix(const float* x1, const float* x2, const int* boxes, float* out, int B, long planes, int H, int W,
                          void* stream) {
  DWN_REQUIRE(W % 4 == 0, "dwn_cutmix: W %% 4 != 0");
  const long total4 = (long)B * planes * H * (W / 4);
  if (total4 == 0) return 0;
  int gx = (int)((total4 + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  cutmix_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(x1, x2, boxes, out, planes, H, W, total4);
  DWN_LAUNCH_CHECK();
  return 0;
}

__global__ void lerp_rows_kernel(const float* __restrict__ t1, const float* __restrict__ t2, const float* __restrict__ lam,
                                 float* __restrict__ out, long row, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float l = lam[i / row];
    out[i] = (1.0f - l) * t1[i] + l * t2[i];
  }
}
extern "C" int dwn_lerp_rows(const float* t1, const float* t2, const float* lam, float* out, int B, long row, void* stream) {
  const long total = (long)B * row;
  if (total == 0) return 0;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  lerp_rows_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(t1, t2, lam, out, row, total);
  DWN_LAUNCH_CHECK();
  return 0;
}

__global__ void scatter_mouse_targets_kernel(const float* __restrict__ compact, const int* __restrict__ ids, int m,
                                             float* __restrict__ out, int n_m, int n_max, int T, long total) {
  const long row = (long)n_m * T;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / row, r = i - b * row;
    out[i] = ids[b] == m ? compact[b * (long)n_max * T + r] : 0.0f;
  }
}
// compact: (B, n_max, T) fp32, rows beyond the sample's own neuron count are ignored; out: (B, n_m, T)
extern "C" int dwn_scatter_mouse_targets(const float* compact, const int* ids, int m, float* out, int B, int n_m, int n_max,
                                         int T, void* stream) {
  DWN_REQUIRE(n_m <= n_max, "dwn_scatter_mouse_targets: n_m %d > n_max %d", n_m, n_max);
  const long total = (long)B * n_m * T;
  if (total == 0) return 0;
  int gx = (int)((total + 255) / 256);
  if (gx > 148 * 16) gx = 148 * 16;
  scatter_mouse_targets_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(compact, ids, m, out, n_m, n_max, T, total);
  DWN_LAUNCH_CHECK();
  return 0;
}
