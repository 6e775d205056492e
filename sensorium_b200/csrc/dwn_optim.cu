// Multi-tensor AdamW (torch.optim.AdamW semantics, configs/true_batch_001.py:45-48) fused with the
// bf16 weight-shadow refresh and the ModelEma update (ema.py:47-55); multi-tensor EMA for buffers.
#include "dwn_common.cuh"

struct DwnTensorEntry {  // 64 bytes, built by the host as an int64[8] row
  float* p;              // parameter (fp32)            | EMA: model tensor (float or int64)
  const float* g;        // gradient                    | EMA: unused
  float* m;              // exp_avg                     | EMA: unused
  float* v;              // exp_avg_sq                  | EMA: unused
  bf16* shadow;          // optional bf16 copy of p     | EMA: unused
  float* ema;            // optional EMA copy of p      | EMA: ema tensor
  long n;                // elements
  long flags;            // bit0: int64 tensor (EMA kernel)
};

constexpr int OPT_CHUNK = 16384;

// steps[t] += active[t]  (per-tensor step counters: tensors without a gradient are skipped entirely), and the two bias
// corrections of the new step count, bc[t] = {1 - b1^t, sqrt(1 - b2^t)}: computed once per tensor here instead of once
// per 16 k-element chunk in the update kernel (two double-precision pow() in front of every CTA's first load)
__global__ void adamw_step_kernel(int* __restrict__ steps, const int* __restrict__ active, int nt, float b1, float b2,
                                  float* __restrict__ bc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  int st = steps[t];
  if (!active || active[t]) steps[t] = ++st;
  bc[2 * t] = (float)(1.0 - pow((double)b1, (double)st));
  bc[2 * t + 1] = (float)sqrt(1.0 - pow((double)b2, (double)st));
}

// Occupancy, not bytes in flight per thread, is what this 7-stream pass wants: 6 CTAs per SM (<= 40 registers) and one
// 16-byte vector per array and iteration run at 82 % of the HBM copy peak; the former unroll-2 body (80 registers, 3 CTAs)
// ran at 67 %, unroll 4 at 54 %, streaming (evict-first) hints changed nothing (tests/gpu_checks/bench_adamw.py)
__global__ void __launch_bounds__(256, 6) adamw_kernel(const DwnTensorEntry* __restrict__ tab,
                                                    const int* __restrict__ chunk_tensor,
                                                    const long* __restrict__ chunk_off, const float* __restrict__ bc,
                                                    const int* __restrict__ active, float lr, float wd, float b1, float b2,
                                                    float eps, float ema_decay, const float* __restrict__ lr_dev) {
  // lr_dev != nullptr: the learning rate is read from device memory (a captured CUDA graph must see the scheduler's
  // current value, a kernel argument would be frozen at capture time)
  if (lr_dev) lr = *lr_dev;
  const int t = chunk_tensor[blockIdx.x];
  if (active && !active[t]) return;
  const DwnTensorEntry e = tab[t];
  const long off = chunk_off[blockIdx.x];
  const long end = min(off + (long)OPT_CHUNK, e.n);
  const float step_size = lr / bc[2 * t];
  const float bc2s = bc[2 * t + 1];
  const float decay = 1.0f - lr * wd;
  auto update = [&](float g, float& p, float& m, float& v) {
    p = p * decay;
    m = m + (g - m) * (1.0f - b1);
    v = v * b2 + (1.0f - b2) * g * g;
    const float denom = sqrtf(v) / bc2s + eps;
    p = p - step_size * (m / denom);
  };
  // 16-byte vector body (chunks start at multiples of 16384 elements, so only the base pointers decide the alignment;
  // gradients that are views into a coalesced data-parallel bucket may be misaligned and take the scalar path)
  const uintptr_t bits = (uintptr_t)e.p | (uintptr_t)e.g | (uintptr_t)e.m | (uintptr_t)e.v | (uintptr_t)e.ema |
                         ((uintptr_t)e.shadow << 1);
  const long nvec = (bits & 15) == 0 ? (end - off) / 4 : 0;
#pragma unroll 1
  for (long q = threadIdx.x; q < nvec; q += blockDim.x) {
    const long i = off + 4 * q;
    const float4 g4 = *reinterpret_cast<const float4*>(e.g + i);
    float4 p4 = *reinterpret_cast<const float4*>(e.p + i);
    float4 m4 = *reinterpret_cast<const float4*>(e.m + i);
    float4 v4 = *reinterpret_cast<const float4*>(e.v + i);
    update(g4.x, p4.x, m4.x, v4.x);
    update(g4.y, p4.y, m4.y, v4.y);
    update(g4.z, p4.z, m4.z, v4.z);
    update(g4.w, p4.w, m4.w, v4.w);
    *reinterpret_cast<float4*>(e.p + i) = p4;
    *reinterpret_cast<float4*>(e.m + i) = m4;
    *reinterpret_cast<float4*>(e.v + i) = v4;
    if (e.shadow) {
      uint2 sh;
      sh.x = pack_bf16x2(p4.x, p4.y);
      sh.y = pack_bf16x2(p4.z, p4.w);
      *reinterpret_cast<uint2*>(e.shadow + i) = sh;
    }
    if (e.ema) {
      float4 a4 = *reinterpret_cast<const float4*>(e.ema + i);
      a4.x = ema_decay * a4.x + (1.0f - ema_decay) * p4.x;
      a4.y = ema_decay * a4.y + (1.0f - ema_decay) * p4.y;
      a4.z = ema_decay * a4.z + (1.0f - ema_decay) * p4.z;
      a4.w = ema_decay * a4.w + (1.0f - ema_decay) * p4.w;
      *reinterpret_cast<float4*>(e.ema + i) = a4;
    }
  }
  for (long i = off + 4 * nvec + threadIdx.x; i < end; i += blockDim.x) {
    const float g = e.g[i];
    float p = e.p[i], m = e.m[i], v = e.v[i];
    update(g, p, m, v);
    e.p[i] = p;
    e.m[i] = m;
    e.v[i] = v;
    if (e.shadow) e.shadow[i] = __float2bfloat16_rn(p);
    if (e.ema) e.ema[i] = ema_decay * e.ema[i] + (1.0f - ema_decay) * p;
  }
}

extern "C" int dwn_adamw(const void* tab, const int* chunk_tensor, const long* chunk_off, int nchunks, int* steps,
                         const int* active, int nt, float lr, float wd, float b1, float b2, float eps, float ema_decay,
                         const float* lr_dev, float* bc_scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  adamw_step_kernel<<<(nt + 127) / 128, 128, 0, st>>>(steps, active, nt, b1, b2, bc_scratch);
  DWN_LAUNCH_CHECK();
  adamw_kernel<<<nchunks, 256, 0, st>>>((const DwnTensorEntry*)tab, chunk_tensor, chunk_off, bc_scratch, active, lr, wd, b1,
                                        b2, eps, ema_decay, lr_dev);
  DWN_LAUNCH_CHECK();
  return 0;
}

// ema <- decay*ema + (1-decay)*model for every state entry; int64 entries go through fp32 and truncate
__global__ void __launch_bounds__(256) ema_kernel(const DwnTensorEntry* __restrict__ tab,
                                                  const int* __restrict__ chunk_tensor,
                                                  const long* __restrict__ chunk_off, float decay) {
  const DwnTensorEntry e = tab[chunk_tensor[blockIdx.x]];
  const long off = chunk_off[blockIdx.x];
  const long end = min(off + (long)OPT_CHUNK, e.n);
  if (e.flags & 1) {
    long long* em = (long long*)e.ema;
    const long long* mo = (const long long*)e.p;
    for (long i = off + threadIdx.x; i < end; i += blockDim.x)
      em[i] = (long long)(decay * (float)em[i] + (1.0f - decay) * (float)mo[i]);
  } else {
    const long nvec = ((((uintptr_t)e.ema | (uintptr_t)e.p) & 15) == 0) ? (end - off) / 4 : 0;
#pragma unroll 2
    for (long q = threadIdx.x; q < nvec; q += blockDim.x) {
      const long i = off + 4 * q;
      float4 a4 = *reinterpret_cast<const float4*>(e.ema + i);
      const float4 p4 = *reinterpret_cast<const float4*>(e.p + i);
      a4.x = decay * a4.x + (1.0f - decay) * p4.x;
      a4.y = decay * a4.y + (1.0f - decay) * p4.y;
      a4.z = decay * a4.z + (1.0f - decay) * p4.z;
      a4.w = decay * a4.w + (1.0f - decay) * p4.w;
      *reinterpret_cast<float4*>(e.ema + i) = a4;
    }
    for (long i = off + 4 * nvec + threadIdx.x; i < end; i += blockDim.x)
      e.ema[i] = decay * e.ema[i] + (1.0f - decay) * e.p[i];
  }
}

extern "C" int dwn_ema(const void* tab, const int* chunk_tensor, const long* chunk_off, int nchunks, float decay,
                       void* stream) {
  ema_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>((const DwnTensorEntry*)tab, chunk_tensor, chunk_off, decay);
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_opt_chunk(void) { return OPT_CHUNK; }

// scale a flat fp32 buffer (gradient averaging after the data-parallel all-reduce)
__global__ void scale_kernel(float* __restrict__ x, long n, float s) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= s;
}
extern "C" int dwn_scale(float* x, long n, float s, void* stream) {
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  scale_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(x, n, s);
  DWN_LAUNCH_CHECK();
  return 0;
}
