// Deterministic per-channel block reduction used by every stats-producing kernel.
#pragma once
#include "dwn_common.cuh"

// Threads are laid out tid -> (cv = tid % cvt, lane = tid / cvt), blockDim.x == cvt*ln.
// Each thread holds vals[NQ][VV] for channels (c0 + cv*VV + j).  The block sums over lanes in a
// fixed order and writes out[q*C + c0 + cv*VV + j].  smem needs blockDim.x*NQ*VV floats.
template <int NQ, int VV>
__device__ __forceinline__ void block_reduce_channels(float (&vals)[NQ][VV], float* smem, int cvt, int ln, float* out,
                                                      int C, int c0) {
  const int tid = threadIdx.x;
  const int cv = tid % cvt, lane = tid / cvt;
  const int width = cvt * VV;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < VV; ++j) smem[(lane * NQ + q) * width + cv * VV + j] = vals[q][j];
  __syncthreads();
  for (int e = tid; e < NQ * width; e += blockDim.x) {
    const int q = e / width, r = e % width;
    float s = 0.f;
    for (int l = 0; l < ln; ++l) s += smem[(l * NQ + q) * width + r];
    out[(long)q * C + c0 + r] = s;
  }
}
