// Probe (not a test): semantics of CUtensorMap elementStrides on sm_100a.  Tensor [H=2][W=16][C=8] bf16 holding its own
// column index; a box over the W dimension with elementStrides[1] = 2 is loaded at w0 = 0 and w0 = 1 for box[1] in {8, 16}.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tests/gpu_checks/build/tma_stride_probe tests/gpu_checks/tma_stride_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int w0, int nbytes, float* out, int nout) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __nv_bfloat16* buf = reinterpret_cast<__nv_bfloat16*>(sm);
  for (int i = threadIdx.x; i < nout; i += blockDim.x) buf[i] = __float2bfloat16(-1.f);
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(sm)), "l"(&map), "r"(b), "r"(0), "r"(w0), "r"(0) : "memory");
  }
  // bounded wait: a wrong byte count must not hang the box
  bool done = false;
  for (int it = 0; it < 2000000 && !done; ++it) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b));
    done = ok != 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = __bfloat162float(buf[i]);
  if (threadIdx.x == 0) out[nout] = done ? 1.f : 0.f;
}

int main() {
  const int H = 2, W = 16, C = 8;
  std::vector<__nv_bfloat16> h(H * W * C);
  for (int y = 0; y < H; ++y) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c) h[(y * W + w) * C + c] = __float2bfloat16((float)(y * 100 + w));
  __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 4096 * 4);
  for (int bw : {8, 16}) for (int w0 : {0, 1}) {
    CUtensorMap map;
    cuuint64_t dims[3] = {C, W, H}; cuuint64_t strides[2] = {C * 2, W * C * 2};
    cuuint32_t box[3] = {C, (cuuint32_t)bw, H}; cuuint32_t estr[3] = {1, 2, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box_w=%d w0=%d encode rc=%d\n", bw, w0, (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int nb : {C * (bw / 2) * H * 2, C * bw * H * 2}) {
      const int nout = C * bw * H;
      probe<<<1, 128, 4096>>>(map, w0, nb, out, nout);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(nout + 1); cudaMemcpy(o.data(), out, (nout + 1) * 4, cudaMemcpyDeviceToHost);
      printf("  expect_tx=%d bytes: err=%d done=%d pixels:", nb, (int)e, (int)o[nout]);
      for (int i = 0; i < nout; i += C) printf(" %g", o[i]);
      printf("\n");
    }
  }
  return 0;
}
