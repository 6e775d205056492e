// Micro-benchmark (not product code): how fast can cp.async.bulk.tensor.4d stage the halo tiles of the spatial
// depth-wise kernels?  Tensor [NP][H][W][C] bf16 (channels-last), box (CC, W, THI+2, 1) with out-of-bounds zero fill,
// 256-thread CTAs, 2 CTAs per SM, ring of STAGES tiles, consumers read every staged byte once from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/gpu_checks/build/tma_box_bench
//        tests/gpu_checks/tma_box_bench.cu        Run: tests/gpu_checks/build/tma_box_bench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

static inline __device__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static inline __device__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
static inline __device__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
static inline __device__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok != 0;
}
static inline __device__ void mbar_wait(uint64_t* b, uint32_t par) {
  for (uint32_t it = 0; !mbar_try(b, par); ++it)
    if (it > (1u << 26)) __trap();
}
static inline __device__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(256, 2)
box_kernel(const __grid_constant__ CUtensorMap map, unsigned* sink, int NP, int THI, int NR, int CC, int nchunks, int nbsh,
           int tile_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[STAGES];
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % nchunks, worker = blockIdx.x / nchunks, nworkers = gridDim.x / nchunks;
  const int ntiles = NP << nbsh, nbm = (1 << nbsh) - 1;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    mbar_expect(&full[s], tile_bytes);
    tma4(smem + (size_t)s * tile_bytes, &map, &full[s], chunk * CC, 0, (t & nbm) * THI - 1, t >> nbsh);
  };
  int t = worker;
  if (tid == 0)
    for (int s = 0; s < STAGES - 1; ++s)
      if (t + s * nworkers < ntiles) issue(t + s * nworkers, s);
  unsigned acc = 0;
  const int nvec = tile_bytes / 16;
  for (int k = 0; t < ntiles; t += nworkers, ++k) {
    const int s = k % STAGES;
    if (tid == 0 && t + (STAGES - 1) * nworkers < ntiles) issue(t + (STAGES - 1) * nworkers, (k + STAGES - 1) % STAGES);
    mbar_wait(&full[s], (k / STAGES) & 1);
    const uint4* src = reinterpret_cast<const uint4*>(smem + (size_t)s * tile_bytes);
    for (int i = tid; i < nvec; i += 256) {
      const uint4 q = src[i];
      acc ^= q.x ^ q.y ^ q.z ^ q.w;
    }
    __syncthreads();  // everybody is done with stage s before it is refilled (next iteration's issue targets (k+STAGES)%STAGES)
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
  if (!enc) { printf("no encode fn\n"); return 1; }
  struct Shape { const char* tag; int NP, H, W, C, CC; };
  const Shape shapes[] = {{"blk0 E (64x64, C=448, CC=16)", 512, 64, 64, 448, 16},
                          {"blk0 E (64x64, C=448, CC=32 half rows)", 512, 64, 64, 448, 32},
                          {"blk1 E (32x32, C=448, CC=32)", 512, 32, 32, 448, 32},
                          {"blk5 E (16x16, C=896, CC=64)", 512, 16, 16, 896, 64},
                          {"blk8 E (8x8, C=1792, CC=128)", 512, 8, 8, 1792, 128}};
  unsigned* sink;
  cudaMalloc(&sink, 4);
  unsigned char* flush;
  cudaMalloc(&flush, 256u << 20);
  for (const Shape& s : shapes) {
    const size_t n = (size_t)s.NP * s.H * s.W * s.C;
    __nv_bfloat16* x;
    cudaMalloc(&x, n * 2);
    cudaMemset(x, 1, n * 2);
    const int THI = 8, NR = THI + 2;
    const int BW = 1024 / s.CC;  // box width in pixels (W, or half a row for the CC=32 variant of blk0)
    for (int stages = 2; stages <= 3; ++stages) {
      CUtensorMap map;
      cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.NP};
      cuuint64_t strides[3] = {(cuuint64_t)s.C * 2, (cuuint64_t)s.W * s.C * 2, (cuuint64_t)s.H * s.W * s.C * 2};
      cuuint32_t box[4] = {(cuuint32_t)s.CC, (cuuint32_t)BW, (cuuint32_t)NR, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int tile_bytes = NR * BW * s.CC * 2;
      const int nchunks = s.C / s.CC * (s.W / BW);  // the half-row variant just covers the left half twice (same bytes)
      int nbsh = 0;
      while ((THI << nbsh) < s.H) ++nbsh;
      (void)nchunks;
      const int P = 148;  // workers per channel chunk, as the product kernels launch
      const size_t smem = (size_t)stages * tile_bytes;
      auto kern = stages == 2 ? box_kernel<2> : box_kernel<3>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float best = 1e9f;
      for (int it = 0; it < 5; ++it) {
        cudaMemsetAsync(flush, it, 256u << 20);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern<<<P * (s.C / s.CC), 256, smem>>>(map, sink, s.NP, THI, NR, s.CC, s.C / s.CC, nbsh, tile_bytes);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
      }
      const double bytes = (double)n * 2 * (BW == s.W ? 1.0 : 0.5);
      printf("%-42s stages=%d tile=%6d B grid=%5d  %.3f ms  %.0f GB/s (algorithmic, halo rows re-read from L2)\n", s.tag, stages,
             tile_bytes, P * (s.C / s.CC), best, bytes / best * 1e-6);
      fflush(stdout);
    }
    cudaFree(x);
  }
  return 0;
}
