// GEMMs of the DwiseNeuro hot path.
//   * bf16 pipeline: persistent, warp-specialised tcgen05 kernel (TMA -> smem ring -> tcgen05.mma with
//     fp32 accumulators in TMEM (double buffered) -> tcgen05.ld epilogue), sm_100a only.
//   * fp32 pipeline: SIMT FFMA tile kernel (the 1e-4 parity mode; TF32 would break the bound).
// Every point-wise Conv3d / grouped Conv1d of the reference (dwiseneuro.py:90-93,117-120,206,277) and
// all their data / weight gradients are expressed as  D[z] (MxN) = A[z] (MxK) * B[z]^T (NxK).
// Operands are described by a "major": K-major = row-major [rows][K]; MN-major = row-major [K][rows].
#include "dwn_common.cuh"
#include "../../include/dwn_b200.h"
#include <cuda.h>
#include <mutex>
#include <cudaTypedefs.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// shared epilogue math
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_beta(float x, float beta) {
  // nn.Softplus(beta, threshold=20)  (dwiseneuro.py:281)
  float bx = beta * x;
  return bx > 20.0f ? x : log1pf(expf(bx)) / beta;
}

struct EpiParams {
  int epi;            // 0 = row-major store, 1 = readout (bias + softplus, transposed to [b][n][t])
  int d_bf16;         // element type of D for epi 0
  void* D;
  long ldd, d_zstride;
  int m_limit, n_limit;
  const float* bias;
  float beta;
  int Tn, n_out_total, row_offset_per_z;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps (clean launch failure) instead of hanging the GPU
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 "version 1" (cute/arch/mma_sm100_desc.hpp layout)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// one warp stores its staged 32-row x 64-column sub-tile (row pitch 272 bytes) with coalesced 16-byte vectors
template <int ESZ>
__device__ __forceinline__ void store_subtile(const uint8_t* my_stage, uint8_t* dbase, long ldd, int row0, int n0, int c0,
                                              int lane, int m_limit, int block_n, int n_limit) {
  constexpr int NVEC = 64 * ESZ / 16;  // 16-byte vectors per row segment: 8 (bf16) or 16 (fp32)
  constexpr int EPV = 16 / ESZ;        // elements per vector
  constexpr int RSTEP = 32 / NVEC;     // rows covered by one warp-wide store: 4 or 2
  const int j = lane % NVEC, r0 = lane / NVEC;
  const int lcol = c0 + j * EPV, gcol = n0 + lcol;
  if (lcol >= block_n || gcol >= n_limit) return;
  const uint8_t* src = my_stage + r0 * 272 + j * 16;
  uint8_t* dst = dbase + ((size_t)(row0 + r0) * ldd + gcol) * ESZ;
  const size_t dstep = (size_t)RSTEP * ldd * ESZ;
  const int rows = m_limit - row0 - r0;  // store while it * RSTEP < rows
#pragma unroll
  for (int it = 0; it < NVEC; ++it) {
    if (it * RSTEP < rows) *reinterpret_cast<uint4*>(dst + it * dstep) = *reinterpret_cast<const uint4*>(src + it * RSTEP * 272);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 GEMM kernel
// ------------------------------------------------------------------------------------------------
constexpr int GT_BM = 128;        // UMMA M (cta_group::1)
constexpr int GT_BK = 64;         // one 128-byte swizzle atom of bf16 along K
constexpr int GT_A_BYTES = GT_BM * GT_BK * 2;
constexpr int GT_STAGE_PITCH = 272;                    // staging row pitch (bytes): 64 fp32 + 16 pad
constexpr int GT_EPI_WARPS = 8;   // two warps per TMEM lane quarter, interleaved over 64-column sub-tiles
constexpr int GT_STAGING_BYTES = GT_EPI_WARPS * 32 * GT_STAGE_PITCH;
constexpr int GT_THREADS = 64 + 32 * GT_EPI_WARPS;   // warp0 TMA, warp1 MMA, warps2.. epilogue
static_assert(GT_STAGE_PITCH == 272, "store_subtile() assumes the 272-byte staging pitch");

constexpr int GT_MAX_SEG = 6;
// Operand maps of one launch.  The K loop runs over up to GT_MAX_SEG "segments" (a_map, b_map, K): one for a plain GEMM,
// two for the dual-K form (D = A B^T + A2 B2^T), six for the fp32-accurate 3-way bf16 split (planes hi/mid/lo of both
// operands, products lo*hi, mid*mid, hi*lo, mid*hi, hi*mid, hi*hi accumulated into the same fp32 TMEM tile).
struct alignas(64) GemmMaps {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

struct GemmTcParams {
  int block_n, b_bytes, stages;
  int nseg, seg_k[GT_MAX_SEG];
  signed char seg_a[GT_MAX_SEG], seg_b[GT_MAX_SEG];
  int num_m_blocks, num_n_blocks, Z;
  int a_zmode, b_zmode, b_batch_rows;
  uint32_t lbo_a, lbo_b;  // debug override of MN-major LBO/SBO (0 = default)
  uint32_t sbo_a, sbo_b;
  EpiParams e;
};

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = GT_A_BYTES + p.b_bytes;
  uint8_t* staging = smem + (size_t)p.stages * stage_bytes;
  uint64_t* bars = (uint64_t*)(staging + GT_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tfull_bar = bars + 2 * p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], GT_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int num_k_blocks = 0;
  for (int sg = 0; sg < p.nseg; ++sg) num_k_blocks += (p.seg_k[sg] + GT_BK - 1) / GT_BK;
  const int total_tiles = p.num_m_blocks * p.num_n_blocks * p.Z;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.num_n_blocks;
      const int m_blk = (tile / p.num_n_blocks) % p.num_m_blocks;
      const int z = tile / (p.num_n_blocks * p.num_m_blocks);
      const int za = p.a_zmode == 1 ? z : 0;
      const int zb = p.b_zmode == 1 ? z : (p.b_zmode == 2 ? (m_blk * GT_BM) / p.b_batch_rows : 0);
      for (int sg = 0; sg < p.nseg; ++sg) {
        const CUtensorMap* ma = &maps.a[p.seg_a[sg]];
        const CUtensorMap* mb = &maps.b[p.seg_b[sg]];
        const int nk = (p.seg_k[sg] + GT_BK - 1) / GT_BK;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + GT_A_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)(GT_A_BYTES + p.block_n * GT_BK * 2));
          const int k0 = kb * GT_BK;
          if (A_MN) {
            tma_load_3d(sa, ma, &full_bar[stage], m_blk * GT_BM, k0, za);
            tma_load_3d(sa + 8192, ma, &full_bar[stage], m_blk * GT_BM + 64, k0, za);
          } else {
            tma_load_3d(sa, ma, &full_bar[stage], k0, m_blk * GT_BM, za);
          }
          if (B_MN) {
            for (int j = 0; j < p.block_n / 64; ++j)
              tma_load_3d(sb + j * 8192, mb, &full_bar[stage], n_blk * p.block_n + j * 64, k0, zb);
          } else {
            tma_load_3d(sb, mb, &full_bar[stage], k0, n_blk * p.block_n, zb);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                           ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
    const uint32_t lbo_a = A_MN ? (p.lbo_a ? p.lbo_a : 8192u) : 16u, sbo_a = A_MN ? (p.sbo_a ? p.sbo_a : 1024u) : 1024u;
    const uint32_t lbo_b = B_MN ? (p.lbo_b ? p.lbo_b : 8192u) : 16u, sbo_b = B_MN ? (p.sbo_b ? p.sbo_b : 1024u) : 1024u;
    const uint32_t kstep_a = A_MN ? 2048u : 32u, kstep_b = B_MN ? 2048u : 32u;  // bytes per UMMA_K=16
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[buf], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + GT_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < GT_BK / 16; ++kk) {
          const uint64_t ad = make_sdesc(sa + kk * kstep_a, lbo_a, sbo_a);
          const uint64_t bd = make_sdesc(sb + kk * kstep_b, lbo_b, sbo_b);
          umma_bf16(tmem_d, ad, bd, idesc, (kb | kk) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_k_blocks - 1) umma_commit(&tfull_bar[buf]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ================= epilogue (8 warps; warp w owns TMEM lanes 32*(w%4)..+31 and every other sub-tile) ======
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;  // 0 / 1: which of the interleaved column sub-tiles this warp stores
    uint8_t* my_stage = staging + (size_t)(warp - 2) * 32 * GT_STAGE_PITCH;
    const EpiParams& e = p.e;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % p.num_n_blocks;
      const int m_blk = (tile / p.num_n_blocks) % p.num_m_blocks;
      const int z = tile / (p.num_n_blocks * p.num_m_blocks);
      const int buf = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[buf], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 256u;
      const int row0 = m_blk * GT_BM + q * 32;  // first row of this warp (within z)
      const int n0 = n_blk * p.block_n;
      if (e.epi == 0) {
        const int esz = e.d_bf16 ? 2 : 4;
        uint8_t* dbase = (uint8_t*)e.D + (size_t)z * e.d_zstride * esz;
        for (int c0 = hsel * 64; c0 < p.block_n; c0 += 128) {
          if (n0 + c0 >= e.n_limit) break;
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr + c0, v0);
          tmem_ld32(taddr + c0 + 32, v1);
          tmem_ld_wait();
          uint8_t* srow = my_stage + lane * GT_STAGE_PITCH;
          if (e.d_bf16) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]));
              w.y = pack_bf16x2(__uint_as_float(v0[j + 2]), __uint_as_float(v0[j + 3]));
              w.z = pack_bf16x2(__uint_as_float(v0[j + 4]), __uint_as_float(v0[j + 5]));
              w.w = pack_bf16x2(__uint_as_float(v0[j + 6]), __uint_as_float(v0[j + 7]));
              *reinterpret_cast<uint4*>(srow + j * 2) = w;
              w.x = pack_bf16x2(__uint_as_float(v1[j]), __uint_as_float(v1[j + 1]));
              w.y = pack_bf16x2(__uint_as_float(v1[j + 2]), __uint_as_float(v1[j + 3]));
              w.z = pack_bf16x2(__uint_as_float(v1[j + 4]), __uint_as_float(v1[j + 5]));
              w.w = pack_bf16x2(__uint_as_float(v1[j + 6]), __uint_as_float(v1[j + 7]));
              *reinterpret_cast<uint4*>(srow + 64 + j * 2) = w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              *reinterpret_cast<uint4*>(srow + j * 4) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
              *reinterpret_cast<uint4*>(srow + 128 + j * 4) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
            }
          }
          __syncwarp();
          // coalesced 16-byte stores: a lane keeps its column vector j and walks down the rows, so the column
          // predicate and the address are loop invariants plus a constant row step (no per-store div/mod)
          if (e.d_bf16) store_subtile<2>(my_stage, dbase, e.ldd, row0, n0, c0, lane, e.m_limit, p.block_n, e.n_limit);
          else store_subtile<4>(my_stage, dbase, e.ldd, row0, n0, c0, lane, e.m_limit, p.block_n, e.n_limit);
          __syncwarp();
        }
      } else {
        // readout: rows = neurons of group z, columns = (b,t); out[b][n][t] fp32
        const int rows_valid = min(e.row_offset_per_z, e.n_out_total - z * e.row_offset_per_z);
        const int row = row0 + lane;
        const bool row_ok = row < rows_valid && row < e.m_limit;
        const int gn = z * e.row_offset_per_z + row;
        const float bias = row_ok ? e.bias[gn] : 0.f;
        float* out = (float*)e.D;
        for (int c0 = hsel * 32; c0 < p.block_n; c0 += 64) {
          if (n0 + c0 >= e.n_limit) break;
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            if ((e.Tn & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = n0 + c0 + j;
                if (c0 + j < p.block_n && col < e.n_limit) {
                  const int b = col / e.Tn, t = col % e.Tn;
                  float4 o;
                  o.x = softplus_beta(__uint_as_float(v[j]) + bias, e.beta);
                  o.y = softplus_beta(__uint_as_float(v[j + 1]) + bias, e.beta);
                  o.z = softplus_beta(__uint_as_float(v[j + 2]) + bias, e.beta);
                  o.w = softplus_beta(__uint_as_float(v[j + 3]) + bias, e.beta);
                  *reinterpret_cast<float4*>(out + ((size_t)b * e.n_out_total + gn) * e.Tn + t) = o;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = n0 + c0 + j;
                if (c0 + j < p.block_n && col < e.n_limit) {
                  const int b = col / e.Tn, t = col % e.Tn;
                  out[((size_t)b * e.n_out_total + gn) * e.Tn + t] = softplus_beta(__uint_as_float(v[j]) + bias, e.beta);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
  }
  return fn;
}

// bf16 operand map.  mn_major=0: memory [Z][rows][K] (ld = row pitch) -> dims (K, rows, Z), box (64, box_rows, 1)
//                    mn_major=1: memory [Z][K][rows] (ld = row pitch) -> dims (rows, K, Z), box (64, 64, 1)
// Encoding is a pure function of (pointer, shape, strides, box): maps are memoised in a direct-mapped table, so an eager
// step re-uses the ~120 maps of the previous step instead of calling the driver for each (SURVEY.md 8b "cached maps";
// a replayed CUDA graph carries its maps as kernel parameters and never gets here).
struct GemmMapKey {
  const void* ptr;
  long rows, K, ld, zstride;
  int mn_major, zdim, box_rows;
  bool operator==(const GemmMapKey& o) const {
    return ptr == o.ptr && rows == o.rows && K == o.K && ld == o.ld && zstride == o.zstride && mn_major == o.mn_major &&
           zdim == o.zdim && box_rows == o.box_rows;
  }
};
static int make_operand_map_uncached(CUtensorMap* map, const void* ptr, int mn_major, long rows, long K, long ld,
                                     long zstride, int zdim, int box_rows);
static int make_operand_map(CUtensorMap* map, const void* ptr, int mn_major, long rows, long K, long ld, long zstride,
                            int zdim, int box_rows) {
  constexpr int NSLOT = 1024;
  static GemmMapKey keys[NSLOT];
  static CUtensorMap vals[NSLOT];
  static bool used[NSLOT];
  static std::mutex mu;
  const GemmMapKey key{ptr, rows, K, ld, zstride, mn_major, zdim, box_rows};
  size_t h = (size_t)((uintptr_t)ptr >> 8) * 1000003u;
  h ^= (size_t)rows * 31 + (size_t)K * 8191 + (size_t)ld * 131071 + (size_t)zstride * 7 + (size_t)mn_major * 524287 +
       (size_t)zdim * 65599 + (size_t)box_rows * 2654435761u;
  const int slot = (int)(h % NSLOT);
  std::lock_guard<std::mutex> lock(mu);
  if (used[slot] && keys[slot] == key) {
    *map = vals[slot];
    return 0;
  }
  if (make_operand_map_uncached(map, ptr, mn_major, rows, K, ld, zstride, zdim, box_rows)) return -1;
  keys[slot] = key;
  vals[slot] = *map;
  used[slot] = true;
  return 0;
}
static int make_operand_map_uncached(CUtensorMap* map, const void* ptr, int mn_major, long rows, long K, long ld,
                                     long zstride, int zdim, int box_rows) {
  auto enc = get_encode_fn();
  if (!enc) return dwn_fail("cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3];
  cuuint64_t strides[2];
  cuuint32_t box[3];
  cuuint32_t estr[3] = {1, 1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    box[0] = 64; box[1] = (cuuint32_t)box_rows;
  } else {
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
    box[0] = 64; box[1] = 64;
  }
  dims[2] = (cuuint64_t)(zdim > 0 ? zdim : 1);
  box[2] = 1;
  strides[0] = (cuuint64_t)ld * 2;
  strides[1] = (cuuint64_t)(zdim > 1 ? zstride : (long)dims[1] * ld) * 2;
  if (strides[1] == 0) strides[1] = strides[0];
  if (((uintptr_t)ptr & 15) || (strides[0] & 15) || (strides[1] & 15))
    return dwn_fail("gemm operand not 16-byte aligned (ptr=%p ld=%ld zstride=%ld)", ptr, ld, zstride);
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dwn_fail("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) ld=%ld", (int)r,
                                         (unsigned long long)dims[0], (unsigned long long)dims[1],
                                         (unsigned long long)dims[2], ld);
  return 0;
}

static void fill_epi(EpiParams& e, const dwn_gemm_desc* d) {
  e.epi = d->epi;
  e.d_bf16 = d->d_dtype == DWN_DT_BF16;
  e.D = d->D;
  e.ldd = d->ldd;
  e.d_zstride = d->d_zstride;
  e.m_limit = d->m_limit > 0 ? d->m_limit : d->M;
  e.n_limit = d->n_limit > 0 ? d->n_limit : d->N;
  e.bias = d->bias;
  e.beta = d->beta;
  e.Tn = d->Tn > 0 ? d->Tn : 1;
  e.n_out_total = d->n_out_total;
  e.row_offset_per_z = d->row_offset_per_z;
}

static int pick_block_n(int N, int b_mn, int d_bf16, int epi) {
  // largest tile <= 256 that divides N (fewer wasted columns); MN-major B needs multiples of 64
  const int step = b_mn ? 64 : 16;
  if (N <= 256) return ((N + step - 1) / step) * step;
  int best = 0;
  for (int bn = 256; bn >= 64; bn -= step)
    if (N % bn == 0) { best = bn; break; }
  if (best >= 128) return best;
  return 256;
}

static int gemm_tc(const dwn_gemm_desc* d, cudaStream_t st) {
  DWN_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->Z > 0, "dwn_gemm: empty problem");
  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.block_n = d->block_n > 0 ? d->block_n : pick_block_n(d->N, d->b_mn, d->d_dtype == DWN_DT_BF16, d->epi);
  DWN_REQUIRE(p.block_n % 16 == 0 && p.block_n <= 256 && (!d->b_mn || p.block_n % 64 == 0), "dwn_gemm: bad block_n %d",
              p.block_n);
  p.b_bytes = ((p.block_n * GT_BK * 2 + 1023) / 1024) * 1024;
  p.num_m_blocks = (d->M + GT_BM - 1) / GT_BM;
  p.num_n_blocks = (d->N + p.block_n - 1) / p.block_n;
  p.Z = d->Z;
  p.a_zmode = d->a_zmode;
  p.b_zmode = d->b_zmode;
  p.b_batch_rows = d->b_batch_rows > 0 ? d->b_batch_rows : 1;
  p.lbo_a = d->dbg_lbo_a; p.sbo_a = d->dbg_sbo_a; p.lbo_b = d->dbg_lbo_b; p.sbo_b = d->dbg_sbo_b;
  fill_epi(p.e, d);
  if (d->epi == 0) {
    const int epv = p.e.d_bf16 ? 8 : 4;
    DWN_REQUIRE(p.e.n_limit % epv == 0 && d->ldd % epv == 0 && ((uintptr_t)d->D & 15) == 0,
                "dwn_gemm: D must be 16-byte vectorisable (n_limit=%d ldd=%ld)", p.e.n_limit, d->ldd);
  }
  const int budget = 227 * 1024 - 1024 - GT_STAGING_BYTES - 256;
  p.stages = budget / (GT_A_BYTES + p.b_bytes);
  if (p.stages > 6) p.stages = 6;
  DWN_REQUIRE(p.stages >= 2, "dwn_gemm: smem budget");
  const size_t smem = 1024 + (size_t)p.stages * (GT_A_BYTES + p.b_bytes) + GT_STAGING_BYTES + 256;

  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int za = d->a_zmode == 1 ? d->Z : 1;
  int zb = d->b_zmode == 1 ? d->Z : 1;
  if (d->b_zmode == 2) zb = (d->M + p.b_batch_rows - 1) / p.b_batch_rows;
  if (make_operand_map(&maps.a[0], d->A, d->a_mn, d->M, d->K, d->lda, d->a_zstride, za, GT_BM)) return -1;
  if (make_operand_map(&maps.b[0], d->B, d->b_mn, d->N, d->K, d->ldb, d->b_zstride, zb, p.block_n)) return -1;
  p.nseg = 1;
  p.seg_k[0] = d->K;
  if (d->split == 3) {
    // fp32-accurate product from three bf16 planes per operand (x = hi + mid + lo, 8 mantissa bits each): the six
    // products whose magnitude is >= 2^-16 of hi*hi, smallest first
    DWN_REQUIRE(d->A2 == nullptr, "dwn_gemm: split and dual-K are exclusive");
    DWN_REQUIRE(d->a_pstride % 8 == 0 && d->b_pstride % 8 == 0, "dwn_gemm: plane strides must be multiples of 8");
    for (int pl = 1; pl < 3; ++pl) {
      if (make_operand_map(&maps.a[pl], (const bf16*)d->A + (size_t)pl * d->a_pstride, d->a_mn, d->M, d->K, d->lda,
                           d->a_zstride, za, GT_BM)) return -1;
      if (make_operand_map(&maps.b[pl], (const bf16*)d->B + (size_t)pl * d->b_pstride, d->b_mn, d->N, d->K, d->ldb,
                           d->b_zstride, zb, p.block_n)) return -1;
    }
    static const signed char sa[6] = {2, 1, 0, 1, 0, 0}, sb[6] = {0, 1, 2, 0, 1, 0};
    p.nseg = 6;
    for (int i = 0; i < 6; ++i) { p.seg_a[i] = sa[i]; p.seg_b[i] = sb[i]; p.seg_k[i] = d->K; }
  } else if (d->A2) {
    DWN_REQUIRE(d->B2 != nullptr && d->K2 > 0, "dwn_gemm: A2 given without B2 / K2");
    if (make_operand_map(&maps.a[1], d->A2, d->a_mn, d->M, d->K2, d->lda2, d->a_zstride, za, GT_BM)) return -1;
    if (make_operand_map(&maps.b[1], d->B2, d->b_mn, d->N, d->K2, d->ldb2, d->b_zstride, zb, p.block_n)) return -1;
    p.nseg = 2;
    p.seg_a[1] = 1; p.seg_b[1] = 1; p.seg_k[1] = d->K2;
  }

  const int total = p.num_m_blocks * p.num_n_blocks * p.Z;
  int grid = dwn_num_sms();
  if (grid > total) grid = total;
#define LAUNCH(AM, BM)                                                                            \
  {                                                                                               \
    auto k = gemm_tc_kernel<AM, BM>;                                                              \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    k<<<grid, GT_THREADS, smem, st>>>(maps, p);                                                   \
  }
  if (d->a_mn) { if (d->b_mn) LAUNCH(1, 1) else LAUNCH(1, 0) } else { if (d->b_mn) LAUNCH(0, 1) else LAUNCH(0, 0) }
#undef LAUNCH
  DWN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SIMT fp32 GEMM (64x64 tile, BK=16, 4x4 micro-tile)
// ------------------------------------------------------------------------------------------------
struct GemmSimtParams {
  const float* A;
  const float* B;
  long a_rs, a_ks, a_zs;  // element strides: row (M), k, z
  long b_rs, b_ks, b_zs;
  int a_kfast, b_kfast;
  int M, N, K;
  int a_zmode, b_zmode, b_batch_rows;
  EpiParams e;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmSimtParams p) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64, z = blockIdx.z;  // M on grid.x: no 65535 limit
  const int za = p.a_zmode == 1 ? z : 0;
  const int zb = p.b_zmode == 1 ? z : (p.b_zmode == 2 ? m0 / p.b_batch_rows : 0);
  const float* A = p.A + (long)za * p.a_zs;
  const float* B = p.B + (long)zb * p.b_zs;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int r, k;
      if (p.a_kfast) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
      float v = 0.f;
      if (m0 + r < p.M && k0 + k < p.K) v = A[(long)(m0 + r) * p.a_rs + (long)(k0 + k) * p.a_ks];
      As[k][r] = v;
      if (p.b_kfast) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
      v = 0.f;
      if (n0 + r < p.N && k0 + k < p.K) v = B[(long)(n0 + r) * p.b_rs + (long)(k0 + k) * p.b_ks];
      Bs[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const EpiParams& e = p.e;
  if (e.epi == 0) {
    float* D = (float*)e.D + (long)z * e.d_zstride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ty * 4 + i;
      if (row >= e.m_limit) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        if (col < e.n_limit) D[(long)row * e.ldd + col] = acc[i][j];
      }
    }
  } else {
    const int rows_valid = min(e.row_offset_per_z, e.n_out_total - z * e.row_offset_per_z);
    float* out = (float*)e.D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ty * 4 + i;
      if (row >= rows_valid || row >= e.m_limit) continue;
      const int gn = z * e.row_offset_per_z + row;
      const float bias = e.bias[gn];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        if (col < e.n_limit) {
          const int b = col / e.Tn, t = col % e.Tn;
          out[((long)b * e.n_out_total + gn) * e.Tn + t] = softplus_beta(acc[i][j] + bias, e.beta);
        }
      }
    }
  }
}

static int gemm_simt(const dwn_gemm_desc* d, cudaStream_t st) {
  DWN_REQUIRE(d->A2 == nullptr, "dwn_gemm(simt): second operand pair unsupported");
  GemmSimtParams p;
  memset(&p, 0, sizeof(p));
  p.A = (const float*)d->A;
  p.B = (const float*)d->B;
  if (d->a_mn) { p.a_rs = 1; p.a_ks = d->lda; } else { p.a_rs = d->lda; p.a_ks = 1; }
  if (d->b_mn) { p.b_rs = 1; p.b_ks = d->ldb; } else { p.b_rs = d->ldb; p.b_ks = 1; }
  p.a_kfast = !d->a_mn;
  p.b_kfast = !d->b_mn;
  p.a_zs = d->a_zstride;
  p.b_zs = d->b_zstride;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_zmode = d->a_zmode; p.b_zmode = d->b_zmode;
  p.b_batch_rows = d->b_batch_rows > 0 ? d->b_batch_rows : 1;
  if (d->b_zmode == 2) DWN_REQUIRE(p.b_batch_rows % 64 == 0, "dwn_gemm(simt): b_batch_rows %% 64 != 0");
  fill_epi(p.e, d);
  DWN_REQUIRE(d->d_dtype == DWN_DT_F32 || d->epi == 1, "dwn_gemm(simt): D must be fp32");
  dim3 grid((d->M + 63) / 64, (d->N + 63) / 64, d->Z);
  DWN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "dwn_gemm(simt): N or Z too large");
  gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
  DWN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dwn_gemm(const dwn_gemm_desc* d, void* stream) {
  if (d->dtype == DWN_DT_BF16) return gemm_tc(d, (cudaStream_t)stream);
  return gemm_simt(d, (cudaStream_t)stream);
}
