// rpe_tc.cu -- the relative-position score term of RPEMultiHeadAttention on TMA + tcgen05 (bf16 path).
//
//   s_p[b,h,n,m] = q_h[b,n,:] . proj_p(E[b,n,m,:])_h = u_h[b,n,:] . E[b,n,m,:],   u_h = W_p,h^T q_h      (PEM/model/transformer.py:369-399)
//
// E (clouds, S, S, 256) bf16 is the 1.27 GB pair embedding: this kernel is the one streaming pass over it per self-attention
// layer -- the HBM-bound "PEM attention" kernel of the roofline report.  The CUDA-core version (attn.cu: rpe_scores_kernel) spends
// one LDG.128 + 4 unpack + 32 FMA + a 16-value shuffle transpose per 8 channels and is issue-bound at ~0.8 of the HBM peak.
// Here no CUDA core touches E:
//   * a query row (b, n) owns the contiguous 197 x 256 bf16 block E[b, n, :, :] (100 KB).  The TMA unit drops it into shared
//     memory as four K-major SWIZZLE_128B slabs [keys][64 channels] -- exactly the A operand of a tcgen05.mma with M = keys;
//   * the four per-head folded queries u_h[b,n,:] (4 x 256 bf16 = 2 KB, contiguous in the projection GEMM's output) arrive by TMA
//     as rows 0..3 of a 16-row B operand whose other rows stay zero;
//   * D^T[m, h] = sum_c E[n,m,c] u_h[c]: 2 key tiles x 16 MMAs (M128 N16 K16) per query row into 16 TMEM columns; the tensor
//     pipe is ~15 % busy, shared memory reads ~20 %, the kernel runs at the speed the 2-stage TMA ring pulls E from HBM;
//   * 8 epilogue warps read 4 columns per (key tile, query row) with tcgen05.ld: lane = key, so each store instruction writes
//     32 consecutive fp32 of s_p[b, h, n, :].
// Persistent: one CTA per SM walks a contiguous range of the clouds*S query rows (rows are independent).
#include <cuda.h>

#include "tc.cuh"

namespace {

constexpr int C = 256, KSLABS = C / 64;
constexpr int SLAB_ROWS = 200;                        // keys per slab (S <= 200), 25 swizzle atoms
constexpr int SLAB_BYTES = SLAB_ROWS * 128;           // 25600
constexpr int E_BYTES = KSLABS * SLAB_BYTES;          // 102400
constexpr int NB = 16;                                // MMA N: 4 heads + 12 zero rows
constexpr int U_SLAB = NB * 128, U_BYTES = KSLABS * U_SLAB;   // 8192
constexpr int STAGE_BYTES = E_BYTES + U_BYTES;        // 110592
constexpr int STAGES = 2;
constexpr int GROUP = 8;                              // query rows per TMEM buffer: 2 key tiles x 8 x 16 = 256 columns
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int SMEM = STAGES * STAGE_BYTES + 1024;

__global__ void __launch_bounds__(THREADS, 1) rpe_scores_tc_kernel(const __grid_constant__ CUtensorMap tmE,
                                                                   const __grid_constant__ CUtensorMap tmU, int S, int total_rows,
                                                                   float* __restrict__ SP, int sp_ld) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // contiguous, balanced row range of this CTA
  const int per = total_rows / gridDim.x, rem = total_rows % gridDim.x;
  const int r0 = blockIdx.x * per + min((int)blockIdx.x, rem), nrows = per + ((int)blockIdx.x < rem ? 1 : 0);
  const int ntile = (S + 127) >> 7;                   // key tiles of 128 (M of the MMA)

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], EPI_WARPS * 32); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmE);
    tc::tma_prefetch_desc(&tmU);
  }
  s6_pdl_trigger();
  // rows 4..15 of every B slab are zero for the lifetime of the kernel (the TMA boxes only ever write rows 0..3)
  for (int i = tid; i < STAGES * U_BYTES / 16; i += THREADS) {
    const int s = i / (U_BYTES / 16), o = i - s * (U_BYTES / 16);
    *reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + E_BYTES + o * 16) = make_uint4(0, 0, 0, 0);
  }
  tc::fence_proxy_async_smem();
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  s6_pdl_wait();                                      // u comes from the projection GEMM right before us

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: one query row per stage
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(KSLABS * S * 128 + KSLABS * 4 * 128);
      for (int i = 0; i < nrows; ++i) {
        const int s = i % STAGES, row = r0 + i;
        tc::mbar_wait(&empty_bar[s], (uint32_t)(((i / STAGES) & 1) ^ 1));
        tc::mbar_arrive_expect_tx(&full_bar[s], tx);
        uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
        for (int kb = 0; kb < KSLABS; ++kb) tc::tma_load_2d(&tmE, &full_bar[s], st + kb * SLAB_BYTES, kb * 64, row * S);
#pragma unroll
        for (int kb = 0; kb < KSLABS; ++kb) tc::tma_load_2d(&tmU, &full_bar[s], st + E_BYTES + kb * U_SLAB, kb * 64, row * 4);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(128, NB);
      int g = 0;
      for (int i = 0; i < nrows; ++i) {
        const int s = i % STAGES, j = i % GROUP, buf = g & 1;
        if (j == 0) {
          tc::mbar_wait(&tmem_empty_bar[buf], (uint32_t)(((g >> 1) & 1) ^ 1));
          tc::tc_fence_after_sync();
        }
        tc::mbar_wait(&full_bar[s], (uint32_t)((i / STAGES) & 1));
        tc::tc_fence_after_sync();
        const uint32_t e_addr = tc::smem_u32(smem + s * STAGE_BYTES), u_addr = e_addr + E_BYTES;
        for (int t = 0; t < ntile; ++t) {
          const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256 + t * 128 + j * NB);
#pragma unroll
          for (int kb = 0; kb < KSLABS; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::umma_bf16(d_addr, tc::umma_desc_sw128(e_addr + kb * SLAB_BYTES + t * (128 * 128) + k * 32),
                            tc::umma_desc_sw128(u_addr + kb * U_SLAB + k * 32), idesc, (kb | k) ? 1u : 0u);
        }
        tc::umma_commit(&empty_bar[s]);
        if (j == GROUP - 1 || i == nrows - 1) {
          tc::umma_commit(&tmem_full_bar[buf]);
          ++g;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp -> (key tile, TMEM lane quarter)
    const int ew = warp - 2, quad = warp & 3, t = ew >> 2;
    const int m = t * 128 + quad * 32 + lane;
    const bool tile_on = t < ntile;
    const size_t SS = (size_t)S * sp_ld;
    int g = 0;
    for (int i0 = 0; i0 < nrows; i0 += GROUP, ++g) {
      const int buf = g & 1, cnt = min(GROUP, nrows - i0);
      tc::mbar_wait(&tmem_full_bar[buf], (uint32_t)((g >> 1) & 1));
      tc::tc_fence_after_sync();
      if (tile_on) {
        uint32_t v[GROUP][4];
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + t * 128);
#pragma unroll
        for (int j = 0; j < GROUP; ++j)
          if (j < cnt)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[j][0]), "=r"(v[j][1]), "=r"(v[j][2]), "=r"(v[j][3])
                         : "r"(t_addr + j * NB));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&tmem_empty_bar[buf]);        // values are in registers: the MMA warp may refill the buffer
        if (m < S) {
          int row = r0 + i0;
          int b = row / S, n = row - b * S;
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            if (j < cnt) {
              float* o = SP + ((size_t)b * 4 * S + n) * sp_ld + m;
#pragma unroll
              for (int h = 0; h < 4; ++h) o[h * SS] = __uint_as_float(v[j][h]);
              if (++n == S) { n = 0; ++b; }
            }
          }
        }
      } else {
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&tmem_empty_bar[buf]);
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

// (rows, 256) bf16 row-major, box = {64 channels, box_rows}, 128-byte swizzle
int make_map(CUtensorMap* map, const void* ptr, long long rows, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return 999;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

}  // namespace

// E (B,S,S,256) bf16 contiguous, U (B*S, 4*256) bf16 contiguous (row = the four folded per-head queries of a token)
// -> SP (B,4,S,S) fp32.  S <= 200.  The tcgen05 / TMA form of sam6d_rpe_scores (PEM/model/transformer.py:369-399).
namespace {
int rpe_scores_tc_launch(const void* E, const void* U, int B, int S, float* SP, int sp_ld, void* stream) {
  S6_REQUIRE(E && U && SP && B >= 0 && S > 0 && S <= SLAB_ROWS && sp_ld >= S);
  S6_REQUIRE((reinterpret_cast<uintptr_t>(E) & 15) == 0 && (reinterpret_cast<uintptr_t>(U) & 15) == 0);
  S6_REQUIRE((long long)B * S * S < 2000000000LL);
  if (B == 0) return 0;
  CUtensorMap tmE, tmU;
  int rc = make_map(&tmE, E, (long long)B * S * S, S);
  if (rc) return rc;
  rc = make_map(&tmU, U, (long long)B * S * 4, 4);
  if (rc) return rc;
  int dev = 0, sms = 0;
  S6_CHECK(cudaGetDevice(&dev));
  S6_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total = B * S, grid = total < sms ? total : sms;
  S6_CHECK(cudaFuncSetAttribute(rpe_scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  S6_CHECK(s6_launch_pdl(rpe_scores_tc_kernel, dim3(grid), dim3(THREADS), SMEM, s6_stream(stream), tmE, tmU, S, total, SP, sp_ld));
  S6_LAUNCH_CHECK();
  return 0;
}
}  // namespace

S6_API int sam6d_rpe_scores_tc(const void* E, const void* U, int B, int S, float* SP, void* stream) {
  return rpe_scores_tc_launch(E, U, B, S, SP, S, stream);
}
// the same with padded score rows: SP (B,4,S,sp_ld), sp_ld >= S (columns [S, sp_ld) are left untouched); with sp_ld a multiple
// of 4 the attention kernel can stream the planes with 16-byte copies (sam6d_attn_tc_bias_ld)
S6_API int sam6d_rpe_scores_tc_ld(const void* E, const void* U, int B, int S, float* SP, int sp_ld, void* stream) {
  return rpe_scores_tc_launch(E, U, B, S, SP, sp_ld, stream);
}
