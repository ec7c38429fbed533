// fine_tc.cu -- the dual-softmax assignment of compute_fine_Rt (PEM/utils/model_utils.py:250-283) without the score matrix.
//
// The reference forms A = F1 F2^T / temp ((B, 2049, 2049) fp32, 538 MB at B = 32) and walks it about a dozen times.  Here the
// normalised bf16 tokens are the only inputs and every pass recomputes its score tile on the tensor cores (69 GFLOP per pass,
// ~40 us of tcgen05 time) and reduces it while it is still in TMEM; nothing of size S x S ever reaches HBM.
//
//   pass ROWSUM (mode 0)   inv[b,i] = 1 / sum_j e_ij,  e_ij = exp(alpha * <a_i, b_j> - shift)     (shift = 1/temp >= any score)
//   pass ARGMAX (mode 1)   lab[b,i] = argmax_j P_ij (first maximum),  P_ij = (e_ij * rowf_i) * (e_ij * colf_j)
//   pass ASSIGN (mode 2)   ARGMAX plus  w_i = sum_j P_ij q4_j.w,  pred_i = sum_j P_ij q4_j.xyz / (w_i + 1e-6)   for rows i >= 1
//
// Column sums and column labels are the same passes with the two token matrices swapped (the score matrix of the swapped pair
// is the transpose), so compute_fine_Rt is: ROWSUM(F1,F2), ROWSUM(F2,F1), ARGMAX(F2,F1), masked points, ASSIGN(F1,F2).
//
// One persistent CTA per SM walks work items (cloud b, 128-row tile); the row tile (4 k-blocks of A, 64 KB) stays in shared
// memory while the 256-column tiles of B stream through a 4-stage TMA ring; tcgen05.mma M128 N256 K16 into two TMEM
// accumulators; 8 epilogue warps (two per TMEM lane quadrant, 128 columns each) keep the per-row state in registers across the
// column tiles and merge their halves through shared memory at the end of the item.
#include <cuda.h>

#include "common.cuh"
#include "tc.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, KB = 4, STAGES = 4;       // K = 256 channels
constexpr int A_KB = BM * BK * 2, B_KB = BN * BK * 2;
constexpr int NUM_THREADS = 64 + 256;
constexpr int SMEM_BYTES = KB * A_KB + STAGES * B_KB + 1024;
constexpr float LOG2E = 1.4426950408889634f;

struct FArgs {
  int B, S, mode, ld_f;
  float a2, s2;                 // alpha * log2(e), shift * log2(e)
  const float* row_f;           // (B, ld_f) factor of the rows    (modes 1, 2)
  const float* col_f;           // (B, ld_f) factor of the columns (modes 1, 2)
  const float4* q4;             // (B, ld_f) masked template points (mode 2)
  float* out_inv;               // mode 0: (B, ld_f)
  int* lab;                     // modes 1, 2: (B, S)
  float* wts; float* pred;      // mode 2: (B, S-1), (B, S-1, 3)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) fine_pass_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB, FArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_res = smem;                       // 4 k-block slabs [128][64] of the row tile
  uint8_t* ring = smem + KB * A_KB;            // stages of [256][64]
  __shared__ __align__(8) uint64_t a_full, a_empty, full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float comb[2][BM][6];             // upper-half partial state of each row, double-buffered over items
  __shared__ float sc[2][BN];                  // column factors of the current / next column tile (modes 1, 2)
  __shared__ __align__(16) float4 sq[2][BN];   // masked template points of the tile (mode 2)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (g.S + BM - 1) / BM, n_tiles = (g.S + BN - 1) / BN;
  const int items = g.B * m_tiles;

  if (tid == 0) {
    tc::mbar_init(&a_full, 1); tc::mbar_init(&a_empty, 1);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], 256); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      long long gk = 0;
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int b = item / m_tiles, mt = item - b * m_tiles;
        tc::mbar_wait(&a_empty, (uint32_t)((it & 1) ^ 1));           // the previous item's MMAs no longer read the row tile
        tc::mbar_arrive_expect_tx(&a_full, KB * A_KB);
        for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(&tmA, &a_full, a_res + kb * A_KB, kb * BK, b * g.S + mt * BM);
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int kb = 0; kb < KB; ++kb, ++gk) {
            const int s = (int)(gk % STAGES);
            tc::mbar_wait(&empty_bar[s], (uint32_t)(((gk / STAGES) & 1) ^ 1));
            tc::mbar_arrive_expect_tx(&full_bar[s], B_KB);
            tc::tma_load_2d(&tmB, &full_bar[s], ring + s * B_KB, kb * BK, b * g.S + nt * BN);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN);
      const uint32_t a_addr = tc::smem_u32(a_res);
      long long gk = 0, tcount = 0;
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        tc::mbar_wait(&a_full, (uint32_t)(it & 1));
        tc::tc_fence_after_sync();
        for (int nt = 0; nt < n_tiles; ++nt, ++tcount) {
          const int acc = (int)(tcount & 1);
          tc::mbar_wait(&tmem_empty_bar[acc], (uint32_t)(((tcount >> 1) & 1) ^ 1));
          tc::tc_fence_after_sync();
          const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < KB; ++kb, ++gk) {
            const int s = (int)(gk % STAGES);
            tc::mbar_wait(&full_bar[s], (uint32_t)((gk / STAGES) & 1));
            tc::tc_fence_after_sync();
            const uint32_t b_addr = tc::smem_u32(ring + s * B_KB);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_bf16(d_addr, tc::umma_desc_sw128(a_addr + kb * A_KB + k * 32), tc::umma_desc_sw128(b_addr + k * 32), idesc,
                            (kb | k) ? 1u : 0u);
            tc::umma_commit(&empty_bar[s]);
          }
          tc::umma_commit(&tmem_full_bar[acc]);
        }
        tc::umma_commit(&a_empty);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = (row of the tile, column half)
    const int quad = warp & 3, half = (warp - 2) >> 2;               // warps 2..9: quadrants 2,3,0,1,2,3,0,1
    const int row = quad * 32 + lane;
    long long tcount = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / m_tiles, mt = item - b * m_tiles;
      const int i = mt * BM + row;                                   // row index inside the cloud
      float rf = 0.f;
      if (MODE != 0 && i < g.S) rf = g.row_f[(size_t)b * g.ld_f + i];
      const float* cf = (MODE != 0) ? g.col_f + (size_t)b * g.ld_f : nullptr;
      const float4* q4 = (MODE == 2) ? g.q4 + (size_t)b * g.ld_f : nullptr;
      float sum = 0.f, bv = -INFINITY, w = 0.f, px = 0.f, py = 0.f, pz = 0.f;
      int bi = 0x7fffffff;
      for (int nt = 0; nt < n_tiles; ++nt, ++tcount) {
        const int acc = (int)(tcount & 1);
        if (MODE != 0) {
          // stage the tile's 256 column factors (and masked points) once per CTA while the MMAs of the tile run; the
          // element loop then reads them as shared-memory broadcasts instead of two dependent global loads per element
          const int et = tid - 64, j = nt * BN + et;
          sc[acc][et] = (j < g.S) ? cf[j] : 0.f;
          if (MODE == 2) sq[acc][et] = (j < g.S) ? q4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          epi_bar();
        }
        tc::mbar_wait(&tmem_full_bar[acc], (uint32_t)((tcount >> 1) & 1));
        tc::tc_fence_after_sync();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * 128);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int j0 = nt * BN + half * 128 + c * 32;
          if (j0 >= g.S) break;                                       // uniform: whole chunk past the last column
          float v[32];
          tc::tmem_ld32(t_addr + c * 32, v);
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int j = j0 + k;
            const float e = (j < g.S) ? ex2(fmaf(v[k], g.a2, -g.s2)) : 0.f;
            if (MODE == 0) {
              sum += e;
            } else {
              const float cj = sc[acc][half * 128 + c * 32 + k];      // 0 past the last column
              const float p = (e * rf) * (e * cj);
              if (p > bv) { bv = p; bi = j; }                         // ascending j inside this thread: first maximum
              if (MODE == 2) {
                const float4 q = sq[acc][half * 128 + c * 32 + k];
                const float pm = p * q.w;
                w += pm; px = fmaf(pm, q.x, px); py = fmaf(pm, q.y, py); pz = fmaf(pm, q.z, pz);
              }
            }
          }
        }
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&tmem_empty_bar[acc]);
      }
      // ---- merge the two column halves of every row (the lower half scanned the lower columns of each tile, but tiles
      // interleave: compare indices explicitly so that the first maximum wins)
      float* cb = comb[it & 1][row];
      if (half == 1) {
        if (MODE == 0) cb[0] = sum;
        else { cb[0] = bv; cb[1] = __int_as_float(bi); if (MODE == 2) { cb[2] = w; cb[3] = px; cb[4] = py; cb[5] = pz; } }
      }
      epi_bar();
      if (half == 0 && i < g.S) {
        if (MODE == 0) {
          g.out_inv[(size_t)b * g.ld_f + i] = 1.f / (sum + cb[0]);
        } else {
          const float ov = cb[0];
          const int oi = __float_as_int(cb[1]);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          if (bi == 0x7fffffff) bi = 0;
          g.lab[(size_t)b * g.S + i] = bi;
          if (MODE == 2 && i >= 1) {
            float ww = w + cb[2], qx = px + cb[3], qy = py + cb[4], qz = pz + cb[5];
            if (bi == 0) { ww = 0.f; qx = 0.f; qy = 0.f; qz = 0.f; }    // background label: the row carries no weight
            const size_t o = (size_t)b * (g.S - 1) + (i - 1);
            const float d = ww + 1e-6f;
            g.wts[o] = ww;
            g.pred[o * 3 + 0] = qx / d; g.pred[o * 3 + 1] = qy / d; g.pred[o * 3 + 2] = qz / d;
          }
        }
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
EncodeFn f_get_encode() {
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
int f_make_map(CUtensorMap* map, const void* ptr, long long rows, int box_rows) {
  EncodeFn enc = f_get_encode();
  if (!enc) return 999;
  cuuint64_t gdim[2] = {256, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {512};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

// masked template points for the ASSIGN pass: q4[b,j] = (pts2[b,j-1], 1) if column j >= 1 carries a non-background label
__global__ void fine_masked_points_kernel(const int* __restrict__ lab2, const float* __restrict__ pts2, int S, int ld, float4* __restrict__ q4) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ld) return;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j >= 1 && j < S && lab2[(size_t)b * S + j] > 0) {
    const float* p = pts2 + ((size_t)b * (S - 1) + (j - 1)) * 3;
    q = make_float4(p[0], p[1], p[2], 1.f);
  }
  q4[(size_t)b * ld + j] = q;
}

}  // namespace

// One pass over the (never materialised) score matrix of Fa (rows) against Fb (columns): both (B*S, 256) bf16, L2-normalised.
// mode 0: out_inv (B,ld_f) = 1 / row sums of exp(alpha <a,b> - shift).
// mode 1: lab (B,S) = row arg-max of P = (e * row_f_i) * (e * col_f_j).
// mode 2: mode 1 plus wts (B,S-1), pred (B,S-1,3) for rows >= 1 from the masked points q4 (B,ld_f) float4.
S6_API int sam6d_fine_pass_tc(const void* Fa, const void* Fb, int B, int S, float alpha, float shift, int mode, const float* row_f,
                              const float* col_f, int ld_f, const float* q4, float* out_inv, int* lab, float* wts, float* pred,
                              void* stream) {
  S6_REQUIRE(Fa && Fb && B >= 0 && S >= 2 && mode >= 0 && mode <= 2 && ld_f >= S);
  S6_REQUIRE(((reinterpret_cast<uintptr_t>(Fa) | reinterpret_cast<uintptr_t>(Fb)) & 15) == 0 && (long long)B * S < 2000000000LL);
  if (mode == 0) S6_REQUIRE(out_inv != nullptr);
  if (mode >= 1) S6_REQUIRE(row_f && col_f && lab);
  if (mode == 2) S6_REQUIRE(q4 && wts && pred && (reinterpret_cast<uintptr_t>(q4) & 15) == 0);
  if (B == 0) return 0;
  CUtensorMap tmA, tmB;
  int rc = f_make_map(&tmA, Fa, (long long)B * S, BM);
  if (rc) return rc;
  rc = f_make_map(&tmB, Fb, (long long)B * S, BN);
  if (rc) return rc;
  int dev = 0, sms = 0;
  S6_CHECK(cudaGetDevice(&dev));
  S6_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int items = B * s6_cdiv(S, BM);
  const int grid = items < sms ? items : sms;
  FArgs g{B, S, mode, ld_f, alpha * LOG2E, shift * LOG2E, row_f, col_f, reinterpret_cast<const float4*>(q4), out_inv, lab, wts, pred};
  cudaStream_t st = s6_stream(stream);
#define FINE_LAUNCH(M)                                                                                              \
  do {                                                                                                              \
    S6_CHECK(cudaFuncSetAttribute(fine_pass_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));   \
    fine_pass_kernel<M><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, g);                                        \
  } while (0)
  if (mode == 0) FINE_LAUNCH(0); else if (mode == 1) FINE_LAUNCH(1); else FINE_LAUNCH(2);
#undef FINE_LAUNCH
  S6_LAUNCH_CHECK();
  return 0;
}

// lab2 (B,S) i32, pts2 (B,S-1,3) -> q4 (B,ld) float4 for sam6d_fine_pass_tc mode 2
S6_API int sam6d_fine_masked_points(const int* lab2, const float* pts2, int B, int S, int ld, float* q4, void* stream) {
  S6_REQUIRE(lab2 && pts2 && q4 && B >= 0 && S >= 2 && ld >= S && (reinterpret_cast<uintptr_t>(q4) & 15) == 0);
  if (B == 0) return 0;
  dim3 grid(s6_cdiv(ld, 256), B);
  fine_masked_points_kernel<<<grid, 256, 0, s6_stream(stream)>>>(lab2, pts2, S, ld, reinterpret_cast<float4*>(q4));
  S6_LAUNCH_CHECK();
  return 0;
}
