// gemm_tma.cu -- persistent, TMA-fed bf16 GEMM on tcgen05:   C = act(A W^T + bias) (+ R),   A (M,K) bf16, W (N,K) bf16.
//
// One CTA per SM loops over 128 x 256 output tiles (m fastest, so CTAs running together share the same W tile in L2):
//   warp 0   TMA producer : cp.async.bulk.tensor 2-D boxes {64 k, 128 rows} of A and {64 k, 256 rows} of W, SWIZZLE_128B,
//                           straight into the UMMA K-major slabs of a 4-stage ring (48 KB per stage); out-of-range rows /
//                           columns are zero-filled by the TMA unit
//   warp 1   MMA issuer   : 4 x tcgen05.mma M128 N256 K16 per stage into one of two 256-column TMEM accumulators,
//                           tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..  epilogue   : tcgen05.ld -> alpha, bias, activation, residual -> fp32 or bf16, transposed through shared memory
//                           into full-line global stores (epilogue.cuh); overlaps the next tile's loads and MMAs through
//                           the second accumulator
// No register staging and no LSU traffic for the operands: the CUDA-core staged kernel (gemm_tc.cu) tops out near 8 GB/s per
// CTA of operand traffic, this one is bounded by L2 -> SM bandwidth and the tensor pipe.
#include <cuda.h>

#include "epilogue.cuh"
#include "tc.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
// Two shapes of the same kernel (227 KB of shared memory decide the split):
//   deep K  (K >= 1024, the ViT-H Linears): 4-stage ring, 4 epilogue warps -- the MMAs of a tile outlast its epilogue
//   short K (the 256/512-wide PEM Linears): a tile is 4-8 k-blocks, the epilogue sets the pace -> 8 epilogue warps (two per
//           TMEM lane quadrant, 4 column chunks each), 3-stage ring
template <int STAGES, int EW>
struct Cfg {
  static constexpr int kThreads = 64 + EW * 32;
  static constexpr int kEpiBytes = EW * epi::WARP_STAGE_FLOATS * 4;
  static constexpr int kSmem = STAGES * STAGE_BYTES + kEpiBytes + 1024;
};

struct Args {
  const float* bias; const void* R; void* C;   // R has the element type of C
  int M, N, K;
  long long ldc, ldr;
  float alpha;
  int act;
  int batch;                 // independent problems stacked along the rows of A and W (score matrices: one per proposal)
  long long a_rpb, w_rpb;    // rows of A / W per problem (a tile may run into the next problem's rows: masked on store)
  long long c_bs, r_bs;      // element strides of C / R between problems
  // VT kernels: output columns >= vt_col0 are V of an attention layer and are written transposed, as the K-major B operand of
  // the P V MMA: vt[(cloud * vt_C + c) * vt_N1 + token], cloud = row / vt_S (saves the transpose pass over V)
  void* vt; int vt_col0, vt_S, vt_N1, vt_C;
  // ... and columns >= vt_col1 = vt_col0 + vt_C go to a SECOND row-major output c2 (ldc2), column j at c2[row * ldc2 + j - vt_col1]
  // (the folded rel-pos queries of an RPE layer: one launch projects q | k, V^T and u)
  int vt_col1; void* c2; long long ldc2;
};

template <typename OT, int ACT, bool HAS_BIAS, bool HAS_RES, int STAGES, int EW, bool VT = false>
__global__ void __launch_bounds__(64 + EW * 32, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmW, Args g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const long long tpb = (long long)m_tiles * n_tiles, ntiles = tpb * g.batch;
  const int nkb = (g.K + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full_bar[a], 1); tc::mbar_init(&tmem_empty_bar[a], EW * 32); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
  }
  s6_pdl_trigger();
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  s6_pdl_wait();                                   // operands / residual may come from the kernel before us

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      long long gk = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long bz = tile / tpb, t = tile - bz * tpb;
        const int m0 = (int)(t % m_tiles) * BM, n0 = (int)(t / m_tiles) * BN;
        const int a_row = (int)(bz * g.a_rpb) + m0, w_row = (int)(bz * g.w_rpb) + n0;
        for (int kb = 0; kb < nkb; ++kb, ++gk) {
          const int s = (int)(gk % STAGES);
          tc::mbar_wait(&empty_bar[s], (uint32_t)(((gk / STAGES) & 1) ^ 1));
          tc::mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* a_slab = smem + s * STAGE_BYTES;
          tc::tma_load_2d(&tmA, &full_bar[s], a_slab, kb * BK, a_row);
          tc::tma_load_2d(&tmW, &full_bar[s], a_slab + A_BYTES, kb * BK, w_row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN);
      long long gk = 0, it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int acc = (int)(it & 1);
        tc::mbar_wait(&tmem_empty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc::tc_fence_after_sync();
        const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb, ++gk) {
          const int s = (int)(gk % STAGES);
          tc::mbar_wait(&full_bar[s], (uint32_t)((gk / STAGES) & 1));
          tc::tc_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16(d_addr, tc::umma_desc_sw128(a_addr + k * 32), tc::umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
          tc::umma_commit(&empty_bar[s]);
        }
        tc::umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp w -> TMEM lanes 32*(w%4) ..
    const int quad = warp & 3;
    const int c_lo = (EW == 8) ? ((warp - 2) >> 2) * (BN / 64) : 0, c_hi = (EW == 8) ? c_lo + BN / 64 : BN / 32;
    float* stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + (warp - 2) * epi::WARP_STAGE_FLOATS;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = (int)(it & 1);
      const long long bz = tile / tpb, t = tile - bz * tpb;
      const int m0 = (int)(t % m_tiles) * BM, n0 = (int)(t / m_tiles) * BN;
      OT* Cb = reinterpret_cast<OT*>(g.C) + bz * g.c_bs;
      const OT* Rb = reinterpret_cast<const OT*>(g.R) + bz * g.r_bs;
      constexpr bool PRE = HAS_RES && (sizeof(OT) == 2) && (EW == 8);   // short-K tiles with a bf16 residual stream
      uint4 pre[PRE ? 4 : 1][4];
      if constexpr (PRE) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int col0 = n0 + (c_lo + cc) * 32;
          if (epi::chunk_vec_ok<OT, OT, true>(col0, g.N, g.ldc, g.ldr))
            epi::prefetch_res_bf16(reinterpret_cast<const __nv_bfloat16*>(Rb), g.ldr, m0 + quad * 32, g.M, col0, lane, pre[cc]);
        }
      }
      tc::mbar_wait(&tmem_full_bar[acc], (uint32_t)((it >> 1) & 1));
      tc::tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (PRE) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c_lo + cc, col0 = n0 + c * 32;
          if (col0 < g.N) {
            float v[32];
            tc::tmem_ld32(t_addr + c * 32, v);
            if (epi::chunk_vec_ok<OT, OT, true>(col0, g.N, g.ldc, g.ldr))   // two call sites: `pre` stays in registers
              epi::process_chunk<OT, ACT, HAS_BIAS, HAS_RES, OT>(v, stage, lane, m0 + quad * 32, g.M, col0, g.N, g.alpha, g.bias,
                                                                 Rb, g.ldr, Cb,
                                                                 g.ldc, pre[cc]);
            else
              epi::process_chunk<OT, ACT, HAS_BIAS, HAS_RES, OT>(v, stage, lane, m0 + quad * 32, g.M, col0, g.N, g.alpha, g.bias,
                                                                 Rb, g.ldr, Cb,
                                                                 g.ldc);
          }
        }
      } else {
        [[maybe_unused]] const int vrow = m0 + quad * 32 + lane, vcloud = VT ? vrow / g.vt_S : 0, vtok = VT ? vrow - vcloud * g.vt_S : 0;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= g.N) break;
          float v[32];
          tc::tmem_ld32(t_addr + c * 32, v);
          if (VT && col0 >= g.vt_col1) {
            epi::process_chunk<OT, ACT, HAS_BIAS, HAS_RES, OT>(v, stage, lane, m0 + quad * 32, g.M, col0, g.N, g.alpha, g.bias, Rb, g.ldr,
                                                               reinterpret_cast<OT*>(g.c2) - g.vt_col1, g.ldc2);
            continue;
          }
          if (VT && col0 >= g.vt_col0) {
            // lane = token: 32 lanes write 32 adjacent tokens of one channel row (64 bytes) per store
            if (vrow < g.M) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.vt) + ((size_t)vcloud * g.vt_C + (col0 - g.vt_col0)) * g.vt_N1 + vtok;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < g.vt_col1) {
                  float x = v[i] * g.alpha;
                  if constexpr (HAS_BIAS) x += __ldg(g.bias + col0 + i);
                  dst[(size_t)i * g.vt_N1] = __float2bfloat16(epi::act_fn<ACT>(x));
                }
            }
            continue;
          }
          epi::process_chunk<OT, ACT, HAS_BIAS, HAS_RES, OT>(v, stage, lane, m0 + quad * 32, g.M, col0, g.N, g.alpha, g.bias,
                                                             Rb, g.ldr, Cb, g.ldc);
        }
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(&tmem_empty_bar[acc]);
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

// 2-D bf16 row-major (rows, K) tensor with row stride ld (elements); box = {64, box_rows}, 128-byte swizzle
int make_map(CUtensorMap* map, const void* ptr, long long rows, long long K, long long ld, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return 999;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

}  // namespace

namespace {

int launch_gemm_tma(const void* A, const void* W, const float* bias, const void* R, void* C, int c_dtype, int M, int N, int K,
                    long long lda, long long ldw, long long ldc, long long ldr, int batch, long long a_rpb, long long w_rpb,
                    long long c_bs, long long r_bs, float alpha, int act, void* stream, void* vt = nullptr, int vt_col0 = 0, int vt_S = 1,
                    int vt_N1 = 0, int vt_col1 = -1, void* c2 = nullptr, long long ldc2 = 0) {
  S6_REQUIRE(A && W && C && M >= 0 && N > 0 && K > 0 && (K % 8) == 0 && (lda % 8) == 0 && (ldw % 8) == 0 && act >= 0 && act <= 2);
  S6_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && batch >= 0);
  S6_REQUIRE(a_rpb * (long long)batch < 2000000000LL && w_rpb * (long long)batch < 2000000000LL);
  if (M == 0 || batch == 0) return 0;
  CUtensorMap tmA, tmW;
  const long long a_rows = batch > 1 ? a_rpb * (batch - 1) + M : M, w_rows = batch > 1 ? w_rpb * (batch - 1) + N : N;
  int rc = make_map(&tmA, A, a_rows, K, lda, BM);
  if (rc) return rc;
  rc = make_map(&tmW, W, w_rows, K, ldw, BN);
  if (rc) return rc;
  int dev = 0, sms = 0;
  S6_CHECK(cudaGetDevice(&dev));
  S6_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long ntiles = (long long)s6_cdiv(M, BM) * s6_cdiv(N, BN) * batch;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  if (vt_col1 < 0) vt_col1 = N;
  Args g{bias, R, C, M, N, K, ldc, ldr, alpha, act, batch, a_rpb, w_rpb, c_bs, r_bs, vt, vt_col0, vt_S, vt_N1, vt_col1 - vt_col0,
         vt_col1, c2, ldc2};
  cudaStream_t st = s6_stream(stream);
  const bool deep_k = K >= 1024;
#define LAUNCH_ONE(OT, ACT, HB, HR, ST, EWN)                                                                           \
  do {                                                                                                                 \
    auto k = gemm_tma_kernel<OT, ACT, HB, HR, ST, EWN>;                                                                \
    S6_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<ST, EWN>::kSmem));               \
    S6_CHECK(s6_launch_pdl(k, dim3(grid), dim3(Cfg<ST, EWN>::kThreads), Cfg<ST, EWN>::kSmem, st, tmA, tmW, g));        \
  } while (0)
#define LAUNCH_TMA(ACT, HB, HR)                                                                                        \
  do {                                                                                                                 \
    if (c_dtype) {                                                                                                     \
      if (deep_k) LAUNCH_ONE(__nv_bfloat16, ACT, HB, HR, 4, 4); else LAUNCH_ONE(__nv_bfloat16, ACT, HB, HR, 3, 8);    \
    } else {                                                                                                           \
      if (deep_k) LAUNCH_ONE(float, ACT, HB, HR, 4, 4); else LAUNCH_ONE(float, ACT, HB, HR, 3, 8);                    \
    }                                                                                                                  \
  } while (0)
  if (vt) {
    // V^T epilogue: bf16 output, bias, no activation / residual (the QKV and KV projections)
    S6_REQUIRE(c_dtype == 1 && bias && !R && act == 0 && batch == 1 && vt_col0 > 0 && vt_col0 < N && (vt_col0 % 32) == 0 && vt_S > 0 &&
               vt_N1 >= vt_S);
    S6_REQUIRE(vt_col1 > vt_col0 && vt_col1 <= N && (vt_col1 % 32) == 0 &&
               (vt_col1 == N || (c2 && (ldc2 % 8) == 0 && ldc2 >= N - vt_col1 && (reinterpret_cast<uintptr_t>(c2) & 15) == 0)));
    if (deep_k) {
      auto k = gemm_tma_kernel<__nv_bfloat16, 0, true, false, 4, 4, true>;
      S6_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<4, 4>::kSmem));
      S6_CHECK(s6_launch_pdl(k, dim3(grid), dim3(Cfg<4, 4>::kThreads), Cfg<4, 4>::kSmem, st, tmA, tmW, g));
    } else {
      auto k = gemm_tma_kernel<__nv_bfloat16, 0, true, false, 3, 8, true>;
      S6_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<3, 8>::kSmem));
      S6_CHECK(s6_launch_pdl(k, dim3(grid), dim3(Cfg<3, 8>::kThreads), Cfg<3, 8>::kSmem, st, tmA, tmW, g));
    }
    S6_LAUNCH_CHECK();
    return 0;
  }
  EPI_DISPATCH(act, bias, R, LAUNCH_TMA);
#undef LAUNCH_TMA
#undef LAUNCH_ONE
  S6_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// A (M,K) bf16 lda, W (N,K) bf16 ldw, C (M,N) fp32 (c_dtype 0) or bf16 (1), bias (N) fp32 or NULL, R (M,N) or NULL: the
// residual has the element type of C (fp32 stream with fp32 output, bf16 stream with bf16 output).
// K % 8 == 0, lda % 8 == 0, ldw % 8 == 0, 16-byte aligned bases.  act: 0 none, 1 ReLU, 2 GELU(erf).
S6_API int sam6d_gemm_tma(const void* A, const void* W, const float* bias, const void* R, void* C, int c_dtype, int M, int N, int K,
                          long long lda, long long ldw, long long ldc, long long ldr, float alpha, int act, void* stream) {
  return launch_gemm_tma(A, W, bias, R, C, c_dtype, M, N, K, lda, ldw, ldc, ldr, 1, 0, 0, 0, 0, alpha, act, stream);
}

// `batch` independent problems C_z = act(alpha A_z W_z^T + bias) (+ R_z): problem z reads rows [z*a_rpb, z*a_rpb + M) of A and
// [z*w_rpb, z*w_rpb + N) of W (w_rpb = 0: shared W) (both matrices are the problems stacked along the rows) and writes C + z*c_bs (elements).
// The cosine score matrices of the matching stages: A = normalised scene tokens, W = normalised template tokens per proposal.
S6_API int sam6d_gemm_tma_batched(const void* A, const void* W, const float* bias, const void* R, void* C, int c_dtype, int M, int N,
                                  int K, long long lda, long long ldw, long long ldc, long long ldr, int batch, long long a_rpb,
                                  long long w_rpb, long long c_bs, long long r_bs, float alpha, int act, void* stream) {
  S6_REQUIRE(a_rpb >= M && (w_rpb >= N || w_rpb == 0));       // w_rpb = 0: one weight matrix shared by every problem
  return launch_gemm_tma(A, W, bias, R, C, c_dtype, M, N, K, lda, ldw, ldc, ldr, batch, a_rpb, w_rpb, c_bs, r_bs, alpha, act, stream);
}

// sam6d_gemm_tma for a fused QKV / KV projection (bf16 output, bias): columns [vt_col0, N) are the values of an attention layer
// and go to Vt instead of C, transposed per cloud of vt_S token rows: Vt[(cloud * (N - vt_col0) + c) * vt_N1 + token] -- the
// layout sam6d_transpose_tokens_bf16 produces and sam6d_attn_tc / sam6d_attn_global_tc consume.  The key-padding columns
// [vt_S, vt_N1) of Vt are not written (the caller keeps them finite, e.g. zeroed once).
S6_API int sam6d_gemm_tma_vt(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, long long lda, long long ldw,
                             long long ldc, void* Vt, int vt_col0, int vt_S, int vt_N1, void* stream) {
  S6_REQUIRE(Vt != nullptr);
  return launch_gemm_tma(A, W, bias, nullptr, C, 1, M, N, K, lda, ldw, ldc, 0, 1, 0, 0, 0, 0, 1.f, 0, stream, Vt, vt_col0, vt_S, vt_N1);
}

// sam6d_gemm_tma_vt with a third column range: [0, vt_col0) -> C, [vt_col0, vt_col1) -> Vt (transposed per cloud), [vt_col1, N) ->
// C2 (M, N - vt_col1) bf16 with row stride ldc2.  One launch for the q | k | v | u projections of an RPE self-attention layer
// (PEM/model/transformer.py:369-394 with proj_p folded into the query).
S6_API int sam6d_gemm_tma_vt2(const void* A, const void* W, const float* bias, void* C, int M, int N, int K, long long lda, long long ldw,
                              long long ldc, void* Vt, int vt_col0, int vt_col1, int vt_S, int vt_N1, void* C2, long long ldc2,
                              void* stream) {
  S6_REQUIRE(Vt != nullptr && C2 != nullptr);
  return launch_gemm_tma(A, W, bias, nullptr, C, 1, M, N, K, lda, ldw, ldc, 0, 1, 0, 0, 0, 0, 1.f, 0, stream, Vt, vt_col0, vt_S, vt_N1, vt_col1,
                         C2, ldc2);
}
