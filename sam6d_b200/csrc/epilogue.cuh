// epilogue.cuh -- shared TMEM epilogue of the tcgen05 GEMMs.
//
// A warp owns 32 accumulator rows (its TMEM lane quadrant); per 32-column chunk:
//   tcgen05.ld (thread = row, 32 columns) -> alpha * acc (+ bias) (activation) (+ residual) -> OT
// Residual loads and output stores go through a per-warp shared-memory transpose (32 x 32 tile, padded rows) so that every
// global access is a fully used 128-byte line: with thread-per-row stores each warp instruction touched 32 different lines
// 16 bytes at a time and the epilogue, not the MMA pipe, set the tile time (1.7 ms vs 0.54 ms on 65536x3840x1280).
// Everything that can be decided at compile time is (activation, bias / residual presence): a runtime activation switch
// if-converts into ~50 predicated erff instructions per element.
#pragma once
#include "tc.cuh"

namespace epi {

constexpr int TILE_LD = 36;                       // floats per staged row: 16-byte aligned, conflict-free for LDS/STS.128
constexpr int WARP_STAGE_FLOATS = 32 * TILE_LD;   // 4.5 KB per warp

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if constexpr (ACT == 1) return fmaxf(x, 0.f);
  if constexpr (ACT == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  return x;
}

// v[32]: this thread's row of the chunk.  stage: this warp's WARP_STAGE_FLOATS floats of shared memory.
// row0: global row of lane 0; col0: first global column of the chunk.  Partial chunks (col0 + 32 > N) take a scalar path.
template <typename T>
struct ident { using type = T; };   // keeps RT out of template argument deduction (callers pass nullptr)
__device__ __forceinline__ float ld_res(const float* p) { return *p; }
__device__ __forceinline__ float ld_res(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename OT, typename RT, bool HAS_RES>
__device__ __forceinline__ bool chunk_vec_ok(int col0, int N, long long ldc, long long ldr) {
  return (col0 + 32 <= N) && ((ldc & (sizeof(OT) == 4 ? 3 : 7)) == 0) && (!HAS_RES || (ldr & (sizeof(RT) == 4 ? 3 : 7)) == 0);
}

// bf16 residual tile of one 32 x 32 chunk in its coalesced register layout: lane l holds 16 B of row (i*8 + l/4), piece l%4.
// Issued before the accumulator is ready, the loads overlap the tile's TMA and MMA time instead of stalling every chunk.
__device__ __forceinline__ void prefetch_res_bf16(const __nv_bfloat16* __restrict__ R, long long ldr, int row0, int M, int col0, int lane,
                                                  uint4 pre[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), q = lane & 3;
    pre[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < M) pre[i] = *reinterpret_cast<const uint4*>(R + (size_t)(row0 + r) * ldr + col0 + q * 8);
  }
}

// RT: residual element type (fp32, or bf16 for the all-bf16 activation flow).  pre: the chunk's residual from
// prefetch_res_bf16 (bf16 residual, vector path only) or nullptr.
template <typename OT, int ACT, bool HAS_BIAS, bool HAS_RES, typename RT = float>
__device__ __forceinline__ void process_chunk(float v[32], float* stage, int lane, int row0, int M, int col0, int N, float alpha,
                                              const float* __restrict__ bias, const typename ident<RT>::type* __restrict__ R, long long ldr,
                                              OT* __restrict__ C, long long ldc, const uint4* pre = nullptr) {
  const bool vec_ok = chunk_vec_ok<OT, RT, HAS_RES>(col0, N, ldc, ldr);
  if (vec_ok) {
    if constexpr (HAS_RES && sizeof(RT) == 2) {
      // bf16 residual tile: a row of the chunk is 64 bytes; lane l reads 16 B of row (i*8 + l/4), piece l%4
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), q = lane & 3;
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (pre) t = pre[i];
        else if (row0 + r < M) t = *reinterpret_cast<const uint4*>(R + (size_t)(row0 + r) * ldr + col0 + q * 8);
        *reinterpret_cast<uint4*>(stage + r * TILE_LD + q * 4) = t;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 t = *reinterpret_cast<const uint4*>(stage + lane * TILE_LD + q * 4);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if constexpr (HAS_BIAS) { b0 = __ldg(reinterpret_cast<const float4*>(bias + col0) + 2 * q); b1 = __ldg(reinterpret_cast<const float4*>(bias + col0) + 2 * q + 1); }
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[q * 8 + 2 * e] = act_fn<ACT>(fmaf(v[q * 8 + 2 * e], alpha, bb[2 * e])) + __uint_as_float(w[e] << 16);
          v[q * 8 + 2 * e + 1] = act_fn<ACT>(fmaf(v[q * 8 + 2 * e + 1], alpha, bb[2 * e + 1])) + __uint_as_float(w[e] & 0xffff0000u);
        }
      }
      __syncwarp();
    } else {
    if constexpr (HAS_RES) {
      // coalesced residual tile -> smem: lane l reads 16 B of row (i*4 + l/8), float4 column l%8
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3), q = lane & 7;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) t = *reinterpret_cast<const float4*>(R + (size_t)(row0 + r) * ldr + col0 + q * 4);
        *reinterpret_cast<float4*>(stage + r * TILE_LD + q * 4) = t;
      }
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (HAS_BIAS) b4 = __ldg(reinterpret_cast<const float4*>(bias + col0) + q);
      if constexpr (HAS_RES) r4 = *reinterpret_cast<const float4*>(stage + lane * TILE_LD + q * 4);
      v[q * 4 + 0] = act_fn<ACT>(fmaf(v[q * 4 + 0], alpha, b4.x)) + r4.x;
      v[q * 4 + 1] = act_fn<ACT>(fmaf(v[q * 4 + 1], alpha, b4.y)) + r4.y;
      v[q * 4 + 2] = act_fn<ACT>(fmaf(v[q * 4 + 2], alpha, b4.z)) + r4.z;
      v[q * 4 + 3] = act_fn<ACT>(fmaf(v[q * 4 + 3], alpha, b4.w)) + r4.w;
    }
    if constexpr (HAS_RES) __syncwarp();
    }
    if constexpr (sizeof(OT) == 4) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stage + lane * TILE_LD + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3), q = lane & 7;
        if (row0 + r < M)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(C) + (size_t)(row0 + r) * ldc + col0 + q * 4) =
              *reinterpret_cast<const float4*>(stage + r * TILE_LD + q * 4);
      }
    } else {
      // bf16: a staged row is 64 bytes = 16 words
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stage + lane * TILE_LD + q * 4) =
            make_uint4(tc::pack_bf16(v[q * 8], v[q * 8 + 1]), tc::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                       tc::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), tc::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), q = lane & 3;
        if (row0 + r < M)
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(C) + (size_t)(row0 + r) * ldc + col0 + q * 8) =
              *reinterpret_cast<const uint4*>(stage + r * TILE_LD + q * 4);
      }
    }
    __syncwarp();
  } else {
    const int row = row0 + lane;
    if (row < M) {
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        if (col < N) {
          float x = v[j] * alpha;
          if constexpr (HAS_BIAS) x += bias[col];
          x = act_fn<ACT>(x);
          if constexpr (HAS_RES) x += ld_res(R + (size_t)row * ldr + col);
          if constexpr (sizeof(OT) == 4) reinterpret_cast<float*>(C)[(size_t)row * ldc + col] = x;
          else reinterpret_cast<__nv_bfloat16*>(C)[(size_t)row * ldc + col] = __float2bfloat16(x);
        }
      }
    }
  }
}

// run-time -> compile-time dispatch of (ACT, HAS_BIAS, HAS_RES)
#define EPI_DISPATCH(ACT_V, BIAS_P, RES_P, ...)                                              \
  do {                                                                                       \
    const int a__ = (ACT_V);                                                                 \
    const bool b__ = (BIAS_P) != nullptr, r__ = (RES_P) != nullptr;                          \
    if (a__ == 0) {                                                                          \
      if (b__) { if (r__) { __VA_ARGS__(0, true, true); } else { __VA_ARGS__(0, true, false); } } \
      else     { if (r__) { __VA_ARGS__(0, false, true); } else { __VA_ARGS__(0, false, false); } } \
    } else if (a__ == 1) {                                                                   \
      if (b__) { if (r__) { __VA_ARGS__(1, true, true); } else { __VA_ARGS__(1, true, false); } } \
      else     { if (r__) { __VA_ARGS__(1, false, true); } else { __VA_ARGS__(1, false, false); } } \
    } else {                                                                                 \
      if (b__) { if (r__) { __VA_ARGS__(2, true, true); } else { __VA_ARGS__(2, true, false); } } \
      else     { if (r__) { __VA_ARGS__(2, false, true); } else { __VA_ARGS__(2, false, false); } } \
    }                                                                                        \
  } while (0)

}  // namespace epi
