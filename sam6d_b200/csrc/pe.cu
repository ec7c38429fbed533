// pe.cu -- PositionalEncoding of the fine stage (PEM/model/fine_point_matching.py:90-125):
//   ball query -> grouped [xyz_j - xyz_i, xyz_j] -> SharedMLP 6->32->64->128 (1x1 conv, BatchNorm(eval), ReLU) -> max over samples.
// The grouped (B,6,N,ns) tensor and the (B,128,N,ns) activations of the reference are never materialised:
// one thread owns one (point, sample) row, keeps the hidden vectors in registers and the three folded weight
// matrices are broadcast from shared memory; padded duplicate samples (which cannot change a max) are skipped.
// This file is the exact fp32 CUDA-core version.
#include "common.cuh"

namespace {

template <int NS>
__global__ void __launch_bounds__(256) pe_mlp_max_kernel(const float* __restrict__ pts, const int* __restrict__ idx,
                                                         const int* __restrict__ cnt, int N,
                                                         const float* __restrict__ W1, const float* __restrict__ B1,   // 32x6, 32
                                                         const float* __restrict__ W2, const float* __restrict__ B2,   // 64x32, 64
                                                         const float* __restrict__ W3, const float* __restrict__ B3,   // 128x64, 128
                                                         float* __restrict__ out, int out_ld, int out_off) {
  __shared__ __align__(16) float w1s[32 * 8];   // rows padded to 8
  __shared__ __align__(16) float w2s[64 * 32];
  __shared__ __align__(16) float w3s[128 * 64];
  __shared__ float b1s[32], b2s[64], b3s[128];
  __shared__ float red[8][32];
  constexpr int PPB = 256 / NS;  // points per CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 32 * 8; e += 256) { int r = e >> 3, c = e & 7; w1s[e] = (c < 6) ? W1[r * 6 + c] : 0.f; }
  for (int e = tid; e < 64 * 32; e += 256) w2s[e] = W2[e];
  for (int e = tid; e < 128 * 64; e += 256) w3s[e] = W3[e];
  if (tid < 32) b1s[tid] = B1[tid];
  if (tid < 64) b2s[tid] = B2[tid];
  if (tid < 128) b3s[tid] = B3[tid];
  __syncthreads();

  const int b = blockIdx.y;
  const int pl = tid / NS, s = tid % NS;
  const int i = blockIdx.x * PPB + pl;
  const bool pvalid = i < N;
  int c = pvalid ? cnt[(size_t)b * N + i] : 0;
  c = max(c, 1);                       // empty ball: the padded index list is all zeros -> one sample, point 0
  const bool active = pvalid && s < c;
  float h2[64];
  if (active) {
    const float* pi = pts + ((size_t)b * N + i) * 3;
    const int j = idx[((size_t)b * N + i) * NS + s];
    const float* pj = pts + ((size_t)b * N + j) * 3;
    float x[6] = {pj[0] - pi[0], pj[1] - pi[1], pj[2] - pi[2], pj[0], pj[1], pj[2]};
    float h1[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      float4 wa = *reinterpret_cast<const float4*>(&w1s[o * 8]);
      float2 wb = *reinterpret_cast<const float2*>(&w1s[o * 8 + 4]);
      float a = b1s[o];
      a = fmaf(wa.x, x[0], a); a = fmaf(wa.y, x[1], a); a = fmaf(wa.z, x[2], a);
      a = fmaf(wa.w, x[3], a); a = fmaf(wb.x, x[4], a); a = fmaf(wb.y, x[5], a);
      h1[o] = fmaxf(a, 0.f);
    }
#pragma unroll 4
    for (int o = 0; o < 64; ++o) {
      float a = b2s[o];
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 w = *reinterpret_cast<const float4*>(&w2s[o * 32 + k]);
        a = fmaf(w.x, h1[k], a); a = fmaf(w.y, h1[k + 1], a); a = fmaf(w.z, h1[k + 2], a); a = fmaf(w.w, h1[k + 3], a);
      }
      h2[o] = fmaxf(a, 0.f);
    }
  }
  // layer 3 in chunks of 32 output channels; max over the NS rows of each point
  for (int oc = 0; oc < 128; oc += 32) {
    float mymax[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) mymax[o] = 0.f;   // ReLU outputs are >= 0, so 0 is neutral for inactive rows
    if (active) {
#pragma unroll 2
      for (int o = 0; o < 32; ++o) {
        float a = b3s[oc + o];
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
          float4 w = *reinterpret_cast<const float4*>(&w3s[(oc + o) * 64 + k]);
          a = fmaf(w.x, h2[k], a); a = fmaf(w.y, h2[k + 1], a); a = fmaf(w.z, h2[k + 2], a); a = fmaf(w.w, h2[k + 3], a);
        }
        mymax[o] = fmaxf(a, 0.f);
      }
    }
    // transpose-reduce max over the 32 lanes of the warp: lane l ends with channel l
    float v = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      float m = warp_max(mymax[o]);
      if (lane == o) v = m;
    }
    if (NS == 32) {
      if (pvalid) out[((size_t)b * N + i) * out_ld + out_off + oc + lane] = v;
    } else {  // NS == 64: two warps per point
      red[warp][lane] = v;
      __syncthreads();
      if ((warp & 1) == 0 && pvalid)
        out[((size_t)b * N + i) * out_ld + out_off + oc + lane] = fmaxf(v, red[warp + 1][lane]);
      __syncthreads();
    }
  }
}

}  // namespace

// pts (B,N,3), idx (B,N,ns) i32 + cnt (B,N) i32 from sam6d_ball_query(pts, pts, ...), folded MLP weights (row-major
// (out,in)), out (B,N,out_ld) f32: channels [out_off, out_off+128) receive max_s MLP([p_j - p_i, p_j]).
S6_API int sam6d_pe_mlp_max(const float* pts, const int* idx, const int* cnt, int B, int N, int ns, const float* W1,
                            const float* B1, const float* W2, const float* B2, const float* W3, const float* B3, float* out,
                            int out_ld, int out_off, void* stream) {
  S6_REQUIRE(pts && idx && cnt && W1 && B1 && W2 && B2 && W3 && B3 && out && B >= 0 && N > 0);
  S6_REQUIRE(ns == 32 || ns == 64);
  if (B == 0) return 0;
  cudaStream_t st = s6_stream(stream);
  if (ns == 32) {
    dim3 grid(s6_cdiv(N, 8), B);
    pe_mlp_max_kernel<32><<<grid, 256, 0, st>>>(pts, idx, cnt, N, W1, B1, W2, B2, W3, B3, out, out_ld, out_off);
  } else {
    dim3 grid(s6_cdiv(N, 4), B);
    pe_mlp_max_kernel<64><<<grid, 256, 0, st>>>(pts, idx, cnt, N, W1, B1, W2, B2, W3, B3, out, out_ld, out_off);
  }
  S6_LAUNCH_CHECK();
  return 0;
}
