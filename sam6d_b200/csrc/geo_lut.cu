// geo_lut.cu -- GeometricStructureEmbedding (PEM/model/transformer.py:334-349) by table interpolation (bf16 token stream).
//
//   E[p,:] = g_d(d_p) + max_{k<3} g_a(a_{p,k}),     g_a(x) = W_a emb(x),   g_d(x) = W_d emb(x) + (b_a + b_d),   p = (cloud, n, m)
//
// emb(x) is the 256-entry sinusoidal embedding of ONE scalar, so g_a and g_d are smooth vector-valued functions of a scalar:
// frequencies <= 1 rad per index unit, angle indices in [0, 12] (angles in [0, pi] / sigma_a), distance indices of points in
// normalised clouds below 12 (the table spans [0, 32): 6.4 object radii).  The reference evaluates them with 4 x 256 sin/cos and two 256 x 256 products PER PAIR (651 GFLOP
// per cloud batch; the tensor-core version of this repo, geo_tc.cu, still spent 1.3 ms per step on it, bound by MUFU, MMA issue and
// its epilogue together).  Here both functions are tabulated once per weight set on a grid of step 1/8 (host side, float64,
// from the fp32 weights: 97 + 257 rows of 256 bf16 = 181 KB) and a pair costs four linear interpolations out of shared memory:
// no sin/cos, no MMA, E written exactly once.  Interpolation error at step 1/8 is < 1e-4 of |E| -- below the bf16 rounding of
// the table and of E itself; measured against the float64 embedding the result is closer than the bf16-operand tensor-core
// product was (rms 1.6e-3 vs 1.8e-3 at |E| ~ 0.56, tools/geo_lut_error.py).
//
// Kernel: persistent, 16 warps per CTA (two pairs in flight per warp), both tables resident in shared memory.  A warp takes 32 consecutive pairs: lane l loads
// the indices of pair l (one coalesced 512-byte read), then for each pair the four indices are broadcast by shuffles and lane l
// interpolates channels [8l, 8l+8) as four packed bf16x2 words per table row (sub / fma / max / add on bf16x2: the same packed
// arithmetic the tensor-core epilogue used), one 16-byte store per lane = one 512-byte row of E per warp instruction.
// Bounds per 64-cloud call: E write 1.27 GB (HBM), 4 KB of table reads per pair (shared-memory bandwidth), ~70 warp
// instructions per pair.
//
// Distance indices outside the table (>= 32): pairs of row 0 / column 0 -- the background point of SAM-6D sits at (100,100,100),
// ~870 index units from everything -- read g_d from `far` (clouds, 2, S, 256), computed exactly (tensor-core distance pass of
// geo_tc.cu) from the 2 S distances of that row and column; any other out-of-range pair takes a slow exact path (sin/cos + a
// 256 x 256 product per pair on CUDA cores, ~5000 warp instructions against ~100 for a table pair: measured on the bench's
// synthetic clouds, whose 20 % gaussian outliers put 3 % of the pairs beyond index 16, a [0, 16) table spent most of its 1.5 ms
// there -- hence the [0, 32) span), so the kernel is correct for any input and fast for the clouds the model produces.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace {


__device__ __forceinline__ uint32_t bf2_sub(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// table position of an index value: row offset (in uint4 units, lane-relative) and interpolation weight.  Computed ONCE per pair by
// the lane that loaded the pair's indices and broadcast by shuffles (per lane and lookup it cost 8 of the ~25 instructions)
__device__ __forceinline__ void lut_pos(float x, float inv_h, int nent, int& row, float& t) {
  const float u = x * inv_h;
  int i = (int)u;
  i = max(0, min(i, nent - 2));
  t = u - (float)i;
  row = i * 32;
}
// linear interpolation between table rows `row` and `row + 1`: lane's 8 channels of g(x), packed bf16x2 arithmetic (t2 = (t, t))
__device__ __forceinline__ uint4 lut_lerp(const uint4* __restrict__ tab, int row, uint32_t t2) {
  const uint4 lo = tab[row], hi = tab[row + 32];
  uint4 r;
  r.x = bf2_fma(t2, bf2_sub(hi.x, lo.x), lo.x);
  r.y = bf2_fma(t2, bf2_sub(hi.y, lo.y), lo.y);
  r.z = bf2_fma(t2, bf2_sub(hi.z, lo.z), lo.z);
  r.w = bf2_fma(t2, bf2_sub(hi.w, lo.w), lo.w);
  return r;
}
// the same interpolation in fp32 (table entries unpacked, no intermediate rounding): 8 channels as floats
__device__ __forceinline__ void lut_lerp_f32(const uint4* __restrict__ tab, int row, float t, float v[8]) {
  const uint4 lo = tab[row], hi = tab[row + 32];
  const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, h[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float l0 = __uint_as_float(l[k] << 16), l1 = __uint_as_float(l[k] & 0xffff0000u);
    const float h0 = __uint_as_float(h[k] << 16), h1 = __uint_as_float(h[k] & 0xffff0000u);
    v[2 * k] = fmaf(t, h0 - l0, l0);
    v[2 * k + 1] = fmaf(t, h1 - l1, l1);
  }
}
__device__ __forceinline__ void unpack8(const uint4& a, float v[8]) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
}

// exact g_d(x) for one pair, warp-cooperative (rare path): lane l evaluates frequencies 4l .. 4l+3, every lane accumulates its 8
// output channels over the 256 embedding entries (weights: W_d^T (in, out) bf16 from global / L2)
__device__ __noinline__ uint4 slow_distance(float x, const float* __restrict__ div_term, const uint4* __restrict__ WdT,
                                            const float* __restrict__ bias, int lane) {
  float e[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) sincosf(x * div_term[lane * 4 + q], &e[2 * q], &e[2 * q + 1]);
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = bias[lane * 8 + c];
  for (int s = 0; s < 32; ++s) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float v = __shfl_sync(0xffffffffu, e[q], s);
      const uint4 w = WdT[(size_t)(s * 8 + q) * 32 + lane];
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(v, __uint_as_float(ww[i] << 16), acc[2 * i]);
        acc[2 * i + 1] = fmaf(v, __uint_as_float(ww[i] & 0xffff0000u), acc[2 * i + 1]);
      }
    }
  }
  return make_uint4(bf2_pack(acc[0], acc[1]), bf2_pack(acc[2], acc[3]), bf2_pack(acc[4], acc[5]), bf2_pack(acc[6], acc[7]));
}

// PRECISE: interpolation, maximum and sum in fp32, ONE rounding to bf16 at the store (rms error 1.2e-3 against 1.6e-3 for the packed
// bf16x2 arithmetic, at about twice the instructions per pair)
// LUT_THREADS / UNROLL: warps per CTA against pairs in flight per warp (the kernel is bound by the shared-memory pipe: ncu shows it
// 60-68 % busy with 32 warps and one pair per iteration, warps waiting on their LDS results)
template <bool PRECISE, int LUT_THREADS, int UNROLL>
__global__ void __launch_bounds__(LUT_THREADS, 1) geo_embed_lut_kernel(const float4* __restrict__ T, long long npairs, int S,
                                                                      const uint4* __restrict__ tabA_g, int na, float inv_ha,
                                                                      const uint4* __restrict__ tabD_g, int nd, float inv_hd,
                                                                      const uint4* __restrict__ far, const float* __restrict__ div_term,
                                                                      const uint4* __restrict__ WdT, const float* __restrict__ bias,
                                                                      uint4* __restrict__ E) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4* tabA = reinterpret_cast<uint4*>(smem_raw);
  uint4* tabD = tabA + na * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int u = tid; u < na * 32; u += LUT_THREADS) tabA[u] = tabA_g[u];
  for (int u = tid; u < nd * 32; u += LUT_THREADS) tabD[u] = tabD_g[u];
  __syncthreads();
  const float d_limit = (float)(nd - 1) / inv_hd;          // distance indices below this are inside the table
  const long long nblocks = (npairs + 31) / 32;
  const long long SS = (long long)S * S;
  for (long long blk = (long long)blockIdx.x * (LUT_THREADS / 32) + warp; blk < nblocks; blk += (long long)gridDim.x * (LUT_THREADS / 32)) {
    const long long base = blk * 32;
    const int cnt = (int)min(32LL, npairs - base);
    float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < cnt) tv = T[base + lane];
    // this lane's pair: table rows and weights of its four indices (a distance outside the table is flagged by row -1)
    int r0, r1, r2, r3;
    float t0, t1, t2, t3;
    lut_pos(tv.x, inv_ha, na, r0, t0);
    lut_pos(tv.y, inv_ha, na, r1, t1);
    lut_pos(tv.z, inv_ha, na, r2, t2);
    lut_pos(tv.w, inv_hd, nd, r3, t3);
    if (!(tv.w < d_limit)) r3 = -1;
    if constexpr (!PRECISE) {                              // the packed variant broadcasts the weight as a bf16x2 word
      t0 = __uint_as_float(bf2_pack(t0, t0)); t1 = __uint_as_float(bf2_pack(t1, t1));
      t2 = __uint_as_float(bf2_pack(t2, t2)); t3 = __uint_as_float(bf2_pack(t3, t3));
    }
    const uint4* tabA_l = tabA + lane;
    const uint4* tabD_l = tabD + lane;
#pragma unroll UNROLL
    for (int j = 0; j < cnt; ++j) {
      const int q0 = __shfl_sync(0xffffffffu, r0, j), q1 = __shfl_sync(0xffffffffu, r1, j);
      const int q2 = __shfl_sync(0xffffffffu, r2, j), q3 = __shfl_sync(0xffffffffu, r3, j);
      const float w0 = __shfl_sync(0xffffffffu, t0, j), w1 = __shfl_sync(0xffffffffu, t1, j);
      const float w2 = __shfl_sync(0xffffffffu, t2, j), w3 = __shfl_sync(0xffffffffu, t3, j);
      uint4 dv;
      const bool in_table = q3 >= 0;                       // warp-uniform: a broadcast value
      if (!in_table) {
        const long long p = base + j;
        const long long c = p / SS;
        const int rem = (int)(p - c * SS), n = rem / S, m = rem - n * S;
        if (n == 0) dv = far[((c * 2 + 0) * S + m) * 32 + lane];
        else if (m == 0) dv = far[((c * 2 + 1) * S + n) * 32 + lane];
        else dv = slow_distance(__shfl_sync(0xffffffffu, tv.w, j), div_term, WdT, bias, lane);
      }
      uint4 o;
      if constexpr (PRECISE) {
        float f0[8], f1[8], f2[8], fd[8];
        lut_lerp_f32(tabA_l, q0, w0, f0);
        lut_lerp_f32(tabA_l, q1, w1, f1);
        lut_lerp_f32(tabA_l, q2, w2, f2);
        if (in_table) lut_lerp_f32(tabD_l, q3, w3, fd);
        else unpack8(dv, fd);
#pragma unroll
        for (int k = 0; k < 8; ++k) fd[k] += fmaxf(fmaxf(f0[k], f1[k]), f2[k]);
        o = make_uint4(bf2_pack(fd[0], fd[1]), bf2_pack(fd[2], fd[3]), bf2_pack(fd[4], fd[5]), bf2_pack(fd[6], fd[7]));
      } else {
        const uint4 v0 = lut_lerp(tabA_l, q0, __float_as_uint(w0));
        const uint4 v1 = lut_lerp(tabA_l, q1, __float_as_uint(w1));
        const uint4 v2 = lut_lerp(tabA_l, q2, __float_as_uint(w2));
        if (in_table) dv = lut_lerp(tabD_l, q3, __float_as_uint(w3));
        o.x = bf2_add(dv.x, bf2_max(bf2_max(v0.x, v1.x), v2.x));
        o.y = bf2_add(dv.y, bf2_max(bf2_max(v0.y, v1.y), v2.y));
        o.z = bf2_add(dv.z, bf2_max(bf2_max(v0.z, v1.z), v2.z));
        o.w = bf2_add(dv.w, bf2_max(bf2_max(v0.w, v1.w), v2.w));
      }
      E[(base + j) * 32 + lane] = o;
    }
  }
}

}  // namespace

// T (clouds*S*S, 4) f32 = (a0, a1, a2, d) indices of every pair; tabA (na, 256) bf16 = W_a emb(i / inv_ha), tabD (nd, 256) bf16 =
// W_d emb(i / inv_hd) + bias; far (clouds, 2, S, 256) bf16 = exact g_d of row 0 ([:,0]) and column 0 ([:,1]) of every cloud;
// div_term (128) f32, WdT (256 in, 256 out) bf16 and bias (256) f32 for the exact fallback of other out-of-table distances;
// precise: 1 = fp32 interpolation with one final rounding, 0 = packed bf16x2 arithmetic  -> E (clouds*S*S, 256) bf16
S6_API int sam6d_geo_embed_lut(const float* T, long long clouds, int S, const void* tabA, int na, float inv_ha, const void* tabD, int nd,
                               float inv_hd, const void* far, const float* div_term, const void* WdT_bf16, const float* bias, void* E,
                               int precise, void* stream) {
  S6_REQUIRE(T && tabA && tabD && far && div_term && WdT_bf16 && bias && E && clouds >= 0 && S > 0 && na >= 2 && nd >= 2);
  S6_REQUIRE(inv_ha > 0.f && inv_hd > 0.f && (long long)(na + nd) * 512 <= 200 * 1024);
  S6_REQUIRE(((reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(tabA) | reinterpret_cast<uintptr_t>(tabD) | reinterpret_cast<uintptr_t>(far) |
               reinterpret_cast<uintptr_t>(WdT_bf16) | reinterpret_cast<uintptr_t>(E)) & 15) == 0);
  const long long npairs = clouds * S * S;
  S6_REQUIRE(npairs < (1LL << 40));
  if (npairs == 0) return 0;
  int dev = 0, sms = 0;
  S6_CHECK(cudaGetDevice(&dev));
  S6_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int smem = (na + nd) * 512;
  // launch shape (SAM6D_GEO_LUT_CFG): 1 = 16 warps x 2 pairs in flight per warp (default: 0.615 ms per 64 clouds); 0 = 32 warps x 1
  // pair (0.666 ms); 2 = 24 warps x 2 pairs (0.671 ms) -- profiles/r02_bench_bf16_v8.json and its cfg lines
  static const int cfg = [] { const char* e = getenv("SAM6D_GEO_LUT_CFG"); return e ? atoi(e) : 1; }();
  const long long nblocks = (npairs + 31) / 32;
  cudaStream_t st = s6_stream(stream);
#define S6_LUT_LAUNCH(P, TH, UN)                                                                                                    \
  do {                                                                                                                               \
    auto kern = geo_embed_lut_kernel<P, TH, UN>;                                                                                     \
    S6_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                                         \
    const long long want = (nblocks + TH / 32 - 1) / (TH / 32);                                                                     \
    const int grid = (int)(want < sms ? want : sms);                                                                                 \
    kern<<<grid, TH, smem, st>>>(reinterpret_cast<const float4*>(T), npairs, S, reinterpret_cast<const uint4*>(tabA), na, inv_ha,     \
                                 reinterpret_cast<const uint4*>(tabD), nd, inv_hd, reinterpret_cast<const uint4*>(far), div_term,  \
                                 reinterpret_cast<const uint4*>(WdT_bf16), bias, reinterpret_cast<uint4*>(E));                       \
  } while (0)
  if (precise) {
    if (cfg == 0) S6_LUT_LAUNCH(true, 1024, 1);
    else if (cfg == 2) S6_LUT_LAUNCH(true, 768, 2);
    else S6_LUT_LAUNCH(true, 512, 2);
  } else {
    if (cfg == 0) S6_LUT_LAUNCH(false, 1024, 1);
    else if (cfg == 2) S6_LUT_LAUNCH(false, 768, 2);
    else S6_LUT_LAUNCH(false, 512, 2);
  }
#undef S6_LUT_LAUNCH
  S6_LAUNCH_CHECK();
  return 0;
}
