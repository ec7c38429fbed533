// coarse.cu -- soft assignment and the hypothesis-and-verify pose initialisation of the coarse stage
// (compute_coarse_Rt, PEM/utils/model_utils.py:187-246).
//
// Pipeline per proposal b (N1 = N2 = n sparse points, the score matrix is (n+1) x (n+1) with row/col 0 = background):
//   1. coarse_assign  : P = softmax_row(A) * softmax_col(A); labels; masked inner block ^1.5      -> W (b, n*n), w1 (b,n)
//   2. coarse_sample  : cdf = cumsum(W) / (sum + 1e-8) (double accumulation like the CPU reference);
//                       idx = searchsorted(cdf, rand)                                              -> (b, 3*n1) pairs
//   3. coarse_hypotheses : 3-point Procrustes per hypothesis + mean residual                      -> Rt (b,n1,12), resid
//   4. coarse_topk    : the n2 smallest residuals (value, then index)                              -> top (b,n2)
//   5. coarse_select  : score = sum(w1) / (sum_i w1_i min_m ||(p_i - t) R - model_m|| + 1e-8); argmax -> init_R, init_t
#include "common.cuh"
#include "svd3.cuh"

namespace {

// ---- 1. soft assignment -------------------------------------------------------------------------------------
// 1024 threads: one CTA per proposal, so the block is the only parallelism there is (every reduction keeps its order: rows are
// reduced by one warp each, columns by one thread each, whatever the block size)
constexpr int ASSIGN_THREADS = 1024;
__global__ void __launch_bounds__(ASSIGN_THREADS) coarse_assign_kernel(const float* __restrict__ A, int S, float* __restrict__ W,
                                                            float* __restrict__ w1out) {
  extern __shared__ float sm[];
  float* a = sm;                 // S*S
  float* rmax = a + S * S;       // S
  float* rsum = rmax + S;
  float* cmax = rsum + S;
  float* csum = cmax + S;
  int* lab1 = (int*)(csum + S);  // S (row labels, index i in 1..S-1)
  int* lab2 = lab1 + S;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Ab = A + (size_t)b * S * S;
  for (int e = tid; e < S * S; e += ASSIGN_THREADS) a[e] = Ab[e];
  __syncthreads();
  // row stats (warp per row) and column stats (thread per column)
  for (int i = warp; i < S; i += ASSIGN_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) m = fmaxf(m, a[i * S + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < S; j += 32) s += expf(a[i * S + j] - m);
    s = warp_sum(s);
    if (lane == 0) { rmax[i] = m; rsum[i] = s; }
  }
  for (int j = tid; j < S; j += ASSIGN_THREADS) {
    float m = -INFINITY;
    for (int i = 0; i < S; ++i) m = fmaxf(m, a[i * S + j]);
    float s = 0.f;
    for (int i = 0; i < S; ++i) s += expf(a[i * S + j] - m);
    cmax[j] = m; csum[j] = s;
  }
  __syncthreads();
  // P in place
  for (int e = tid; e < S * S; e += ASSIGN_THREADS) {
    int i = e / S, j = e - i * S;
    float v = a[e];
    a[e] = (expf(v - rmax[i]) / rsum[i]) * (expf(v - cmax[j]) / csum[j]);
  }
  __syncthreads();
  // labels: first maximal index (torch.max)
  for (int i = warp; i < S; i += ASSIGN_THREADS / 32) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = lane; j < S; j += 32) argmax_first(bv, bi, a[i * S + j], j);
    warp_argmax_first(bv, bi);
    if (lane == 0) lab1[i] = bi;
  }
  for (int j = tid; j < S; j += ASSIGN_THREADS) {
    float bv = -INFINITY; int bi = 0;
    for (int i = 0; i < S; ++i) { float v = a[i * S + j]; if (v > bv) { bv = v; bi = i; } }
    lab2[j] = bi;
  }
  __syncthreads();
  const int n = S - 1;
  float* Wb = W + (size_t)b * n * n;
  for (int e = tid; e < n * n; e += ASSIGN_THREADS) {
    int i = e / n, j = e - i * n;
    float v = a[(i + 1) * S + (j + 1)];
    v = v * (lab1[i + 1] > 0 ? 1.f : 0.f) * (lab2[j + 1] > 0 ? 1.f : 0.f);
    Wb[e] = v * sqrtf(v);   // ** 1.5
  }
  for (int i = tid; i < n; i += ASSIGN_THREADS) w1out[(size_t)b * n + i] = lab1[i + 1] > 0 ? 1.f : 0.f;
}

// ---- 2. cdf + searchsorted ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) coarse_sample_kernel(const float* __restrict__ W, int L, const float* __restrict__ rand,
                                                             int nr, int* __restrict__ idx_out) {
  extern __shared__ float cdf[];  // L floats
  __shared__ double part[1024];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* Wb = W + (size_t)b * L;
  const int per = (L + 1023) / 1024;
  const int beg = min(tid * per, L), end = min(beg + per, L);
  double s = 0.0;
  for (int i = beg; i < end; ++i) s += (double)Wb[i];
  part[tid] = s;
  __syncthreads();
  // exclusive scan of the 1024 partials (Hillis-Steele in double)
  for (int o = 1; o < 1024; o <<= 1) {
    double v = (tid >= o) ? part[tid - o] : 0.0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  double run = (tid == 0) ? 0.0 : part[tid - 1];
  for (int i = beg; i < end; ++i) { run += (double)Wb[i]; cdf[i] = (float)run; }
  __syncthreads();
  const float denom = cdf[L - 1] + 1e-8f;
  __syncthreads();
  for (int i = tid; i < L; i += 1024) cdf[i] = cdf[i] / denom;
  __syncthreads();
  // searchsorted(right=False): first i with cdf[i] >= v; L if none
  for (int r = tid; r < nr; r += 1024) {
    float v = rand[(size_t)b * nr + r];
    int lo = 0, hi = L;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] < v) lo = mid + 1; else hi = mid;
    }
    idx_out[(size_t)b * nr + r] = lo;
  }
}

// ---- 3. hypotheses --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) coarse_hyp_kernel(const int* __restrict__ idx, const float* __restrict__ pts1,
                                                         const float* __restrict__ pts2, int n, int n1, float* __restrict__ Rt,
                                                         float* __restrict__ resid) {
  const int b = blockIdx.y;
  const int hpt = blockIdx.x * blockDim.x + threadIdx.x;
  if (hpt >= n1) return;
  const int* id = idx + ((size_t)b * n1 + hpt) * 3;
  float p1[3][3], p2[3][3];
  int i1s[3], i2s[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int v = id[k];
    int i1 = min(v / n, n - 1), i2 = min(v % n, n - 1);
    i1s[k] = i1; i2s[k] = i2;
    const float* a = pts1 + ((size_t)b * n + i1) * 3;
    const float* c = pts2 + ((size_t)b * n + i2) * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) { p1[k][d] = a[d]; p2[k][d] = c[d]; }
  }
  // The triplet is drawn with replacement (model_utils.py:218-226).  A repeated point in either cloud leaves collinear
  // centred points, i.e. a rank-1 cross-covariance (rank 0 when one cloud contributes a single point): svd3.cuh.
  const int eq1 = (i1s[0] == i1s[1]) + (i1s[0] == i1s[2]) + (i1s[1] == i1s[2]);
  const int eq2 = (i2s[0] == i2s[1]) + (i2s[0] == i2s[2]) + (i2s[1] == i2s[2]);
  const bool rank0 = (eq1 == 3) || (eq2 == 3);
  const bool rank1 = !rank0 && (eq1 + eq2 > 0);
  // weighted_procrustes(src = p2, ref = p1, weights = 1, thresh 0.5, eps 1e-5): w = 1 / (3 + 1e-5)
  const float w = 1.f / (3.f + 1e-5f);
  float cs[3], cr[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    cs[d] = p2[0][d] * w + p2[1][d] * w + p2[2][d] * w;
    cr[d] = p1[0][d] * w + p1[1][d] * w + p1[2][d] * w;
  }
  double H[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += (double)(p2[k][i] - cs[i]) * (double)(w * (p1[k][j] - cr[j]));
      H[i][j] = s;
    }
  double Rd[3][3];
  if (rank0) {   // H is zero up to the 1e-5 of the weight normalisation: the reference's svd(0) gives U = V = I
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rd[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    procrustes_rotation(H, Rd, rank1);
  }
  float R[3][3], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = (float)Rd[i][j];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = cr[i] - (R[i][0] * cs[0] + R[i][1] * cs[1] + R[i][2] * cs[2]);
  // residual: mean_k || (p1_k - t) R - p2_k ||
  float rs = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = p1[k][0] - t[0], y = p1[k][1] - t[1], z = p1[k][2] - t[2];
    float ex = x * R[0][0] + y * R[1][0] + z * R[2][0] - p2[k][0];
    float ey = x * R[0][1] + y * R[1][1] + z * R[2][1] - p2[k][1];
    float ez = x * R[0][2] + y * R[1][2] + z * R[2][2] - p2[k][2];
    rs += sqrtf(ex * ex + ey * ey + ez * ez);
  }
  float* o = Rt + ((size_t)b * n1 + hpt) * 12;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i * 3 + j] = R[i][j];
  o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
  resid[(size_t)b * n1 + hpt] = rs / 3.f;
}

// ---- 4. top-k smallest (bitonic sort of (value, index) keys in shared memory) ---------------------------------
__global__ void __launch_bounds__(1024) topk_smallest_kernel(const float* __restrict__ v, int n, int npow2, int k,
                                                             int* __restrict__ out) {
  extern __shared__ unsigned long long keys[];  // npow2
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < npow2; i += 1024) {
    unsigned long long key = ~0ull;
    if (i < n) {
      float f = v[(size_t)b * n + i];
      unsigned u = __float_as_uint(f);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // order-preserving map; NaN sorts last-ish
      key = ((unsigned long long)u << 32) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npow2 / 2; i += 1024) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool asc = ((lo & size) == 0);
        unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == asc) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += 1024) out[(size_t)b * k + i] = (int)(keys[i] & 0xffffffffu);
}

// ---- 5. pose selection ----------------------------------------------------------------------------------------
// grid = (ceil(n2 / SEL_PP), B); a thread owns one point of pts1 under SEL_PP hypotheses; the CAD samples sit in shared memory as
// (x, y, z, |m|^2) quadruples, so the inner loop is one 16-byte broadcast load and 4 instructions per (hypothesis, point, sample):
// min_m (|x|^2 - 2 x.m + |m|^2) = |x|^2 + min_m (|m|^2 - 2 x.m), the clamp at 0 commutes with the minimum
// (pairwise_distance, model_utils.py:98-111).
// (Measured alternatives: a packed-fp32 FFMA2 variant with 8 hypotheses per CTA, 550 us against 362 us for this one -- FFMA2 issues at
// half rate; a uniform 8^3 grid over the CAD samples walked shell by shell per thread, bit-identical scores but 3.9 ms -- the
// walks of a warp's 32 points diverge, and at 1024 samples the regular scan below is only ~110 warp instructions per point.)
constexpr int SEL_PP = 4, SEL_THREADS = 224;
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__global__ void __launch_bounds__(SEL_THREADS) coarse_select_kernel(const float* __restrict__ Rt, const int* __restrict__ top, int n1,
                                                                    int n2, const float* __restrict__ pts1, const float* __restrict__ w1,
                                                                    int n, const float* __restrict__ model, int nm,
                                                                    float* __restrict__ scores) {
  extern __shared__ float4 smq[];   // nm quadruples
  __shared__ float red[2 * SEL_PP][SEL_THREADS / 32];
  __shared__ float rts[SEL_PP][12];
  const int pose0 = blockIdx.x * SEL_PP, b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < nm; i += SEL_THREADS) {
    const float* q = model + ((size_t)b * nm + i) * 3;
    const float x = q[0], y = q[1], z = q[2];
    smq[i] = make_float4(x, y, z, x * x + y * y + z * z);
  }
  if (tid < SEL_PP * 12) {
    const int pp = tid / 12, e = tid - pp * 12;
    const int pose = min(pose0 + pp, n2 - 1);
    rts[pp][e] = Rt[((size_t)b * n1 + top[(size_t)b * n2 + pose]) * 12 + e];
  }
  __syncthreads();
  float num = 0.f, den[SEL_PP];
#pragma unroll
  for (int pp = 0; pp < SEL_PP; ++pp) den[pp] = 0.f;
  for (int i = tid; i < n; i += SEL_THREADS) {
    const float* p = pts1 + ((size_t)b * n + i) * 3;
    const float px = p[0], py = p[1], pz = p[2];
    float ax[SEL_PP], ay[SEL_PP], az[SEL_PP], x2[SEL_PP], best[SEL_PP];
#pragma unroll
    for (int pp = 0; pp < SEL_PP; ++pp) {
      const float* R = rts[pp];
      const float x = px - R[9], y = py - R[10], z = pz - R[11];
      const float tx = x * R[0] + y * R[3] + z * R[6];
      const float ty = x * R[1] + y * R[4] + z * R[7];
      const float tz = x * R[2] + y * R[5] + z * R[8];
      x2[pp] = tx * tx + ty * ty + tz * tz;
      ax[pp] = -2.f * tx; ay[pp] = -2.f * ty; az[pp] = -2.f * tz;
      best[pp] = INFINITY;
    }
    // two CAD samples per step and a three-input minimum (FMNMX3): 3.5 instead of 4 issue slots per (hypothesis, point, sample);
    // a minimum is exact, so the grouping does not change the result
    int m = 0;
#pragma unroll 2
    for (; m + 1 < nm; m += 2) {
      const float4 q = smq[m], q2 = smq[m + 1];
#pragma unroll
      for (int pp = 0; pp < SEL_PP; ++pp)
        best[pp] = min3f(best[pp], fmaf(ax[pp], q.x, fmaf(ay[pp], q.y, fmaf(az[pp], q.z, q.w))),
                         fmaf(ax[pp], q2.x, fmaf(ay[pp], q2.y, fmaf(az[pp], q2.z, q2.w))));
    }
    if (m < nm) {
      const float4 q = smq[m];
#pragma unroll
      for (int pp = 0; pp < SEL_PP; ++pp) best[pp] = fminf(best[pp], fmaf(ax[pp], q.x, fmaf(ay[pp], q.y, fmaf(az[pp], q.z, q.w))));
    }
    const float wi = w1[(size_t)b * n + i];
    num += wi;
#pragma unroll
    for (int pp = 0; pp < SEL_PP; ++pp) den[pp] += sqrtf(fmaxf(x2[pp] + best[pp], 0.f)) * wi;
  }
  num = warp_sum(num);
#pragma unroll
  for (int pp = 0; pp < SEL_PP; ++pp) den[pp] = warp_sum(den[pp]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int pp = 0; pp < SEL_PP; ++pp) { red[pp][tid >> 5] = num; red[SEL_PP + pp][tid >> 5] = den[pp]; }
  }
  __syncthreads();
  if (tid < SEL_PP && pose0 + tid < n2) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < SEL_THREADS / 32; ++w) { a += red[tid][w]; c += red[SEL_PP + tid][w]; }
    scores[(size_t)b * n2 + pose0 + tid] = a / (c + 1e-8f);
  }
}

__global__ void coarse_pick_kernel(const float* __restrict__ scores, const int* __restrict__ top, const float* __restrict__ Rt,
                                   int n1, int n2, float* __restrict__ R, float* __restrict__ t) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int i = lane; i < n2; i += 32) argmax_first(bv, bi, scores[(size_t)b * n2 + i], i);
  warp_argmax_first(bv, bi);
  if (bi == 0x7fffffff) bi = 0;   // all-NaN row: keep the first hypothesis
  const float* rt = Rt + ((size_t)b * n1 + top[(size_t)b * n2 + bi]) * 12;
  if (lane < 9) R[(size_t)b * 9 + lane] = rt[lane];
  if (lane < 3) t[(size_t)b * 3 + lane] = rt[9 + lane];
}

}  // namespace

// A (B,S,S) f32 -> W (B,(S-1)^2) masked soft assignment ^1.5, w1 (B,S-1)      (model_utils.py:206-216)
S6_API int sam6d_coarse_assign(const float* A, int B, int S, float* W, float* w1, void* stream) {
  S6_REQUIRE(A && W && w1 && B >= 0 && S >= 2);
  if (B == 0) return 0;
  size_t smem = ((size_t)S * S + 6 * S) * sizeof(float);
  S6_REQUIRE(smem <= 220 * 1024);
  S6_CHECK(cudaFuncSetAttribute(coarse_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coarse_assign_kernel<<<B, ASSIGN_THREADS, smem, s6_stream(stream)>>>(A, S, W, w1);
  S6_LAUNCH_CHECK();
  return 0;
}

// W (B,L) f32, rand (B,nr) f32 in [0,1) -> idx (B,nr) i32 in [0,L]             (model_utils.py:218-220)
S6_API int sam6d_coarse_sample(const float* W, int B, int L, const float* rand, int nr, int* idx, void* stream) {
  S6_REQUIRE(W && rand && idx && B >= 0 && L > 0 && nr >= 0);
  if (B == 0 || nr == 0) return 0;
  size_t smem = (size_t)L * sizeof(float);
  S6_REQUIRE(smem <= 200 * 1024);
  S6_CHECK(cudaFuncSetAttribute(coarse_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coarse_sample_kernel<<<B, 1024, smem, s6_stream(stream)>>>(W, L, rand, nr, idx);
  S6_LAUNCH_CHECK();
  return 0;
}

// idx (B,n1,3) flat pair indices -> Rt (B,n1,12) = [R row-major, t], resid (B,n1)   (model_utils.py:221-234)
S6_API int sam6d_coarse_hypotheses(const int* idx, const float* pts1, const float* pts2, int B, int n, int n1, float* Rt,
                                   float* resid, void* stream) {
  S6_REQUIRE(idx && pts1 && pts2 && Rt && resid && B >= 0 && n > 0 && n1 >= 0);
  if (B == 0 || n1 == 0) return 0;
  dim3 grid(s6_cdiv(n1, 128), B);
  coarse_hyp_kernel<<<grid, 128, 0, s6_stream(stream)>>>(idx, pts1, pts2, n, n1, Rt, resid);
  S6_LAUNCH_CHECK();
  return 0;
}

// v (B,n) -> out (B,k): indices of the k smallest values, ascending by (value, index)   (model_utils.py:235)
S6_API int sam6d_topk_smallest(const float* v, int B, int n, int k, int* out, void* stream) {
  S6_REQUIRE(v && out && B >= 0 && n > 0 && k > 0 && k <= n);
  if (B == 0) return 0;
  int npow2 = 2;
  while (npow2 < n) npow2 <<= 1;
  size_t smem = (size_t)npow2 * sizeof(unsigned long long);
  S6_REQUIRE(smem <= 200 * 1024);
  S6_CHECK(cudaFuncSetAttribute(topk_smallest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_smallest_kernel<<<B, 1024, smem, s6_stream(stream)>>>(v, n, npow2, k, out);
  S6_LAUNCH_CHECK();
  return 0;
}

// scores the n2 retained hypotheses against the CAD samples and returns the best (model_utils.py:239-246)
// scratch: scores (B,n2) f32
S6_API int sam6d_coarse_select(const float* Rt, const int* top, int B, int n1, int n2, const float* pts1, const float* w1, int n,
                               const float* model, int nm, float* scores, float* R, float* t, void* stream) {
  S6_REQUIRE(Rt && top && pts1 && w1 && model && scores && R && t && B >= 0 && n > 0 && nm > 0 && n2 > 0);
  if (B == 0) return 0;
  size_t smem = (size_t)nm * 4 * sizeof(float);
  S6_REQUIRE(smem <= 200 * 1024);
  S6_CHECK(cudaFuncSetAttribute(coarse_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(s6_cdiv(n2, SEL_PP), B);
  coarse_select_kernel<<<grid, SEL_THREADS, smem, s6_stream(stream)>>>(Rt, top, n1, n2, pts1, w1, n, model, nm, scores);
  S6_LAUNCH_CHECK();
  coarse_pick_kernel<<<B, 32, 0, s6_stream(stream)>>>(scores, top, Rt, n1, n2, R, t);
  S6_LAUNCH_CHECK();
  return 0;
}
