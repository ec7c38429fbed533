// ism.cu -- per-proposal template scoring of the Instance Segmentation Model
// (PairwiseSimilarity, ISM/model/loss.py:21-44; compute_semantic_score + best_template_pose,
//  ISM/model/detector.py:198-207,260-296).
//
//   s[p,o,t]  = clamp(cos(q_p, r_{o,t}), 0, 1)            on L2-normalised descriptors
//   S[p,o]    = mean(top5_t s[p,o,:])                     ('avg_5')
//   o*[p]     = argmax_o S[p,o]   (first max),  score[p] = S[p,o*]
//   t*[p]     = argmax_t s[p,o*,:] (first max)
// One CTA per proposal: the query row stays in registers, reference descriptors stream once from L2/HBM with
// coalesced float4 reads (a warp per (o,t) row), reductions by warp shuffles; the reference's P-fold replication of the
// reference descriptors (722 MB at P=200, O=21) never exists.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) template_score_kernel(const float* __restrict__ Qn, const float* __restrict__ Rn, int O, int T,
                                                             int C, float* __restrict__ sim_out, float* __restrict__ obj_score,
                                                             int* __restrict__ best_obj, float* __restrict__ best_score,
                                                             int* __restrict__ best_tmpl) {
  extern __shared__ float sm[];
  float* q = sm;              // C  (query / max(||query||, eps))
  float* sim = q + C;         // O*T
  float* so = sim + O * T;    // O
  __shared__ float red[8];
  __shared__ int s_best;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float eps = 1e-8f;
  // F.cosine_similarity re-normalises its (already unit) inputs: x / max(||x||, 1e-8)
  float ss = 0.f;
  for (int c = tid; c < C; c += 256) { float v = Qn[(size_t)p * C + c]; q[c] = v; ss += v * v; }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float qn = 0.f;
  for (int w = 0; w < 8; ++w) qn += red[w];
  qn = fmaxf(sqrtf(qn), eps);
  __syncthreads();
  for (int c = tid; c < C; c += 256) q[c] = q[c] / qn;
  __syncthreads();
  for (int row = warp; row < O * T; row += 8) {
    const float* r = Rn + (size_t)row * C;
    float dot = 0.f, rr = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(r + c);
      float4 u = *reinterpret_cast<const float4*>(q + c);
      dot = fmaf(u.x, v.x, dot); dot = fmaf(u.y, v.y, dot); dot = fmaf(u.z, v.z, dot); dot = fmaf(u.w, v.w, dot);
      rr = fmaf(v.x, v.x, rr); rr = fmaf(v.y, v.y, rr); rr = fmaf(v.z, v.z, rr); rr = fmaf(v.w, v.w, rr);
    }
    dot = warp_sum(dot); rr = warp_sum(rr);
    if (lane == 0) {
      float s = dot / fmaxf(sqrtf(rr), eps);
      s = fminf(fmaxf(s, 0.f), 1.f);
      sim[row] = s;
      if (sim_out) sim_out[(size_t)p * O * T + row] = s;
    }
  }
  __syncthreads();
  // avg_5: mean of the 5 largest (all T values when T < 5), summed in descending order like topk -> mean
  for (int o = tid; o < O; o += 256) {
    float top[5] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < T; ++t) {
      float v = sim[o * T + t];
      if (v > top[4]) {
        int k = 4;
        while (k > 0 && v > top[k - 1]) { top[k] = top[k - 1]; --k; }
        top[k] = v;
      }
    }
    const int kk = min(T, 5);
    float s = 0.f;
    for (int k = 0; k < kk; ++k) s += top[k];
    s /= (float)kk;
    so[o] = s;
    if (obj_score) obj_score[(size_t)p * O + o] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int o = lane; o < O; o += 32) argmax_first(bv, bi, so[o], o);
    warp_argmax_first(bv, bi);
    if (bi == 0x7fffffff) bi = 0;
    if (lane == 0) { s_best = bi; best_obj[p] = bi; best_score[p] = so[bi]; }
  }
  __syncthreads();
  if (warp == 0) {
    const int o = s_best;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int t = lane; t < T; t += 32) argmax_first(bv, bi, sim[o * T + t], t);
    warp_argmax_first(bv, bi);
    if (lane == 0) best_tmpl[p] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

}  // namespace

// Qn (P,C) and Rn (O,T,C): descriptors already passed through F.normalize (sam6d_l2norm_rows).  C % 4 == 0.
// sim_out (P,O,T) and obj_score (P,O) are optional (may be null).
S6_API int sam6d_template_score(const float* Qn, const float* Rn, int P, int O, int T, int C, float* sim_out, float* obj_score,
                                int* best_obj, float* best_score, int* best_tmpl, void* stream) {
  S6_REQUIRE(Qn && Rn && best_obj && best_score && best_tmpl && P >= 0 && O > 0 && T > 0 && C > 0 && C % 4 == 0);
  if (P == 0) return 0;
  size_t smem = ((size_t)C + (size_t)O * T + O) * sizeof(float);
  S6_REQUIRE(smem <= 200 * 1024);
  S6_CHECK(cudaFuncSetAttribute(template_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  template_score_kernel<<<P, 256, smem, s6_stream(stream)>>>(Qn, Rn, O, T, C, sim_out, obj_score, best_obj, best_score, best_tmpl);
  S6_LAUNCH_CHECK();
  return 0;
}
