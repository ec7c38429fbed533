// ism_geo.cu -- geometric score of the ISM proposals (ISM/model/detector.py:209-258, 311-323; ISM/utils/trimesh_utils.py:77-105;
// ISM/utils/bbox_utils.py:197-221): the CAD samples, rotated by the best template's pose and moved to the centroid of the
// proposal's masked depth, are projected into the image; the IoU of their bounding box with the proposal box is the score.
//
// The reference materialises, per proposal, the masked depth image and three back-projected coordinate images in float64
// (N x H x W x 4 doubles: 630 MB for 64 proposals at 640 x 480) and reduces each with its own pass, then forms the posed and
// projected clouds as (N, 2048, 3) tensors.  Here:
//   query_translation_kernel   one CTA per proposal walks mask x depth once, accumulating sum X, sum Y, sum Z (float64, as the
//                              reference's dtypes make it) and the valid count in registers -> (N,3) float32
//   project_iou_kernel         one CTA per proposal: rotate + translate + project + truncate + clamp each sample in registers,
//                              min / max by warp shuffles, box and IoU by thread 0 (integer areas, one float32 division)
#include "common.cuh"

namespace {

constexpr int TR_THREADS = 512, PJ_THREADS = 256;

__global__ void __launch_bounds__(TR_THREADS) query_translation_kernel(const float* __restrict__ masks, const int* __restrict__ depth, int H, int W,
                                                                       const double* __restrict__ K, double scale, float* __restrict__ translate) {
  __shared__ double red[3][TR_THREADS / 32];
  __shared__ int redn[TR_THREADS / 32];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  const float* m = masks + (size_t)n * H * W;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int cnt = 0;
  const int HW = H * W;
  for (int i = tid; i < HW; i += TR_THREADS) {
    const float md = m[i] * (float)depth[i];                 // float32 mask x int32 depth -> float32
    if (md == 0.f) continue;                                 // Z = 0: not valid whatever the scale
    const double Z = (double)md * scale / 1000.0;
    if (Z > 0.0) {
      const int v = i / W, u = i - v * W;
      sx += ((double)u - cx) * Z / fx;
      sy += ((double)v - cy) * Z / fy;
      sz += Z;
      ++cnt;
    }
  }
  sx = warp_sum_d(sx); sy = warp_sum_d(sy); sz = warp_sum_d(sz);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) { red[0][warp] = sx; red[1][warp] = sy; red[2][warp] = sz; redn[warp] = cnt; }
  __syncthreads();
  if (tid == 0) {
    double X = 0.0, Y = 0.0, Z = 0.0;
    int c = 0;
    for (int w = 0; w < TR_THREADS / 32; ++w) { X += red[0][w]; Y += red[1][w]; Z += red[2][w]; c += redn[w]; }
    // count_nonzero(valid) + 1e-8 is a float32 tensor in the reference: the 1e-8 survives only for an empty mask
    const double den = (double)((float)c + 1e-8f);
    translate[n * 3 + 0] = (float)(X / den);
    translate[n * 3 + 1] = (float)(Y / den);
    translate[n * 3 + 2] = (float)(Z / den);
  }
}

__global__ void __launch_bounds__(PJ_THREADS) project_iou_kernel(const float* __restrict__ poses, const float* __restrict__ pointcloud, int npc,
                                                                 const long long* __restrict__ best_pose, const long long* __restrict__ pred_obj,
                                                                 const float* __restrict__ translate, const double* __restrict__ Kd, int H, int W,
                                                                 const long long* __restrict__ boxes, int* __restrict__ image_vu,
                                                                 int* __restrict__ xyxy, float* __restrict__ iou, unsigned char* __restrict__ ok) {
  __shared__ int red[4][PJ_THREADS / 32];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = poses + (size_t)best_pose[n] * 16;
  const float r00 = P[0], r01 = P[1], r02 = P[2], r10 = P[4], r11 = P[5], r12 = P[6], r20 = P[8], r21 = P[9], r22 = P[10];
  const float t0 = translate[n * 3], t1 = translate[n * 3 + 1], t2 = translate[n * 3 + 2];
  float K[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) K[i] = (float)Kd[i];
  const float* pc = pointcloud + (size_t)pred_obj[n] * npc * 3;
  int umin = INT_MAX, vmin = INT_MAX, umax = INT_MIN, vmax = INT_MIN;
  for (int k = tid; k < npc; k += PJ_THREADS) {
    const float x = pc[k * 3], y = pc[k * 3 + 1], z = pc[k * 3 + 2];
    const float px = fmaf(r02, z, fmaf(r01, y, r00 * x)) + t0;
    const float py = fmaf(r12, z, fmaf(r11, y, r10 * x)) + t1;
    const float pz = fmaf(r22, z, fmaf(r21, y, r20 * x)) + t2;
    const float hu = fmaf(K[2], pz, fmaf(K[1], py, K[0] * px));
    const float hv = fmaf(K[5], pz, fmaf(K[4], py, K[3] * px));
    const float hw = fmaf(K[8], pz, fmaf(K[7], py, K[6] * px));
    int u = __float2int_rz(hu / hw), v = __float2int_rz(hv / hw);      // .to(torch.int): truncation toward zero
    u = min(max(u, 0), W - 1);
    v = min(max(v, 0), H - 1);
    if (image_vu) { image_vu[((size_t)n * npc + k) * 2] = u; image_vu[((size_t)n * npc + k) * 2 + 1] = v; }
    umin = min(umin, u); umax = max(umax, u); vmin = min(vmin, v); vmax = max(vmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o)); vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o)); vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { red[0][warp] = umin; red[1][warp] = vmin; red[2][warp] = umax; red[3][warp] = vmax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < PJ_THREADS / 32; ++w) {
      umin = min(umin, red[0][w]); vmin = min(vmin, red[1][w]); umax = max(umax, red[2][w]); vmax = max(vmax, red[3][w]);
    }
    xyxy[n * 4] = umin; xyxy[n * 4 + 1] = vmin; xyxy[n * 4 + 2] = umax; xyxy[n * 4 + 3] = vmax;
    const long long bx0 = boxes[n * 4], by0 = boxes[n * 4 + 1], bx1 = boxes[n * 4 + 2], by1 = boxes[n * 4 + 3];
    const long long iw = min((long long)umax, bx1) - max((long long)umin, bx0), ih = min((long long)vmax, by1) - max((long long)vmin, by0);
    const long long inter = iw * ih;
    const long long area_a = (long long)(umax - umin) * (vmax - vmin), area_b = (bx1 - bx0) * (by1 - by0);
    ok[n] = (iw > 0 && ih > 0) ? 1 : 0;
    iou[n] = (float)inter / (float)(area_a + area_b - inter);           // integer tensors, true division in float32
  }
}

}  // namespace

// Calculate_the_query_translation (detector.py:237-250) -> depth_image_to_pointcloud_translate_torch (trimesh_utils.py:77-105):
// masks (N,H,W) f32 0/1, depth (H,W) i32, K (3,3) f64 row-major on the device, depth_scale -> translate (N,3) f32
S6_API int sam6d_query_translation(const float* masks, const int* depth, int N, int H, int W, const double* K, double depth_scale,
                                   float* translate, void* stream) {
  S6_REQUIRE(masks && depth && K && translate && N >= 0 && H > 0 && W > 0);
  if (N == 0) return 0;
  query_translation_kernel<<<N, TR_THREADS, 0, s6_stream(stream)>>>(masks, depth, H, W, K, depth_scale, translate);
  S6_LAUNCH_CHECK();
  return 0;
}

// project_template_to_image (detector.py:209-235) + the IoU of compute_geometric_score (detector.py:311-323, bbox_utils.py:197-221):
// poses (T,4,4) f32, pointcloud (O,npc,3) f32, best_pose / pred_obj (N) i64, translate (N,3) f32, K (3,3) f64, boxes (N,4) i64 xyxy
// -> image_vu (N,npc,2) i32 (NULL to skip), xyxy (N,4) i32, iou (N) f32, ok (N) u8 (the proposal's intersection is non-empty)
S6_API int sam6d_project_template_iou(const float* poses, int T, const float* pointcloud, int O, int npc, const long long* best_pose,
                                      const long long* pred_obj, const float* translate, const double* K, int N, int H, int W,
                                      const long long* boxes, int* image_vu, int* xyxy, float* iou, unsigned char* ok, void* stream) {
  S6_REQUIRE(poses && pointcloud && best_pose && pred_obj && translate && K && boxes && xyxy && iou && ok);
  S6_REQUIRE(T > 0 && O > 0 && npc > 0 && N >= 0 && H > 0 && W > 0);
  if (N == 0) return 0;
  project_iou_kernel<<<N, PJ_THREADS, 0, s6_stream(stream)>>>(poses, pointcloud, npc, best_pose, pred_obj, translate, K, H, W, boxes, image_vu,
                                                               xyxy, iou, ok);
  S6_LAUNCH_CHECK();
  return 0;
}
