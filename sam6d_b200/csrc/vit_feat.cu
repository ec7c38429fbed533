// vit_feat.cu -- pixel features of the PEM RGB branch without the (B,256,224,224) map (PEM/model/feature_extraction.py:100-108,
// PEM/utils/model_utils.py:69-81).
//
// The reference reshapes the upscaling Linear's output (B, 14*14, 4*4*C) into a (B, C, 56, 56) map, F.interpolate's it
// bilinearly to the image size (1.6 GB at B = 32) and then gathers C channels at the 2048 chosen pixels of every image.
// Here one warp per chosen pixel reads its four source taps straight from the Linear output -- feature (h, w) of the 56 x 56
// map is the contiguous channel block ((h%4)*4 + w%4) of token (h/4)*14 + w/4 -- and blends them with PyTorch's
// align_corners=False weights (src = max(scale * (dst + 0.5) - 0.5, 0)).
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ld_f(const T* p);
template <>
__device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) bilinear_gather_kernel(const T* __restrict__ up, const long long* __restrict__ choose, int K,
                                                              int G, int sub, int C, int H, int W, float* __restrict__ out) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= K) return;
  const long long pix = choose[(size_t)b * K + k];
  const int Y = (int)(pix / W), X = (int)(pix - (long long)Y * W);
  const int Hs = G * sub;
  const float sh = (float)Hs / (float)H, sw = (float)Hs / (float)W;
  const float sy = fmaxf(sh * ((float)Y + 0.5f) - 0.5f, 0.f), sx = fmaxf(sw * ((float)X + 0.5f) - 0.5f, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Hs - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1, lx1 = sx - (float)x0, lx0 = 1.f - lx1;
  const size_t row = (size_t)sub * sub * C;                    // floats per token
  const T* base = up + (size_t)b * G * G * row;
  auto tap = [&](int h, int w) { return base + ((size_t)(h / sub) * G + (w / sub)) * row + (size_t)((h % sub) * sub + (w % sub)) * C; };
  const T* p00 = tap(y0, x0);
  const T* p01 = tap(y0, x1);
  const T* p10 = tap(y1, x0);
  const T* p11 = tap(y1, x1);
  float* o = out + ((size_t)b * K + k) * C;
  for (int c = lane; c < C; c += 32)
    o[c] = ly0 * (lx0 * ld_f(p00 + c) + lx1 * ld_f(p01 + c)) + ly1 * (lx0 * ld_f(p10 + c) + lx1 * ld_f(p11 + c));
}

}  // namespace

// up: (B, G*G, sub*sub*C) fp32 (up_is_bf16 = 0) or bf16, the output of ViT_AE.output_upscaling; choose (B,K) int64 pixel indices
// y*W + x of the H x W image -> out (B,K,C) fp32 = get_chosen_pixel_feats(F.interpolate(map, (H,W), "bilinear"), choose)
S6_API int sam6d_bilinear_gather(const void* up, int up_is_bf16, const long long* choose, int B, int K, int G, int sub, int C, int H,
                                 int W, float* out, void* stream) {
  S6_REQUIRE(up && choose && out && B >= 0 && K >= 0 && G > 0 && sub > 0 && C > 0 && H > 0 && W > 0);
  if (B == 0 || K == 0) return 0;
  S6_REQUIRE(B <= 65535);
  dim3 grid(s6_cdiv(K, 8), B);
  if (up_is_bf16)
    bilinear_gather_kernel<__nv_bfloat16><<<grid, 256, 0, s6_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(up), choose, K, G, sub,
                                                                             C, H, W, out);
  else
    bilinear_gather_kernel<float><<<grid, 256, 0, s6_stream(stream)>>>(reinterpret_cast<const float*>(up), choose, K, G, sub, C, H, W, out);
  S6_LAUNCH_CHECK();
  return 0;
}
