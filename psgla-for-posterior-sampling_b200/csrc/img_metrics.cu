// PSNR / SSIM of stacks of samples against the ground-truth image, on the device: replaces the per-sample
// `sample.cpu()` + skimage calls of the reference's post-processing (sampling_images.py:373-384, :411-433), which copy
// ~1000 samples to the host per image.  Definitions follow skimage 0.24 as the reference calls it:
//   PSNR(im, x, data_range=1)                       = 10 log10(R^2 / mean((im - x)^2))
//   ssim(im, x, data_range=1, channel_axis=2)       : 7x7 uniform window, sample covariance (cov_norm = 49/48), K1 = 0.01,
//     K2 = 0.03, SSIM map averaged over the interior [3, H-3) x [3, W-3), then over the channels.
// (skimage is absent offline: parity is against the oracle's restatement, which is itself unpinned -- DESIGN.md.)
#include "common.cuh"

namespace psgla {

constexpr int MT = 32;         // output tile
constexpr int MH = MT + 6;     // tile + 7x7 window halo

// grid: (tiles, C, n).  acc[n][0] += sum of squared error over this tile; acc[n][1] += sum of the SSIM map.
__global__ void __launch_bounds__(256)
psnr_ssim_kernel(int H, int W, const float* __restrict__ stack, const float* __restrict__ ref, float c1, float c2,
                 double* __restrict__ acc) {
  __shared__ float sx[MH][MH + 1], sy[MH][MH + 1];
  __shared__ double red[2][8];
  const int tiles_x = (W + MT - 1) / MT;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c = blockIdx.y, n = blockIdx.z;
  const size_t plane = (size_t)H * W;
  const float* xp = stack + ((size_t)n * gridDim.y + c) * plane;
  const float* rp = ref + (size_t)c * plane;
  const int x0 = tx * MT, y0 = ty * MT;
  // stage tile + halo (window centred on the output pixel: rows y-3..y+3); out-of-image entries are never used
  for (int i = threadIdx.x; i < MH * MH; i += 256) {
    const int r = i / MH, q = i % MH;
    const int gy = y0 - 3 + r, gx = x0 - 3 + q;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    sx[r][q] = in ? xp[(size_t)gy * W + gx] : 0.f;
    sy[r][q] = in ? rp[(size_t)gy * W + gx] : 0.f;
  }
  __syncthreads();
  double sse = 0.0, ssum = 0.0;
  for (int i = threadIdx.x; i < MT * MT; i += 256) {
    const int r = i / MT, q = i % MT;
    const int gy = y0 + r, gx = x0 + q;
    if (gy >= H || gx >= W) continue;
    const float d = sy[r + 3][q + 3] - sx[r + 3][q + 3];
    sse += (double)d * d;
    if (gy < 3 || gy >= H - 3 || gx < 3 || gx >= W - 3) continue;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float u = sy[r + dy][q + dx], v = sx[r + dy][q + dx];  // u: reference image, v: sample
        a += u, b += v;
        aa = fmaf(u, u, aa), bb = fmaf(v, v, bb), ab = fmaf(u, v, ab);
      }
    const float inv = 1.f / 49.f, cov_norm = 49.f / 48.f;
    const float ux = a * inv, uy = b * inv;
    const float vx = cov_norm * (aa * inv - ux * ux), vy = cov_norm * (bb * inv - uy * uy), vxy = cov_norm * (ab * inv - ux * uy);
    const float s = ((2.f * ux * uy + c1) * (2.f * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    ssum += (double)s;
  }
  // block reduction (8 warps)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sse += __shfl_down_sync(0xffffffffu, sse, o);
    ssum += __shfl_down_sync(0xffffffffu, ssum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sse;
    red[1][threadIdx.x >> 5] = ssum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) a += red[0][w], b += red[1][w];
    atomicAdd(&acc[2 * n], a);
    atomicAdd(&acc[2 * n + 1], b);
  }
}

__global__ void psnr_ssim_finish_kernel(int n, double n_px, double n_interior, double range2, const double* __restrict__ acc,
                                        float* __restrict__ psnr, float* __restrict__ ssim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mse = acc[2 * i] / n_px;
  if (psnr) psnr[i] = (float)(10.0 * log10(range2 / mse));  // +inf for identical images, like skimage
  if (ssim) ssim[i] = (float)(acc[2 * i + 1] / n_interior);
}

}  // namespace psgla

using namespace psgla;

extern "C" size_t psgla_img_metrics_workspace_bytes(int n) { return (size_t)(n > 0 ? n : 0) * 2 * sizeof(double); }

extern "C" int psgla_img_psnr_ssim(psgla_img_shape s, const float* stack_dev, const float* ref_dev, float data_range,
                                   void* workspace_dev, size_t workspace_bytes, float* psnr_out_dev, float* ssim_out_dev,
                                   void* stream) {
  PSGLA_REQUIRE(s.B > 0 && s.C > 0 && s.H > 0 && s.W > 0, "psgla_img_psnr_ssim: bad shape");
  PSGLA_REQUIRE(stack_dev && ref_dev && workspace_dev && (psnr_out_dev || ssim_out_dev), "psgla_img_psnr_ssim: null pointer");
  PSGLA_REQUIRE(!ssim_out_dev || (s.H >= 7 && s.W >= 7), "SSIM needs images of at least 7 x 7 (the window), got %d x %d", s.H, s.W);
  PSGLA_REQUIRE(data_range > 0, "data_range must be positive");
  PSGLA_REQUIRE(s.B <= 65535 && s.C <= 65535, "at most 65535 images / channels per call");
  if (workspace_bytes < psgla_img_metrics_workspace_bytes(s.B))
    return set_error(PSGLA_E_WORKSPACE, "workspace of %zu bytes is smaller than the %zu needed", workspace_bytes,
                     psgla_img_metrics_workspace_bytes(s.B));
  cudaStream_t st = (cudaStream_t)stream;
  PSGLA_CUDA_TRY(cudaMemsetAsync(workspace_dev, 0, psgla_img_metrics_workspace_bytes(s.B), st));
  const int tiles = ((s.W + MT - 1) / MT) * ((s.H + MT - 1) / MT);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  psnr_ssim_kernel<<<dim3(tiles, s.C, s.B), 256, 0, st>>>(s.H, s.W, stack_dev, ref_dev, c1, c2, (double*)workspace_dev);
  PSGLA_CUDA_TRY(cudaGetLastError());
  const double n_px = (double)s.C * s.H * s.W;
  const double n_int = (double)s.C * (s.H >= 7 ? s.H - 6 : 0) * (s.W >= 7 ? s.W - 6 : 0);
  psnr_ssim_finish_kernel<<<(s.B + 127) / 128, 128, 0, st>>>(s.B, n_px, n_int > 0 ? n_int : 1.0, (double)data_range * data_range,
                                                            (const double*)workspace_dev, psnr_out_dev, ssim_out_dev);
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}
