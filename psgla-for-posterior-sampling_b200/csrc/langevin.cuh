// The per-element Langevin "pre" arithmetic shared by the stand-alone pre kernels (img_elementwise.cu) and by the last conv
// layer's epilogue, which applies it to the iterate it has just produced (conv_tc.cu, fused post + next pre).
//   PSGLA   Y = X + (delta/lambd) grad + sqrt(2) s Z                      restoration_algorithms.py:232-236
//   PnP-ULA X + delta (-(X - proj)/lambd + grad) + sqrt(2 delta) Z        restoration_algorithms.py:104-115
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace psgla {

// N(0,1) for element e (linear index within one chain's [C][H][W]) of `chain` at `iteration`:
// Philox counter (e >> 2, iteration, chain), component e & 3.
__device__ __forceinline__ void normal_quad(uint64_t seed, uint64_t chain, uint32_t iteration, uint32_t quad,
                                            float (&z)[4]) {
  philox_normal4(seed, chain, quad, iteration, z[0], z[1], z[2], z[3]);
}
__device__ __forceinline__ float normal_at(uint64_t seed, uint64_t chain, uint32_t iteration, uint32_t e) {
  float z[4];
  normal_quad(seed, chain, iteration, e >> 2, z);
  return z[e & 3];
}

// ---- the stream of torch.randn(shape, generator=torch.Generator("cuda").manual_seed(seed)) (restoration_algorithms.py:
// 86-87,104 / :212-213,232), reproduced bit for bit so that a seed alone replays the reference's CUDA noise.
// torch fills a tensor of `numel` floats with a grid-stride kernel of T = 256 * grid threads, each thread owning the Philox
// subsequence t = its global index, started at the generator's offset (in 32-bit outputs, a multiple of 4); one
// curand_normal4 per stride serves elements t + T (4 loop + j), j = 0..3.  curand's Box-Muller: u = x 2^-32 + 2^-33,
// v = y 2^-32 2pi + 2^-33 2pi, s = sqrtf(-2 logf(u)), (s sin v, s cos v) with __sincosf.  Library logf / sqrtf here on
// purpose (the accurate ones torch is built with); the constants are curand_globals.h's rounded literals.
__device__ __forceinline__ float torch_cuda_normal_at(uint64_t seed, uint64_t offset, uint32_t T, uint64_t li) {
  const uint64_t r = li / T;
  const uint32_t t = (uint32_t)(li - r * T);
  const uint64_t ctr = (offset >> 2) + (r >> 2);
  const uint32_t j = (uint32_t)r & 3u;
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = t, c3 = 0;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t a = j < 2 ? c0 : c2, b = j < 2 ? c1 : c3;
  const float u = a * 2.3283064e-10f + (2.3283064e-10f / 2);
  const float v = b * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
  const float s = sqrtf(-2.0f * logf(u));
  float sn, cs;
  __sincosf(v, &sn, &cs);
  return ((j & 1u) ? cs : sn) * s;
}

struct PreArgs {
  int alg;
  float gain_data, noise_scale, proj_gain, c_min, c_max, x_gain, den_in_c3;
  unsigned long long seed;
  long long chain_id0;
  unsigned int iteration;
  int noise_mode;              // PSGLA_NOISE_PHILOX or PSGLA_NOISE_TORCH_CUDA
  unsigned int torch_threads;  // T of the torch launch
  unsigned long long torch_offset;
  long long chw;               // elements of one chain (3 H W): global linear index = b chw + e
};

// The N(0,1) draw of element e of local chain b in the selected stream; `quad` variants serve 4 consecutive elements.
__device__ __forceinline__ float draw_at(const PreArgs& a, int b, uint32_t e) {
  if (a.noise_mode == PSGLA_NOISE_TORCH_CUDA)
    return torch_cuda_normal_at(a.seed, a.torch_offset, a.torch_threads, (uint64_t)b * (uint64_t)a.chw + e);
  return normal_at(a.seed, (unsigned long long)(a.chain_id0 + b), a.iteration, e);
}
__device__ __forceinline__ void draw_quad(const PreArgs& a, int b, uint32_t e0, float (&z)[4]) {  // e0 % 4 == 0
  if (a.noise_mode == PSGLA_NOISE_TORCH_CUDA) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      z[j] = torch_cuda_normal_at(a.seed, a.torch_offset, a.torch_threads, (uint64_t)b * (uint64_t)a.chw + e0 + j);
    return;
  }
  normal_quad(a.seed, (unsigned long long)(a.chain_id0 + b), a.iteration, e0 >> 2, z);
}

__device__ __forceinline__ float langevin_base(const PreArgs& a, float x, float neg_grad_unscaled, float z) {
  // neg_grad_unscaled = mask (x - y)  resp.  A^T(A x - y); the data term is  -gain_data * that.
  float base = fmaf(-a.gain_data, neg_grad_unscaled, fmaf(a.x_gain, x, x));
  if (a.alg == PSGLA_ALG_PNPULA) {
    const float proj = fminf(fmaxf(x, a.c_min), a.c_max);
    base = fmaf(-a.proj_gain, x - proj, base);
  }
  return fmaf(a.noise_scale, z, base);
}

__device__ __forceinline__ void store_nhwc16(__nv_bfloat16* dst_pixel, float c0, float c1, float c2, float c3) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(c0, c1);
  const __nv_bfloat162 b = __floats2bfloat162_rn(c2, c3);
  uint4* d = reinterpret_cast<uint4*>(dst_pixel);
  d[0] = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b), 0u, 0u);
  d[1] = make_uint4(0u, 0u, 0u, 0u);
}

// host: validates a psgla_pre_params and converts it (img_elementwise.cu); `chw` is left 0 for the caller to set
int fill_pre(const psgla_pre_params* p, PreArgs* a);

}  // namespace psgla
