// DRUNet denoiser (deepinv.models.DRUNet as constructed at sampling_images.py:136 and called at
// restoration_algorithms.py:238 / sampling_images.py:156): KAIR UNetRes with nc = [64, 128, 256, 512], nb = 4 residual
// blocks per stage, stride-2 2x2 convs down, 2x2 transposed convs up, no biases, ReLU; the input is the image plus a
// constant noise-level channel and the output is the denoised image (architecture restated in oracle/image_oracle.py).
//
// Layer schedule (68 launches): head 4->64 (first-layer SS kernel on the NHWC16 input, channel 3 = sigma), the 16
// full-resolution 64-channel convs on the weight-resident TS kernel (conv_tc.cu), every other layer on the streamed
// implicit-GEMM kernel (conv_gemm.cu); residual and U-Net skip additions are folded into the epilogue of the conv that
// produces the sum; the 64->3 tail runs the fused Langevin "post" epilogue.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv_api.cuh"

namespace psgla {

enum DruKind { DRU_HEAD, DRU_C64, DRU_GEMM3, DRU_DOWN, DRU_UP, DRU_TAIL };

struct DruLayer {
  DruKind kind;
  int cin, cout;
  size_t w_off, bytes;
};

static size_t dru_layer_bytes(DruKind k, int cin, int cout) {
  size_t b = 0;
  switch (k) {
    case DRU_HEAD: b = (size_t)9 * 64 * 16 * 2 + 64 * 4; break;   // swizzled weights + (zero) bias
    case DRU_C64: b = (size_t)9 * 64 * 64 * 2 + 64 * 4; break;
    case DRU_TAIL: b = (size_t)9 * 16 * 64 * 2 + 16 * 4; break;
    case DRU_GEMM3: b = (size_t)9 * cout * cin * 2; break;
    case DRU_DOWN:
    case DRU_UP: b = (size_t)4 * cout * cin * 2; break;
  }
  return (b + 1023) / 1024 * 1024;
}

// The 64 weight tensors in state-dict order: m_head, m_down1.{0..3}.res.{0,2}, m_down1.4, m_down2..., m_down3...,
// m_body.{0..3}.res.{0,2}, m_up3.0, m_up3.{1..4}.res.{0,2}, m_up2..., m_up1..., m_tail.
static const std::vector<DruLayer>& dru_plan() {
  static std::vector<DruLayer> plan;
  if (!plan.empty()) return plan;
  std::vector<DruLayer> v;
  size_t off = 0;
  auto add = [&](DruKind k, int cin, int cout) {
    const size_t b = dru_layer_bytes(k, cin, cout);
    v.push_back({k, cin, cout, off, b});
    off += b;
  };
  const int nc[4] = {64, 128, 256, 512};
  add(DRU_HEAD, 4, 64);
  for (int s = 0; s < 3; ++s) {
    for (int i = 0; i < 8; ++i) add(s == 0 ? DRU_C64 : DRU_GEMM3, nc[s], nc[s]);
    add(DRU_DOWN, nc[s], nc[s + 1]);
  }
  for (int i = 0; i < 8; ++i) add(DRU_GEMM3, 512, 512);
  for (int s = 2; s >= 0; --s) {
    add(DRU_UP, nc[s + 1], nc[s]);
    for (int i = 0; i < 8; ++i) add(s == 0 ? DRU_C64 : DRU_GEMM3, nc[s], nc[s]);
  }
  add(DRU_TAIL, 64, 3);
  plan.swap(v);
  return plan;
}

static inline uint16_t dru_bf16(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

struct DruBuffers {
  uint8_t* buf[4][4];  // [scale][P, A, B, C]
};
static size_t dru_scale_bytes(int B, int H, int W, int s) {
  const size_t n = (size_t)B * (H >> s) * (W >> s) * (64 << s) * 2;
  return (n + 1023) / 1024 * 1024;
}
static size_t dru_workspace_bytes(int B, int H, int W) {
  size_t t = 0;
  for (int s = 0; s < 4; ++s) t += 4 * dru_scale_bytes(B, H, W, s);
  return t + 1024;
}

}  // namespace psgla

using namespace psgla;

extern "C" int psgla_drunet_num_weights(void) { return (int)dru_plan().size(); }

extern "C" size_t psgla_drunet_packed_bytes(void) {
  const auto& plan = dru_plan();
  return plan.back().w_off + plan.back().bytes;
}

extern "C" int psgla_drunet_pack_weights(const float* const* weights_host, void* packed_dev, void* stream) {
  PSGLA_REQUIRE(weights_host && packed_dev, "psgla_drunet_pack_weights: null pointer");
  const auto& plan = dru_plan();
  std::vector<uint8_t> host(psgla_drunet_packed_bytes(), 0);
  for (size_t l = 0; l < plan.size(); ++l) {
    const DruLayer& L = plan[l];
    const float* w = weights_host[l];
    PSGLA_REQUIRE(w != nullptr, "DRUNet weight tensor %zu is null", l);
    uint8_t* dst = host.data() + L.w_off;
    switch (L.kind) {
      case DRU_HEAD: pack_conv3x3_swizzled(w, 64, 4, 64, 16, dst); break;
      case DRU_C64: pack_conv3x3_swizzled(w, 64, 64, 64, 64, dst); break;
      case DRU_TAIL: pack_conv3x3_swizzled(w, 3, 64, 16, 64, dst); break;
      case DRU_GEMM3:  // OIHW [cout][cin][3][3] -> [tap][cout][cin]
      case DRU_DOWN: {  // OIHW [cout][cin][2][2] -> [tap][cout][cin]
        const int taps = L.kind == DRU_GEMM3 ? 9 : 4;
        uint16_t* d = reinterpret_cast<uint16_t*>(dst);
        for (int n = 0; n < L.cout; ++n)
          for (int k = 0; k < L.cin; ++k)
            for (int t = 0; t < taps; ++t)
              d[((size_t)t * L.cout + n) * L.cin + k] = dru_bf16(w[((size_t)n * L.cin + k) * taps + t]);
        break;
      }
      case DRU_UP: {  // ConvTranspose2d weight [cin][cout][2][2] -> [quadrant][cout][cin]
        uint16_t* d = reinterpret_cast<uint16_t*>(dst);
        for (int k = 0; k < L.cin; ++k)
          for (int n = 0; n < L.cout; ++n)
            for (int t = 0; t < 4; ++t)
              d[((size_t)t * L.cout + n) * L.cin + k] = dru_bf16(w[((size_t)k * L.cout + n) * 4 + t]);
        break;
      }
    }
  }
  PSGLA_CUDA_TRY(cudaMemcpyAsync(packed_dev, host.data(), host.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  PSGLA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));  // `host` dies at return
  return PSGLA_OK;
}

extern "C" size_t psgla_drunet_workspace_bytes(psgla_img_shape s) { return dru_workspace_bytes(s.B, s.H, s.W); }

extern "C" int psgla_drunet_denoise_post(const void* packed_dev, psgla_img_shape shape, const void* den_in_dev,
                                         void* workspace_dev, size_t workspace_bytes, const float* base_dev,
                                         const psgla_post_params* post, float* x_out_dev, float* sample_dev,
                                         float* mean_dev, float* mean2_dev, void* stream) {
  return psgla_drunet_denoise_post_next(packed_dev, shape, den_in_dev, workspace_dev, workspace_bytes, base_dev, post, x_out_dev,
                                        sample_dev, mean_dev, mean2_dev, nullptr, stream);
}

extern "C" int psgla_drunet_denoise_post_next(const void* packed_dev, psgla_img_shape shape, const void* den_in_dev,
                                              void* workspace_dev, size_t workspace_bytes, const float* base_dev,
                                              const psgla_post_params* post, float* x_out_dev, float* sample_dev,
                                              float* mean_dev, float* mean2_dev, const psgla_next_pre* next, void* stream) {
  PSGLA_REQUIRE(packed_dev && den_in_dev && workspace_dev && post && x_out_dev, "psgla_drunet_denoise_post: null pointer");
  {
    const int rcn = check_next_pre(next, shape);
    if (rcn) return rcn;
  }
  PSGLA_REQUIRE((mean_dev == nullptr) == (mean2_dev == nullptr), "mean and mean2 must be given together");
  PSGLA_REQUIRE(shape.B > 0 && shape.C == 3 && shape.H > 0 && shape.W > 0, "image shape must be [B>0][3][H>0][W>0]");
  PSGLA_REQUIRE(shape.H % 8 == 0 && shape.W % 8 == 0, "DRUNet needs H and W to be multiples of 8 (got %d x %d): crop or pad "
                "the image (the reference's deepinv pads internally by an unverifiable rule)", shape.H, shape.W);
  const int B = shape.B, H = shape.H, W = shape.W;
  if (workspace_bytes < dru_workspace_bytes(B, H, W))
    return set_error(PSGLA_E_WORKSPACE, "workspace of %zu bytes is smaller than the %zu needed", workspace_bytes,
                     dru_workspace_bytes(B, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const auto& plan = dru_plan();
  const uint8_t* packed = (const uint8_t*)packed_dev;
  DruBuffers bufs;
  {
    uint8_t* pp = (uint8_t*)(((uintptr_t)workspace_dev + 1023) & ~(uintptr_t)1023);
    for (int s = 0; s < 4; ++s)
      for (int j = 0; j < 4; ++j) {
        bufs.buf[s][j] = pp;
        pp += dru_scale_bytes(B, H, W, s);
      }
  }
  size_t li = 0;
  int rc = 0;
  auto wptr = [&](size_t l) { return packed + plan[l].w_off; };
  auto bias_of = [&](size_t l, int nout_pad, int cin_pad) {
    return reinterpret_cast<const float*>(packed + plan[l].w_off + (size_t)9 * nout_pad * cin_pad * 2);
  };
  // four residual blocks at scale s starting from `cur`; the last block also adds `skip` (U-Net skip connection)
  auto resblocks = [&](int s, uint8_t* cur, const uint8_t* skip) -> uint8_t* {
    const int C = 64 << s, h = H >> s, w = W >> s;
    for (int blk = 0; blk < 4 && !rc; ++blk) {
      uint8_t* scratch[2];
      int n = 0;
      for (int j = 1; j < 4 && n < 2; ++j)
        if (bufs.buf[s][j] != cur) scratch[n++] = bufs.buf[s][j];
      uint8_t* tmp = scratch[0];
      uint8_t* nxt = scratch[1];
      const uint8_t* extra = (blk == 3) ? skip : nullptr;
      if (s == 0) {
        rc = conv64_hidden(cur, tmp, wptr(li), bias_of(li, 64, 64), B, h, w, 1, nullptr, nullptr, st);
        if (!rc) rc = conv64_hidden(tmp, nxt, wptr(li + 1), bias_of(li + 1, 64, 64), B, h, w, 0, cur, extra, st);
      } else {
        rc = conv_gemm_layer(0, B, h, w, C, C, wptr(li), cur, nullptr, nullptr, tmp, 1, st);
        if (!rc) rc = conv_gemm_layer(0, B, h, w, C, C, wptr(li + 1), tmp, cur, extra, nxt, 0, st);
      }
      li += 2;
      cur = nxt;
    }
    return cur;
  };

  // head: x1
  rc = conv_first16(den_in_dev, bufs.buf[0][0], wptr(0), bias_of(0, 64, 16), B, H, W, 0, st);
  if (rc) return rc;
  li = 1;
  uint8_t* cur = bufs.buf[0][0];
  for (int s = 0; s < 3; ++s) {  // encoder
    cur = resblocks(s, cur, nullptr);
    if (rc) return rc;
    rc = conv_gemm_layer(1, B, H >> s, W >> s, 64 << s, 128 << s, wptr(li), cur, nullptr, nullptr, bufs.buf[s + 1][0], 0, st);
    if (rc) return rc;
    ++li;
    cur = bufs.buf[s + 1][0];
  }
  cur = resblocks(3, cur, bufs.buf[3][0]);  // body, + x4
  if (rc) return rc;
  for (int s = 2; s >= 0; --s) {  // decoder
    rc = conv_gemm_layer(2, B, H >> (s + 1), W >> (s + 1), 128 << s, 64 << s, wptr(li), cur, nullptr, nullptr, bufs.buf[s][1], 0, st);
    if (rc) return rc;
    ++li;
    cur = resblocks(s, bufs.buf[s][1], bufs.buf[s][0]);  // + x3 / x2 / x1
    if (rc) return rc;
  }
  // tail + fused Langevin post:  X+ = base_scale * base + gain * D
  return conv_last_post(cur, wptr(li), bias_of(li, 16, 64), B, H, W, base_dev, post->base_scale, post->gain, post->w_old,
                        post->w_new, x_out_dev, sample_dev, mean_dev, mean2_dev, next, st);
}
