// DnCNN conv3x3 layers as implicit GEMM on tcgen05 tensor cores (sm_100a), activations bf16 NHWC, fp32 accumulate.
//
// Replaces the cuDNN fp32 convolutions behind `denoiser.forward` (restoration_algorithms.py:238,
// sampling_images.py:156; architecture: deepinv.models.DnCNN, see oracle/image_oracle.py).
//
// Mapping (DESIGN.md "conv kernel"):
//   GEMM view per output image row segment:  D[128 pixels x NOUT] = sum over 9 taps  A_tap[128 x CIN] * W_tap[NOUT x CIN]^T
//   * A_tap is a *window* into input rows kept in shared memory: a work item walks down a 128-pixel-wide strip,
//     every input row (130 pixels = 128 + halo, CIN channels, 128B- or 32B-swizzled by TMA) is loaded ONCE into a
//     ring of NSTAGE slots and used by the 3 output rows around it; the horizontal tap offset dx is a +dx*row_bytes
//     shift of the UMMA shared-memory descriptor's start address, the vertical tap dy picks the ring slot.
//     TMA out-of-bounds zero fill provides the convolution's zero padding left/right; rows above/below the image
//     are simply skipped (their taps contribute zero).
//   * W (9 taps, bf16, K-major, pre-swizzled on the host) stays resident in shared memory for the CTA's lifetime.
//   * accumulators live in TMEM (2 stages x NOUT columns); 4 epilogue warps read them with tcgen05.ld while the
//     single MMA thread already works on the next row; one TMA thread keeps the ring full.
//   Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
//   Epilogues: bias + ReLU -> bf16 NHWC (hidden layers); or the fused Langevin "post" step for the last layer:
//   X+ = base + gain * (conv + bias), sample thinning and running E[X], E[X^2] (restoration_algorithms.py:238-262).
#include <cuda_bf16.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "sm100.cuh"

namespace psgla {

using namespace sm100;

constexpr int TILE_M = 128;
constexpr int BOX_W = TILE_M + 2;
constexpr int NSTAGE = 8;
constexpr int NACC = 2;
constexpr int CONV_THREADS = 192;

constexpr int round_up_c(int v, int a) { return (v + a - 1) / a * a; }

template <int CIN, int NOUT>
struct ConvCfg {
  static constexpr int ROW_BYTES = CIN * 2;
  static constexpr int BOX_BYTES = BOX_W * ROW_BYTES;
  static constexpr int SLOT_BYTES = round_up_c(BOX_BYTES, 1024);
  static constexpr uint32_t LAYOUT = (CIN == 64) ? LAYOUT_SW128 : LAYOUT_SW32;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int TAP_BYTES = NOUT * ROW_BYTES;
  static constexpr int W_BYTES = 9 * TAP_BYTES;
  static constexpr int OFF_RING = round_up_c(W_BYTES, 1024);
  static constexpr int OFF_BIAS = OFF_RING + NSTAGE * SLOT_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 256;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + slack to align the dynamic base to 1024
  static constexpr int TMEM_COLS = (NACC * NOUT) < 32 ? 32 : NACC * NOUT;
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two");
};

struct ConvParams {
  int B, H, W;
  int R, strips, row_blocks, n_items;
  const uint8_t* weights;  // 9 taps, swizzled
  const float* bias;       // NOUT floats
  __nv_bfloat16* out;      // EPI_HIDDEN
  int relu;
  int desc_mode;  // 0: base_offset = 0 (swizzle anchored at 1024 B); 1: base_offset = (start >> 7) & 7
  // EPI_POST
  const float* base;
  float* x_out;
  float* sample;
  float* mean;
  float* mean2;
  float gain, w_old, w_new;
};

enum { EPI_HIDDEN = 0, EPI_POST = 2 };

struct ItemCoord {
  int b, y0, rcur, x0, ylo, yhi;
};
__device__ __forceinline__ ItemCoord decode_item(const ConvParams& p, int item) {
  ItemCoord c;
  const int sx = item % p.strips;
  const int t = item / p.strips;
  const int ry = t % p.row_blocks;
  c.b = t / p.row_blocks;
  c.y0 = ry * p.R;
  c.rcur = min(p.R, p.H - c.y0);
  c.x0 = sx * TILE_M;
  c.ylo = max(c.y0 - 1, 0);
  c.yhi = min(c.y0 + c.rcur, p.H - 1);
  return c;
}

template <int CIN, int NOUT, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmap, const ConvParams p) {
  using Cfg = ConvCfg<CIN, NOUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* ring = smem + Cfg::OFF_RING;
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + NACC;
  uint64_t* wbar = tempty + NACC;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one elected arrive per epilogue warp
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NOUT) bias_s[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      tma_prefetch_desc(&tmap);
      mbar_expect_tx(wbar, Cfg::W_BYTES);
      bulk_load(smem_w, p.weights, Cfg::W_BYTES, wbar);
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
          const uint32_t slot = L % NSTAGE, use = L / NSTAGE;
          mbar_wait(&empty[slot], (use & 1) ^ 1);
          mbar_expect_tx(&full[slot], Cfg::BOX_BYTES);
          tma_load_4d(ring + slot * Cfg::SLOT_BYTES, &tmap, &full[slot], 0, c.x0 - 1, y, c.b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, NOUT);
      const uint32_t ring_addr = smem_u32(ring), w_addr = smem_u32(smem_w);
      mbar_wait(wbar, 0);
      tc_fence_after();
      uint32_t L0 = 0, T = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        int waited = 0;
        const int ylast = c.y0 + c.rcur - 1;
        for (int y = c.y0; y <= ylast; ++y, ++T) {
          const int need = min(y + 1, c.yhi) - c.ylo + 1;
          while (waited < need) {
            const uint32_t q = L0 + waited;
            mbar_wait(&full[q % NSTAGE], (q / NSTAGE) & 1);
            ++waited;
          }
          const uint32_t acc = T % NACC;
          mbar_wait(&tempty[acc], ((T / NACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * NOUT;
          uint32_t accumulate = 0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= p.H) continue;
            const uint32_t q = L0 + (uint32_t)(yy - c.ylo);
            const uint32_t a_row = ring_addr + (q % NSTAGE) * Cfg::SLOT_BYTES;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint32_t a0 = a_row + dx * Cfg::ROW_BYTES;
              const uint32_t b0 = w_addr + (dy * 3 + dx) * Cfg::TAP_BYTES;
              const uint32_t boff = p.desc_mode == 1 ? ((a0 >> 7) & 7) : 0;
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k) {
                const uint64_t adesc = make_smem_desc(a0 + k * 32, Cfg::SBO, Cfg::LAYOUT, boff);
                const uint64_t bdesc = make_smem_desc(b0 + k * 32, Cfg::SBO, Cfg::LAYOUT, 0);
                umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit(&tfull[acc]);
          if (y - 1 >= c.ylo) umma_commit(&empty[(L0 + (uint32_t)(y - 1 - c.ylo)) % NSTAGE]);
          if (y == ylast)
            for (int yy = y; yy <= c.yhi; ++yy) umma_commit(&empty[(L0 + (uint32_t)(yy - c.ylo)) % NSTAGE]);
        }
        L0 += (uint32_t)(c.yhi - c.ylo + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q4 = warp & 3;
    uint32_t T = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      const int x = c.x0 + q4 * 32 + lane;
      const bool valid = x < p.W;
      for (int y = c.y0; y < c.y0 + c.rcur; ++y, ++T) {
        const uint32_t acc = T % NACC;
        mbar_wait(&tfull[acc], (T / NACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * NOUT;
        if (EPI == EPI_HIDDEN) {
          __nv_bfloat16* outp = p.out + (((size_t)c.b * p.H + y) * p.W + x) * NOUT;
#pragma unroll
          for (int half = 0; half < NOUT / 32; ++half) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + half * 32, v);
            tmem_ld_wait();
            if (half == NOUT / 32 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
            }
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float f0 = __uint_as_float(v[2 * j]) + bias_s[half * 32 + 2 * j];
              float f1 = __uint_as_float(v[2 * j + 1]) + bias_s[half * 32 + 2 * j + 1];
              if (p.relu) {
                f0 = fmaxf(f0, 0.f);
                f1 = fmaxf(f1, 0.f);
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              packed[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(outp + half * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
          }
        } else {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          if (valid) {
            const size_t plane = (size_t)p.H * p.W;
            const size_t idx0 = ((size_t)c.b * 3) * plane + (size_t)y * p.W + x;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              const size_t idx = idx0 + ch * plane;
              const float r = __uint_as_float(v[ch]) + bias_s[ch];
              const float xn = p.base ? fmaf(p.gain, r, p.base[idx]) : r;
              p.x_out[idx] = xn;
              if (p.sample) p.sample[idx] = xn;
              if (p.mean) {
                // three rounded fp32 operations each, as the reference's eager ops (restoration_algorithms.py:257-258)
                p.mean[idx] = __fadd_rn(__fmul_rn(p.w_old, p.mean[idx]), __fmul_rn(p.w_new, xn));
                p.mean2[idx] = __fadd_rn(__fmul_rn(p.w_old, p.mean2[idx]), __fmul_rn(p.w_new, __fmul_rn(xn, xn)));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
PFN_tensorMapEncodeTiled get_tensor_map_encoder() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(ptr);
  });
  return fn;
}

static int make_act_tensor_map(CUtensorMap* map, const void* ptr, int B, int H, int W, int C) {
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return set_error(PSGLA_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)BOX_W, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = (C == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PSGLA_OK;
}

static int desc_mode_from_env() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("PSGLA_DESC_MODE");
    mode = e ? atoi(e) : 0;
  }
  return mode;
}

static void plan_items(ConvParams* p) {
  p->strips = (p->W + TILE_M - 1) / TILE_M;
  const int sms = num_sms();
  int best = 1;
  const int cands[] = {32, 16, 8, 4, 2, 1};
  for (int R : cands) {
    if (R > p->H && R != 1) continue;
    const long long items = (long long)p->B * p->strips * ((p->H + R - 1) / R);
    best = R;
    if (items >= 2LL * sms) break;
  }
  const char* e = getenv("PSGLA_CONV_ROWS");
  if (e && atoi(e) > 0) best = atoi(e);
  p->R = best;
  p->row_blocks = (p->H + best - 1) / best;
  p->n_items = p->B * p->strips * p->row_blocks;
}

template <int CIN, int NOUT, int EPI>
static int launch_conv(const void* in, ConvParams p, cudaStream_t st) {
  using Cfg = ConvCfg<CIN, NOUT>;
  CUtensorMap map;
  int rc = make_act_tensor_map(&map, in, p.B, p.H, p.W, CIN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(conv3x3_kernel<CIN, NOUT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  plan_items(&p);
  p.desc_mode = desc_mode_from_env();
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  conv3x3_kernel<CIN, NOUT, EPI><<<grid, CONV_THREADS, Cfg::SMEM_BYTES, st>>>(map, p);
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}

// ---- packed weight layout: per layer [weights (9 taps, swizzled) | bias fp32], each layer 1024 B aligned
struct LayerInfo {
  int cin, nout;  // padded
  size_t w_off, b_off;
};
static LayerInfo layer_info(int depth, int layer) {
  LayerInfo li{};
  size_t off = 0;
  for (int l = 0; l <= layer; ++l) {
    const int cin = (l == 0) ? 16 : 64;
    const int nout = (l == depth - 1) ? 16 : 64;
    li.cin = cin;
    li.nout = nout;
    li.w_off = off;
    const size_t wbytes = (size_t)9 * nout * cin * 2;
    li.b_off = off + wbytes;
    off = (li.b_off + (size_t)nout * 4 + 1023) / 1024 * 1024;
  }
  return li;
}
static size_t packed_bytes(int depth) {
  LayerInfo li = layer_info(depth, depth - 1);
  return (li.b_off + (size_t)li.nout * 4 + 1023) / 1024 * 1024;
}

static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace psgla

using namespace psgla;

extern "C" size_t psgla_dncnn_packed_bytes(int depth) { return depth >= 2 ? packed_bytes(depth) : 0; }

extern "C" int psgla_dncnn_pack_weights(int depth, const float* const* weights_host, const float* const* biases_host,
                                        void* packed_dev, void* stream) {
  PSGLA_REQUIRE(depth >= 2 && weights_host && packed_dev, "psgla_dncnn_pack_weights: bad argument");
  std::vector<uint8_t> host(packed_bytes(depth), 0);
  for (int l = 0; l < depth; ++l) {
    const LayerInfo li = layer_info(depth, l);
    const int cin_real = (l == 0) ? 3 : 64;
    const int nout_real = (l == depth - 1) ? 3 : 64;
    const int row_bytes = li.cin * 2;
    const float* w = weights_host[l];  // OIHW [nout_real][cin_real][3][3]
    PSGLA_REQUIRE(w != nullptr, "layer %d: null weight pointer", l);
    for (int tap = 0; tap < 9; ++tap)
      for (int n = 0; n < nout_real; ++n)
        for (int k = 0; k < cin_real; ++k) {
          const float val = w[((size_t)n * cin_real + k) * 9 + tap];
          const int kbyte = k * 2;
          int chunk = kbyte >> 4;
          chunk ^= (row_bytes == 128) ? (n & 7) : ((n >> 2) & 1);
          const size_t phys = li.w_off + (size_t)tap * li.nout * row_bytes + (size_t)n * row_bytes + chunk * 16 + (kbyte & 15);
          const uint16_t h = f32_to_bf16_rn(val);
          std::memcpy(&host[phys], &h, 2);
        }
    if (biases_host && biases_host[l])
      std::memcpy(&host[li.b_off], biases_host[l], (size_t)nout_real * 4);
  }
  PSGLA_CUDA_TRY(cudaMemcpyAsync(packed_dev, host.data(), host.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  PSGLA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));  // `host` dies at return
  return PSGLA_OK;
}

extern "C" size_t psgla_dncnn_workspace_bytes(psgla_img_shape s) {
  return 2 * ((size_t)s.B * s.H * s.W * 64 * 2 + 1024);
}

static int check_shape(const psgla_img_shape& s) {
  PSGLA_REQUIRE(s.B > 0 && s.H > 0 && s.W > 0 && s.C == 3, "image shape must be [B>0][3][H>0][W>0], got [%d][%d][%d][%d]",
                s.B, s.C, s.H, s.W);
  return PSGLA_OK;
}

static ConvParams base_params(const psgla_img_shape& s, const uint8_t* packed, const LayerInfo& li) {
  ConvParams p{};
  p.B = s.B;
  p.H = s.H;
  p.W = s.W;
  p.weights = packed + li.w_off;
  p.bias = reinterpret_cast<const float*>(packed + li.b_off);
  return p;
}

extern "C" int psgla_conv3x3_layer(const void* packed_dev, int depth, int layer, psgla_img_shape shape,
                                   const void* in_dev, void* out_dev, int relu, void* stream) {
  PSGLA_REQUIRE(packed_dev && in_dev && out_dev && depth >= 2 && layer >= 0 && layer < depth,
                "psgla_conv3x3_layer: bad argument");
  int rc = check_shape(shape);
  if (rc) return rc;
  const LayerInfo li = layer_info(depth, layer);
  ConvParams p = base_params(shape, (const uint8_t*)packed_dev, li);
  p.relu = relu;
  cudaStream_t st = (cudaStream_t)stream;
  if (layer == depth - 1) {  // raw conv + bias -> fp32 NCHW
    p.x_out = (float*)out_dev;
    return launch_conv<64, 16, EPI_POST>(in_dev, p, st);
  }
  p.out = (__nv_bfloat16*)out_dev;
  return layer == 0 ? launch_conv<16, 64, EPI_HIDDEN>(in_dev, p, st) : launch_conv<64, 64, EPI_HIDDEN>(in_dev, p, st);
}

extern "C" int psgla_dncnn_residual_post(int depth, const void* packed_dev, psgla_img_shape shape,
                                         const void* den_in_dev, void* workspace_dev, size_t workspace_bytes,
                                         const float* base_dev, const psgla_post_params* post, float* x_out_dev,
                                         float* sample_dev, float* mean_dev, float* mean2_dev, void* stream) {
  PSGLA_REQUIRE(packed_dev && den_in_dev && workspace_dev && post && x_out_dev && depth >= 2,
                "psgla_dncnn_residual_post: bad argument");
  PSGLA_REQUIRE((mean_dev == nullptr) == (mean2_dev == nullptr), "mean and mean2 must be given together");
  int rc = check_shape(shape);
  if (rc) return rc;
  if (workspace_bytes < psgla_dncnn_workspace_bytes(shape))
    return set_error(PSGLA_E_WORKSPACE, "workspace of %zu bytes is smaller than the %zu needed", workspace_bytes,
                     psgla_dncnn_workspace_bytes(shape));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t half = ((size_t)shape.B * shape.H * shape.W * 64 * 2 + 1023) / 1024 * 1024;
  uint8_t* ws[2] = {(uint8_t*)workspace_dev, (uint8_t*)workspace_dev + half};
  const uint8_t* packed = (const uint8_t*)packed_dev;
  const void* cur = den_in_dev;
  for (int l = 0; l < depth - 1; ++l) {
    const LayerInfo li = layer_info(depth, l);
    ConvParams p = base_params(shape, packed, li);
    p.relu = 1;
    p.out = (__nv_bfloat16*)ws[l & 1];
    rc = (l == 0) ? launch_conv<16, 64, EPI_HIDDEN>(cur, p, st) : launch_conv<64, 64, EPI_HIDDEN>(cur, p, st);
    if (rc) return rc;
    cur = ws[l & 1];
  }
  const LayerInfo li = layer_info(depth, depth - 1);
  ConvParams p = base_params(shape, packed, li);
  p.base = base_dev;
  p.x_out = x_out_dev;
  p.sample = sample_dev;
  p.mean = mean_dev;
  p.mean2 = mean2_dev;
  p.gain = post->gain;
  p.w_old = post->w_old;
  p.w_new = post->w_new;
  return launch_conv<64, 16, EPI_POST>(cur, p, st);
}

// ------------------------------------------------------------------------------------------------ descriptor self-test
namespace psgla {
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ d, int row_shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                // 136 rows x 128 B = 17408 B
  uint8_t* sb = smem + 18 * 1024;    // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], 136 * 128 + 64 * 128);
    tma_load_2d(sa, &map_a, &bar[0], 0, 0);
    tma_load_2d(sb, &map_b, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sa) + row_shift * 128, b0 = smem_u32(sb);
    const uint32_t boff = mode == 1 ? ((a0 >> 7) & 7) : 0;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k)
      umma_bf16(tbase, make_smem_desc(a0 + k * 32, 1024, LAYOUT_SW128, boff),
                make_smem_desc(b0 + k * 32, 1024, LAYOUT_SW128, 0), idesc, k > 0);
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tbase + ((uint32_t)(warp * 32) << 16) + half * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[(size_t)(warp * 32 + lane) * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 64);
  }
}
}  // namespace psgla

extern "C" int psgla_selftest_umma(const void* a_dev, const void* b_dev, float* d_dev, int row_shift, int mode,
                                   void* stream) {
  PSGLA_REQUIRE(a_dev && b_dev && d_dev && row_shift >= 0 && row_shift <= 8, "psgla_selftest_umma: bad argument");
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return set_error(PSGLA_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  CUtensorMap ma, mb;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {64, 136};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 136};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_dev), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "tensor map A: CUresult %d", (int)r);
  }
  {
    const cuuint64_t dims[2] = {64, 64};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 64};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_dev), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "tensor map B: CUresult %d", (int)r);
  }
  const int smem = 30 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, d_dev, row_shift, mode);
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}
