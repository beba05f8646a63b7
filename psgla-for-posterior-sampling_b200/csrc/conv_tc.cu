// 3x3 convolution layers with shared-memory-resident weights as implicit GEMM on tcgen05 tensor cores (sm_100a),
// activations bf16 NHWC, fp32 accumulate: all of DnCNN and the 64-channel full-resolution layers of DRUNet.
//
// Replaces the cuDNN fp32 convolutions behind `denoiser.forward` (restoration_algorithms.py:238,
// sampling_images.py:156; architecture: deepinv.models.DnCNN / DRUNet, see oracle/image_oracle.py).
//
// GEMM view per 128-pixel output row segment:  D[128 x NOUT] = sum over 9 taps  A_tap[128 x CIN] * W_tap[NOUT x CIN]^T.
// A work item is a (chain, 128-pixel strip, block of R rows); every input row (130 pixels = 128 + halo, TMA out-of-bounds
// zero fill = the convolution's padding) is loaded ONCE and serves the three output rows around it.  Kernels in this file:
//   conv3x3_kernel<CIN,NOUT,EPI>   SS form: A_tap = window into the shared-memory row ring (tap dx = descriptor start + dx rows,
//                                  tap dy = ring slot).  Used for the 3(16)-channel first layer; A/B variant for 64 channels.
//   conv3x3_ts_kernel<NOUT,EPI>    TS form: four loader warps copy each row into TENSOR MEMORY three times (shifted by dx) and
//                                  the MMAs take A from there (N/2 instead of 32 + N/4 cycles per MMA at N = 64).  Last layer
//                                  (NOUT = 16, fused Langevin post / next pre epilogue); single-CTA form of the hidden layers.
//   conv3x3_ts2_kernel<NOUT,RES>   the hidden layers: a CTA PAIR (cta_group::2, M = 256) on two adjacent strips, weights split
//                                  32 / 32 output channels between the two shared memories; RES = residual input by TMA.
//   conv3x3_ts_chain_kernel        experiment: 18 layers in one persistent launch with a grid barrier (off by default).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), [warps 2..5 = loaders (TS forms)], 8 epilogue warps
// in two groups (TMEM -> bias / residual / ReLU -> bf16 -> swizzled staging box -> TMA store; or the fused Langevin step,
// restoration_algorithms.py:238-262, for the last layer).  Consecutive layers walk their items in opposite directions
// (ConvParams::reverse) so that each starts on what the previous one left in L2.  DESIGN.md section 4 has the measurements.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "conv_api.cuh"
#include "langevin.cuh"
#include "sm100.cuh"

namespace psgla {

using namespace sm100;

constexpr int TILE_M = 128;
constexpr int BOX_W = TILE_M + 2;
constexpr int NSTAGE_64 = 6;   // input-row ring slots, 64-channel rows (17 KB each); 16-channel rows: ConvCfg::NSTAGE
constexpr int NACC = 4;       // TMEM accumulator stages; stage s is drained by epilogue group s & 1
constexpr int EPI_WARPS = 8;  // two groups of four warps (one warp per TMEM lane quarter)
constexpr int CONV_THREADS = 64 + 32 * EPI_WARPS;

constexpr int round_up_c(int v, int a) { return (v + a - 1) / a * a; }

enum { EPI_HIDDEN = 0, EPI_POST = 2 };

// f[0..7] += eight bf16 values packed in a 16-byte vector
__device__ __forceinline__ void add_bf16x8(float (&f)[8], const uint4 r) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] += __uint_as_float(w[i] << 16);
    f[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int CIN, int NOUT, int EPI>
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
  // 32-pixel output boxes per epilogue warp: two (the warp fills one while the TMA store of the previous row still reads
  // the other) where shared memory allows, i.e. not next to 72 KB of weights and a 64-channel ring
  // Ring depth.  A slot stays occupied for three output rows, so NSTAGE - 3 rows are in flight ahead of the MMAs; the
  // 16-channel first layer consumes a row in ~430 cycles against ~2 us of TMA latency from HBM and needs a deep ring
  // (6 slots: 1 560 cycles per row measured), its rows are only 5 KB.
  static constexpr int NSTAGE = (CIN == 16) ? 20 : NSTAGE_64;
  static constexpr int STAGE_BUFS = (CIN == 16) ? 2 : 1;
  static constexpr int STAGE_BYTES = (EPI == EPI_HIDDEN) ? STAGE_BUFS * 32 * NOUT * 2 : 0;
  static constexpr int OFF_STAGE = OFF_RING + NSTAGE * SLOT_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGE + EPI_WARPS * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 256;
  static constexpr int BAR_BYTES = 512;
  static_assert((2 * NSTAGE + 2 * NACC + 1) * 8 + 4 <= BAR_BYTES, "barrier block overflows");
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;  // + slack to align the dynamic base to 1024
  static constexpr int TMEM_COLS = (NACC * NOUT) < 32 ? 32 : NACC * NOUT;
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of one CTA");
};

struct ConvParams {
  int B, H, W;
  int R, strips, row_blocks, n_items;
  int reverse;  // walk the work items back to front (see decode_item)
  const uint8_t* weights;  // 9 taps, swizzled
  const float* bias;       // NOUT floats
  int relu;
  // EPI_HIDDEN: optional bf16 NHWC tensors of the output's shape added before the ReLU (DRUNet residual / skip adds)
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  // EPI_POST
  const float* base;
  float* x_out;
  float* sample;
  float* mean;
  float* mean2;
  float gain, base_scale, w_old, w_new;
  // EPI_POST, optional: the "pre" step of the NEXT iteration applied to the iterate this epilogue produces (inpainting):
  // nx_base = langevin_base(X+), nx_den_in = bf16 NHWC16 of it (PSGLA) or of X+ (PnP-ULA); nx_base may alias base.
  int nx_enable;
  PreArgs nx;
  const float* nx_mask;
  const float* nx_y;
  int nx_mask_B, nx_y_B;
  float* nx_base;
  __nv_bfloat16* nx_den_in;
};

static int set_next_pre(ConvParams* p, const psgla_next_pre* next);  // host: fills the nx_* fields (defined with the API)

struct ItemCoord {
  int b, y0, rcur, x0, ylo, yhi;
};
// Items are dealt to the persistent CTAs in index order, so a layer finishes with the END of the activation tensor freshly
// written -- and a 268 MB tensor (32 chains of 256 x 256 x 64 bf16) leaves roughly its last third in the 126 MB L2.  Consecutive
// layers therefore walk the items in opposite directions (reverse = layer parity): each layer starts on what the previous one
// wrote last and reads it from L2 instead of HBM.  (In the pair kernel items 2i and 2i + 1 swap ranks under the reversal and
// still share chain and rows.)
__device__ __forceinline__ ItemCoord decode_item(const ConvParams& p, int item) {
  if (p.reverse) item = p.n_items - 1 - item;
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

// relu(a), relu(b) (or a, b) rounded to nearest-even bf16 and packed {lo = a, hi = b}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, bool relu) {
  uint32_t d;
  if (relu)
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// ------------------------------------------------------------------------------------------------ epilogues
// Eight epilogue warps form two groups of four (one warp per TMEM lane quarter q4); group g drains the output rows
// with T % 2 == g, where T counts the CTA's output rows in issue order and accumulator stage = T % NACC_.

// Hidden layers: TMEM -> +bias -> ReLU -> bf16 -> 128B-swizzled staging box in shared memory -> one TMA store of
// 32 pixels x NOUT channels per warp and row (clipped at the image edge by the tensor map).
template <int NOUT, int NACC_>
__device__ __forceinline__ void epilogue_hidden(const ConvParams& p, const CUtensorMap* tmap_out, uint8_t* stage,
                                                const float* bias_s, uint64_t* tfull, uint64_t* tempty,
                                                uint32_t tmem_base, int grp, int q4, int lane, uint32_t& T,
                                                uint32_t tempty_cluster = 0, int stage_bufs = 1) {
  // tempty_cluster != 0 (CTA-pair kernel): the accumulator-free barriers live in the leader CTA, at this cluster address.
  // T: the CTA's running output-row counter (accumulator stage and mbarrier phase); it carries over when one kernel
  // runs several layers back to back.  bias_s may point to shared or global memory.
  // stage_bufs == 2: the warp alternates between two staging boxes, so a row is staged while the previous row's TMA
  // store is still reading its box (the store's read latency otherwise serialises with the warp's work on every row)
  uint32_t nrow = 0;
  const float4* bias4 = reinterpret_cast<const float4*>(bias_s);
  const bool relu = p.relu != 0;
  if (lane == 0) tma_prefetch_desc(tmap_out);
  // the residual tensors come from earlier kernels and are now read ahead of the accumulator (i.e. before anything in this
  // warp depends on the producer's own griddepcontrol.wait)
  if (p.res1 != nullptr) griddep_wait();
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const ItemCoord c = decode_item(p, item);
    const int xw = c.x0 + q4 * 32;  // first pixel of this warp's 32-pixel box
    for (int y = c.y0; y < c.y0 + c.rcur; ++y, ++T) {
      if ((int)(T & 1) != grp) continue;
      const uint32_t acc = T % NACC_;
      // residual inputs (DRUNet) are fetched while the row's MMAs are still in flight
      const bool has_res = p.res1 != nullptr && xw + lane < p.W;
      const size_t roff = (((size_t)c.b * p.H + y) * p.W + (xw + lane)) * NOUT;
      uint4 rr[NOUT / 8];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < NOUT / 8; ++j) rr[j] = *reinterpret_cast<const uint4*>(p.res1 + roff + 8 * j);
      }
      mbar_wait(&tfull[acc], (T / NACC_) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * NOUT;
      uint32_t v[NOUT];
#pragma unroll
      for (int h = 0; h < NOUT / 32; ++h) tmem_ld_32x32b_x32(taddr + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[h * 32]));
      tmem_ld_wait();
      tc_fence_before();
      // the staging box of the previous row must have been read by its TMA store before it is overwritten
      uint8_t* stage_cur = stage + (stage_bufs == 2 ? (nrow & 1u) * (32 * NOUT * 2) : 0);
      const uint32_t stage_row = smem_u32(stage_cur) + lane * (NOUT * 2);
      ++nrow;
      if (lane == 0) {
        if (stage_bufs == 2)
          bulk_wait_group_read1();
        else
          bulk_wait_group_read0();
        if (tempty_cluster)
          mbar_arrive_remote(tempty_cluster + acc * 8u);
        else
          mbar_arrive(&tempty[acc]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NOUT / 8; ++j) {  // 16-byte chunk j = channels 8j..8j+7
        const float4 b0 = bias4[2 * j], b1 = bias4[2 * j + 1];
        float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                      __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                      __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                      __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
        if (has_res) {
          add_bf16x8(f, rr[j]);
          if (p.res2) add_bf16x8(f, *reinterpret_cast<const uint4*>(p.res2 + roff + 8 * j));
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1], relu);
        o.y = pack_bf16x2(f[2], f[3], relu);
        o.z = pack_bf16x2(f[4], f[5], relu);
        o.w = pack_bf16x2(f[6], f[7], relu);
        st_shared_v4(stage_row + ((uint32_t)(j ^ (lane & 7)) << 4), o);  // 128B swizzle: chunk ^= row & 7
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (xw < p.W) tma_store_4d(tmap_out, stage_cur, 0, xw, y, c.b);
        bulk_commit_group();  // also when nothing was stored: wait_group.read 1 counts one group per row
      }
    }
  }
  if (lane == 0) bulk_wait_group0();
}

// The same with the first residual tensor fetched by TMA (pair kernel, DRUNet's 64-channel residual blocks).  At 64 chains of
// 320 x 480 the layer moves 3.8 GB and is HBM-bound (580 us at the measured copy bandwidth against 540 us of MMAs); per-thread
// 16-byte loads of the residual reached only ~4 TB/s in total (950 us per layer, ncu).  Here lane 0 loads the warp's
// 32-pixel residual box of its NEXT row straight into the staging box that row will use (the two boxes alternate), one row
// ahead; the lanes then read-modify-write the box in shared memory and the TMA store ships it.  Box ownership: the load for
// row n + 1 into box b' is issued after cp.async.bulk.wait_group.read 1, i.e. once the store of row n - 1 has finished
// reading b'; generic writes precede the store by fence.proxy.async as before.
template <int NOUT, int NACC_>
__device__ __forceinline__ void epilogue_hidden_tmares(const ConvParams& p, const CUtensorMap* tmap_out,
                                                       const CUtensorMap* tmap_res, uint8_t* stage, uint64_t* rbar,
                                                       const float* bias_s, uint64_t* tfull, uint64_t* tempty,
                                                       uint32_t tmem_base, int grp, int q4, int lane, uint32_t tempty_cluster) {
  constexpr uint32_t BOX = 32 * NOUT * 2;
  struct Iter {
    int item, y, yend;
    uint32_t T;
    ItemCoord c;
    bool done;
  };
  auto step = [&](Iter& r) {
    ++r.y;
    ++r.T;
    if (r.y >= r.yend) {
      r.item += gridDim.x;
      if (r.item >= p.n_items) {
        r.done = true;
      } else {
        r.c = decode_item(p, r.item);
        r.y = r.c.y0;
        r.yend = r.c.y0 + r.c.rcur;
      }
    }
  };
  auto settle = [&](Iter& r) {
    while (!r.done && (int)(r.T & 1) != grp) step(r);
  };
  auto issue = [&](const Iter& r, uint32_t box) {  // lane 0: residual box of row r -> staging box `box`
    const int xw = r.c.x0 + q4 * 32;
    if (xw < p.W) {
      mbar_expect_tx(&rbar[box], BOX);
      tma_load_4d(stage + box * BOX, tmap_res, &rbar[box], 0, xw, r.y, r.c.b);
    }
  };
  const float4* bias4 = reinterpret_cast<const float4*>(bias_s);
  const bool relu = p.relu != 0;
  if (lane == 0) {
    tma_prefetch_desc(tmap_out);
    tma_prefetch_desc(tmap_res);
  }
  griddep_wait();  // the residual tensor was written by an earlier kernel
  Iter it;
  it.item = blockIdx.x;
  it.T = 0;
  it.done = it.item >= p.n_items;
  if (!it.done) {
    it.c = decode_item(p, it.item);
    it.y = it.c.y0;
    it.yend = it.c.y0 + it.c.rcur;
  }
  settle(it);
  uint32_t n = 0, cnt[2] = {0, 0};
  if (!it.done && lane == 0) issue(it, 0);
  while (!it.done) {
    Iter nxt = it;
    step(nxt);
    settle(nxt);
    const ItemCoord& c = it.c;
    const int y = it.y;
    const uint32_t T = it.T;
    const uint32_t box = n & 1u;
    const int xw = c.x0 + q4 * 32;
    const bool row_has = xw < p.W;
    const uint32_t acc = T % NACC_;
    // second residual tensor (U-Net skip, one layer per scale): per-thread loads, issued before the accumulator is awaited
    const bool has_res2 = p.res2 != nullptr && xw + lane < p.W;
    const size_t roff = (((size_t)c.b * p.H + y) * p.W + (xw + lane)) * NOUT;
    uint4 rr2[NOUT / 8];
    if (has_res2) {
#pragma unroll
      for (int j = 0; j < NOUT / 8; ++j) rr2[j] = *reinterpret_cast<const uint4*>(p.res2 + roff + 8 * j);
    }
    mbar_wait(&tfull[acc], (T / NACC_) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * NOUT;
    uint32_t v[NOUT];
#pragma unroll
    for (int h = 0; h < NOUT / 32; ++h) tmem_ld_32x32b_x32(taddr + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[h * 32]));
    tmem_ld_wait();
    tc_fence_before();
    if (lane == 0) {
      if (tempty_cluster)
        mbar_arrive_remote(tempty_cluster + acc * 8u);
      else
        mbar_arrive(&tempty[acc]);
    }
    if (row_has) {
      mbar_wait(&rbar[box], cnt[box] & 1u);  // this row's residual box has landed (pixels beyond W: zero fill)
      ++cnt[box];
    }
    uint8_t* stage_cur = stage + box * BOX;
    const uint32_t stage_row = smem_u32(stage_cur) + lane * (NOUT * 2);
#pragma unroll
    for (int j = 0; j < NOUT / 8; ++j) {
      const float4 b0 = bias4[2 * j], b1 = bias4[2 * j + 1];
      float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                    __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                    __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                    __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
      const uint32_t saddr = stage_row + ((uint32_t)(j ^ (lane & 7)) << 4);  // 128B swizzle: chunk ^= row & 7
      if (row_has) add_bf16x8(f, ld_shared_v4(saddr));
      if (has_res2) add_bf16x8(f, rr2[j]);
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1], relu);
      o.y = pack_bf16x2(f[2], f[3], relu);
      o.z = pack_bf16x2(f[4], f[5], relu);
      o.w = pack_bf16x2(f[6], f[7], relu);
      st_shared_v4(saddr, o);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (row_has) tma_store_4d(tmap_out, stage_cur, 0, xw, y, c.b);
      bulk_commit_group();
      if (!nxt.done) {
        bulk_wait_group_read1();  // the store of the previous row no longer reads the other box
        issue(nxt, box ^ 1u);
      }
    }
    __syncwarp();  // no lane touches the other box before lane 0 has seen it released
    ++n;
    it = nxt;
  }
  if (lane == 0) bulk_wait_group0();
}

// Last layer: the fused Langevin "post" step, fp32 NCHW (restoration_algorithms.py:238-262 / :115-135), optionally followed
// by the next iteration's "pre" on the fresh iterate.
// The layer is HBM-bound (128 B of activations in, ~70-140 B of fp32 state in and out per pixel) and its MMAs take only a few
// hundred cycles per row, so nothing hides a DRAM round trip behind them: the epilogue therefore walks ITS rows with a
// one-row-ahead register prefetch of everything it reads from global memory, and draws the row's noise before it waits
// for the accumulator.
struct PostRowIter {
  int item, y, yend;
  uint32_t T;
  ItemCoord c;
  bool done;
};
struct PostRowData {
  float bse[3], m1[3], m2[3], nmask[3], nobs[3];
};

template <int NOUT, int NACC_>
__device__ __forceinline__ void epilogue_post(const ConvParams& p, const float* bias_s, uint64_t* tfull, uint64_t* tempty,
                                              uint32_t tmem_base, int grp, int q4, int lane) {
  griddep_wait();  // base / running moments were written by earlier kernels
  const size_t plane = (size_t)p.H * p.W;
  auto step = [&](PostRowIter& r) {
    ++r.y;
    ++r.T;
    if (r.y >= r.yend) {
      r.item += gridDim.x;
      if (r.item >= p.n_items) {
        r.done = true;
      } else {
        r.c = decode_item(p, r.item);
        r.y = r.c.y0;
        r.yend = r.c.y0 + r.c.rcur;
      }
    }
  };
  auto settle = [&](PostRowIter& r) {  // forward to the next row this epilogue group drains
    while (!r.done && (int)(r.T & 1) != grp) step(r);
  };
  auto fetch = [&](const PostRowIter& r, PostRowData& d) {
    const int x = r.c.x0 + q4 * 32 + lane;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) d.bse[ch] = d.m1[ch] = d.m2[ch] = d.nmask[ch] = d.nobs[ch] = 0.f;
    if (r.done || x >= p.W) return;
    const size_t e0 = (size_t)r.y * p.W + x;
    const size_t idx0 = ((size_t)r.c.b * 3) * plane + e0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (p.base) d.bse[ch] = p.base[idx0 + ch * plane];
      if (p.mean) {
        d.m1[ch] = p.mean[idx0 + ch * plane];
        d.m2[ch] = p.mean2[idx0 + ch * plane];
      }
    }
    if (p.nx_enable) {
      const size_t mi = ((size_t)(p.nx_mask_B > 1 ? r.c.b : 0) * 3) * plane + e0;
      const size_t yi = ((size_t)(p.nx_y_B > 1 ? r.c.b : 0) * 3) * plane + e0;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        d.nmask[ch] = p.nx_mask[mi + ch * plane];
        d.nobs[ch] = p.nx_y[yi + ch * plane];
      }
    }
  };

  PostRowIter it;
  it.item = blockIdx.x;
  it.T = 0;
  it.done = it.item >= p.n_items;
  if (!it.done) {
    it.c = decode_item(p, it.item);
    it.y = it.c.y0;
    it.yend = it.c.y0 + it.c.rcur;
  }
  settle(it);
  PostRowData cur;
  fetch(it, cur);
  while (!it.done) {
    PostRowIter nxt = it;
    step(nxt);
    settle(nxt);
    PostRowData nd;
    fetch(nxt, nd);  // in flight while this row is processed
    const ItemCoord& c = it.c;
    const int y = it.y;
    const uint32_t T = it.T;
    const int x = c.x0 + q4 * 32 + lane;
    const bool valid = x < p.W;
    const size_t idx0 = ((size_t)c.b * 3) * plane + (size_t)y * p.W + x;
    float z[3] = {0.f, 0.f, 0.f};
    if (p.nx_enable) {
      if (p.nx.noise_mode == PSGLA_NOISE_PHILOX && (p.W & 3) == 0) {
        // Library stream: one Philox call serves four consecutive elements, and lanes 4k .. 4k+3 hold four consecutive
        // pixels (x0, the warp offset and W are multiples of 4).  Lane 4k + ch draws channel ch's quad, the four lanes
        // exchange components by shuffle: one Philox call per lane instead of three.
        const int sub = lane & 3;
        float z4[4] = {0.f, 0.f, 0.f, 0.f};
        if (sub < 3 && x - sub < p.W)
          draw_quad(p.nx, c.b, (uint32_t)((size_t)sub * plane + (size_t)y * p.W + (size_t)(x - sub)), z4);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int src = (lane & ~3) + ch;
          const float t0 = __shfl_sync(0xffffffffu, z4[0], src), t1 = __shfl_sync(0xffffffffu, z4[1], src);
          const float t2 = __shfl_sync(0xffffffffu, z4[2], src), t3 = __shfl_sync(0xffffffffu, z4[3], src);
          z[ch] = sub == 0 ? t0 : (sub == 1 ? t1 : (sub == 2 ? t2 : t3));
        }
      } else if (valid) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) z[ch] = draw_at(p.nx, c.b, (uint32_t)((size_t)ch * plane + (size_t)y * p.W + x));
      }
    }
    const uint32_t acc = T % NACC_;
    mbar_wait(&tfull[acc], (T / NACC_) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * NOUT;
    uint32_t v[16];
    tmem_ld_32x32b_x16(taddr, v);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[acc]);
    if (valid) {
      float xnew[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const size_t idx = idx0 + ch * plane;
        const float r = __uint_as_float(v[ch]) + bias_s[ch];
        const float xn = p.base ? fmaf(p.gain, r, p.base_scale * cur.bse[ch]) : r;
        xnew[ch] = xn;
        p.x_out[idx] = xn;
        if (p.sample) p.sample[idx] = xn;
        if (p.mean) {
          // three rounded fp32 operations each, as the reference's eager ops (restoration_algorithms.py:257-258)
          p.mean[idx] = __fadd_rn(__fmul_rn(p.w_old, cur.m1[ch]), __fmul_rn(p.w_new, xn));
          p.mean2[idx] = __fadd_rn(__fmul_rn(p.w_old, cur.m2[ch]), __fmul_rn(p.w_new, __fmul_rn(xn, xn)));
        }
      }
      if (p.nx_enable) {
        // the next iteration's Langevin "pre" on the fresh iterate: same arithmetic and the same noise element as
        // pre_inpaint_kernel (img_elementwise.cu), so fused and unfused runs agree bit for bit
        float din[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float bv = langevin_base(p.nx, xnew[ch], cur.nmask[ch] * (xnew[ch] - cur.nobs[ch]), z[ch]);
          p.nx_base[idx0 + ch * plane] = bv;
          din[ch] = (p.nx.alg == PSGLA_ALG_PNPULA) ? xnew[ch] : bv;
        }
        store_nhwc16(p.nx_den_in + (((size_t)c.b * plane) + (size_t)y * p.W + x) * 16, din[0], din[1], din[2], p.nx.den_in_c3);
      }
    }
    it = nxt;
    cur = nd;
  }
}

// ------------------------------------------------------------------------------------------------ SS kernel (A from smem)
template <int CIN, int NOUT, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const ConvParams p) {
  using Cfg = ConvCfg<CIN, NOUT, EPI>;
  constexpr int NSTAGE = Cfg::NSTAGE;
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

  // Programmatic dependent launch: let the next layer's CTAs start their prologue as ours retire; everything below that
  // touches memory written by an earlier kernel sits behind griddep_wait().
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one elected arrive per warp of the group that drains this stage
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
      bulk_load(smem_w, p.weights, Cfg::W_BYTES, wbar);  // weights are constant across launches: no dependency wait
      griddep_wait();
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
    // ---------------------------------------------------------------- MMA issuer
    // The whole warp walks the (warp-uniform) control flow so that barrier addresses and descriptors live in uniform
    // registers; one elected lane issues the tcgen05 instructions.  Descriptors differ only in their 14-bit start
    // address field, so each MMA costs two integer adds.
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, NOUT);
    constexpr uint32_t DESC_HI = (Cfg::SBO >> 4) | (1u << 14) | (Cfg::LAYOUT << 29);  // SBO | version 1 | swizzle mode
    const uint32_t ring_lo = (smem_u32(ring) >> 4) | 0x10000u;                         // start address >> 4 | LBO = 1
    const uint32_t w_lo = (smem_u32(smem_w) >> 4) | 0x10000u;
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
        if (elect_one()) {
          uint32_t accumulate = 0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= p.H) continue;
            const uint32_t q = L0 + (uint32_t)(yy - c.ylo);
            const uint32_t a_lo = ring_lo + (q % NSTAGE) * (uint32_t)(Cfg::SLOT_BYTES >> 4);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              // tap (dy, dx): the A operand is the input row shifted by dx pixels.  The swizzle pattern is anchored at
              // absolute 1024-byte boundaries, so shifting the start address by whole 128-byte rows needs no
              // base-offset correction (verified by psgla_selftest_umma).
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k) {
                const uint32_t al = a_lo + (uint32_t)((dx * Cfg::ROW_BYTES + k * 32) >> 4);
                const uint32_t bl = w_lo + (uint32_t)(((dy * 3 + dx) * Cfg::TAP_BYTES + k * 32) >> 4);
                umma_bf16(d_tmem, ((uint64_t)DESC_HI << 32) | al, ((uint64_t)DESC_HI << 32) | bl, idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit(&tfull[acc]);
          if (y - 1 >= c.ylo) umma_commit(&empty[(L0 + (uint32_t)(y - 1 - c.ylo)) % NSTAGE]);
          if (y == ylast)
            for (int yy = y; yy <= c.yhi; ++yy) umma_commit(&empty[(L0 + (uint32_t)(yy - c.ylo)) % NSTAGE]);
        }
        __syncwarp();
      }
      L0 += (uint32_t)(c.yhi - c.ylo + 1);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps
    const int ew = warp - 2;
    uint32_t T = 0;
    if (EPI == EPI_HIDDEN)
      epilogue_hidden<NOUT, NACC>(p, &tmap_out, smem + Cfg::OFF_STAGE + ew * Cfg::STAGE_BYTES, bias_s, tfull, tempty,
                                  tmem_base, ew >> 2, warp & 3, lane, T, 0, Cfg::STAGE_BUFS);
    else
      epilogue_post<NOUT, NACC>(p, bias_s, tfull, tempty, tmem_base, ew >> 2, warp & 3, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ TS kernel (A from TMEM)
// Measured on B200 (psgla_selftest_mma_rate): an M128 x N64 x K16 bf16 MMA takes 72 cycles with both operands in shared
// memory (the 4 KB A fetch is exposed) but 41 cycles with A in tensor memory (floor 32).  For the 64-input-channel layers
// four "loader" warps therefore copy every input row from the TMA ring into TMEM three times, shifted by dx = 0, 1, 2
// pixels (TMEM lanes are pixels and cannot be shifted by the MMA), and the MMAs read A from there:
//   TMEM columns [0, NACC_TS * NOUT)            accumulators (one stage per epilogue group)
//                [128 + s*96 + dx*32 + k*8 ...)  A ring: slot s = input row mod 4, shift dx, K-step k (8 columns = 16 bf16)
// Warps: 0 TMA producer, 1 MMA issuer, 2-5 loaders (TMEM lane quarter = warp & 3), 6-13 epilogue (two groups).
constexpr int TS_NSTAGE = 4;   // shared-memory staging slots of the TMA ring
constexpr int TS_NA = 4;       // input rows resident in TMEM
constexpr int TS_NACC = 2;
constexpr int TS_A_COL0 = 128;
constexpr int TS_THREADS = 64 + 128 + 32 * EPI_WARPS;

template <int NOUT, int EPI>
struct ConvTsCfg {
  static constexpr int ROW_BYTES = 128;
  static constexpr int BOX_BYTES = BOX_W * ROW_BYTES;
  static constexpr int SLOT_BYTES = round_up_c(BOX_BYTES, 1024);
  static constexpr int TAP_BYTES = NOUT * ROW_BYTES;
  static constexpr int W_BYTES = 9 * TAP_BYTES;
  static constexpr int OFF_RING = round_up_c(W_BYTES, 1024);
  static constexpr int STAGE_BUFS = 2;
  static constexpr int STAGE_BYTES = (EPI == EPI_HIDDEN) ? STAGE_BUFS * 32 * NOUT * 2 : 0;
  // staging ring depth: the last layer (N = 16: 9 x 4 MMAs of ~9 cycles per row, no output staging) outruns a 4-slot
  // ring by far and has the shared memory for a deep one
  static constexpr int NSTAGE = (NOUT == 16) ? 10 : TS_NSTAGE;
  static constexpr int OFF_STAGE = OFF_RING + NSTAGE * SLOT_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGE + EPI_WARPS * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 256;
  static constexpr int BAR_BYTES = 512;
  static_assert((2 * NSTAGE + 2 * TS_NA + 2 * TS_NACC + 3) * 8 + 4 <= BAR_BYTES, "barrier block overflows");
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
  static_assert(TS_NACC * NOUT <= TS_A_COL0 && TS_A_COL0 + TS_NA * 96 <= 512, "TMEM plan does not fit 512 columns");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of one CTA");
};

template <int NOUT, int EPI>
__global__ void __launch_bounds__(TS_THREADS, 1)
conv3x3_ts_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const ConvParams p) {
  using Cfg = ConvTsCfg<NOUT, EPI>;
  constexpr int NST = Cfg::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* ring = smem + Cfg::OFF_RING;
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);  // TMA landed a row in the staging ring
  uint64_t* empty = full + NST;                                 // loaders have copied it out
  uint64_t* afull = empty + NST;                                // the row's three shifted copies are in TMEM
  uint64_t* aempty = afull + TS_NA;                                   // every MMA reading them has completed
  uint64_t* tfull = aempty + TS_NA;
  uint64_t* tempty = tfull + TS_NACC;
  uint64_t* wbar = tempty + TS_NACC;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 4);
    }
    for (int i = 0; i < TS_NA; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < TS_NACC; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, 512);
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
      griddep_wait();
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
          const uint32_t slot = L % NST;
          mbar_wait(&empty[slot], ((L / NST) & 1) ^ 1);
          mbar_expect_tx(&full[slot], Cfg::BOX_BYTES);
          tma_load_4d(ring + slot * Cfg::SLOT_BYTES, &tmap, &full[slot], 0, c.x0 - 1, y, c.b);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform, one elected lane)
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, NOUT);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
    const uint32_t w_lo = (smem_u32(smem_w) >> 4) | 0x10000u;
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
          mbar_wait(&afull[q % TS_NA], (q / TS_NA) & 1);
          ++waited;
        }
        const uint32_t acc = T % TS_NACC;
        mbar_wait(&tempty[acc], ((T / TS_NACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NOUT;
        if (elect_one()) {
          uint32_t accumulate = 0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = y + dy - 1;
            if (yy < 0 || yy >= p.H) continue;
            const uint32_t q = L0 + (uint32_t)(yy - c.ylo);
            const uint32_t a_t = tmem_base + TS_A_COL0 + (q % TS_NA) * 96u;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t bl = w_lo + (uint32_t)(((dy * 3 + dx) * Cfg::TAP_BYTES + k * 32) >> 4);
                umma_bf16_ts(d_tmem, a_t + dx * 32 + k * 8, ((uint64_t)DESC_HI << 32) | bl, idesc, accumulate);
                accumulate = 1;
              }
            }
            // input row y - 1 is dead once its 12 MMAs (dy = 0) retire: free its TMEM slot now, two thirds of a row before
            // the accumulator completes, so that the loaders run a full row ahead
            if (dy == 0 && y - 1 >= c.ylo) umma_commit(&aempty[(L0 + (uint32_t)(y - 1 - c.ylo)) % TS_NA]);
          }
          umma_commit(&tfull[acc]);
          if (y == ylast)
            for (int yy = y; yy <= c.yhi; ++yy) umma_commit(&aempty[(L0 + (uint32_t)(yy - c.ylo)) % TS_NA]);
        }
        __syncwarp();
      }
      L0 += (uint32_t)(c.yhi - c.ylo + 1);
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- loaders: staging ring -> registers -> TMEM
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;  // TMEM lane = pixel of the 128-pixel strip; box row m + dx is pixel x0 - 1 + m + dx
    const uint32_t ring_addr = smem_u32(ring);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + TS_A_COL0;
    uint32_t L = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
        const uint32_t slot = L % NST, as = L % TS_NA;
        mbar_wait(&full[slot], (L / NST) & 1);
        mbar_wait(&aempty[as], ((L / TS_NA) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tile = ring_addr + slot * Cfg::SLOT_BYTES;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          uint32_t v[32];
          ld_swizzled_row128(tile, m + dx, v);
          tmem_st_32x32b_x32(lane_taddr + as * 96u + dx * 32u, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&empty[slot]);
          mbar_arrive(&afull[as]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 2 groups x 4 warps
    const int ew = warp - 6;
    uint32_t T = 0;
    if (EPI == EPI_HIDDEN)
      epilogue_hidden<NOUT, TS_NACC>(p, &tmap_out, smem + Cfg::OFF_STAGE + ew * Cfg::STAGE_BYTES, bias_s, tfull, tempty,
                                     tmem_base, ew >> 2, warp & 3, lane, T, 0, Cfg::STAGE_BUFS);
    else
      epilogue_post<NOUT, TS_NACC>(p, bias_s, tfull, tempty, tmem_base, ew >> 2, warp & 3, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair TS kernel
// The TS kernel above is bound by the shared-memory port, not the tensor pipe: per 128-pixel output row a CTA moves 72 KB
// of B operand (36 MMAs x 2 KB of weights) next to the loaders', the TMA ring's and the epilogue's traffic.  cta_group::2
// halves the dominant term: two CTAs on the SM pair of one TPC run ONE M = 256 MMA per (tap, K-step) -- 128 pixels (TMEM
// lanes) in each CTA, the 64 output channels' weights split 32 / 32 between the two shared memories -- so every CTA reads
// 1 KB instead of 2 KB of B per MMA and keeps only half of the layer's weights (36 KB).
// The pair works on the two strips (2 sp, 2 sp + 1) of the same chain and row block, so every row count is identical in
// both CTAs; the strip count is padded to even (a strip beyond the image loads zeros and stores nothing).
// Roles per CTA as in the TS kernel; only the leader's warp 1 issues MMAs.  Barriers the leader's issuer waits on (afull,
// tempty, wready) live in the leader and collect arrivals from both CTAs; barriers it signals (aempty, tfull, done) are
// arrived in both CTAs by one multicast tcgen05.commit.
constexpr int TS2_NSTAGE = 6;

template <int NOUT>
struct ConvTs2Cfg {
  static constexpr int ROW_BYTES = 128;
  static constexpr int BOX_BYTES = BOX_W * ROW_BYTES;
  static constexpr int SLOT_BYTES = round_up_c(BOX_BYTES, 1024);
  static constexpr int TAP_BYTES_FULL = NOUT * ROW_BYTES;       // one tap of the packed layer (all output channels)
  static constexpr int TAP_BYTES = (NOUT / 2) * ROW_BYTES;      // this CTA's half
  static constexpr int W_BYTES = 9 * TAP_BYTES;
  static constexpr int OFF_RING = round_up_c(W_BYTES, 1024);
  static constexpr int STAGE_BUFS = 2;
  static constexpr int STAGE_BYTES = STAGE_BUFS * 32 * NOUT * 2;
  static constexpr int OFF_STAGE = OFF_RING + TS2_NSTAGE * SLOT_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGE + EPI_WARPS * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 256;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
  static_assert(TS_NACC * NOUT <= TS_A_COL0 && TS_A_COL0 + TS_NA * 96 <= 512, "TMEM plan does not fit 512 columns");
  static_assert((2 * TS2_NSTAGE + 2 * TS_NA + 2 * TS_NACC + 3 + 2 * EPI_WARPS) * 8 + 4 <= BAR_BYTES, "barrier block overflows");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of one CTA");
};

// RES: the layer has a residual input and fetches it by TMA (epilogue_hidden_tmares); a separate instantiation so that the
// plain layers (all of DnCNN) keep their own register allocation and code (sharing one kernel cost them 5 %, ncu).
template <int NOUT, bool RES>
__global__ void __launch_bounds__(TS_THREADS, 1)
conv3x3_ts2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out,
                   const __grid_constant__ CUtensorMap tmap_res, const ConvParams p) {
  using Cfg = ConvTs2Cfg<NOUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* ring = smem + Cfg::OFF_RING;
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);  // local: TMA landed a row in the staging ring
  uint64_t* empty = full + TS2_NSTAGE;                                // local: loaders have copied it out
  uint64_t* afull = empty + TS2_NSTAGE;                               // leader: both CTAs' copies of the row are in TMEM
  uint64_t* aempty = afull + TS_NA;                                   // both (multicast): MMAs reading them completed
  uint64_t* tfull = aempty + TS_NA;                                   // both (multicast): accumulator stage complete
  uint64_t* tempty = tfull + TS_NACC;                                 // leader: both CTAs' epilogues drained the stage
  uint64_t* wbar = tempty + TS_NACC;                                  // local: this CTA's half of the weights landed
  uint64_t* wready = wbar + 1;                                        // leader: the peer's half landed
  uint64_t* done = wready + 1;                                        // both (multicast): every MMA of the launch completed
  uint64_t* rbar = done + 1;                                          // local: residual boxes (two per epilogue warp) landed
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(rbar + 2 * EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < TS2_NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 4);
    }
    for (int i = 0; i < TS_NA; ++i) {
      mbar_init(&afull[i], 8);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < TS_NACC; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(wbar, 1);
    mbar_init(wready, 1);
    mbar_init(done, 1);
    if (RES)
      for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&rbar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_ptr_s, 512);
    tmem_relinquish2();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NOUT) bias_s[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  tc_fence_before();
  cluster_sync();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t afull_c = mapa_shared(smem_u32(afull), 0);
  const uint32_t tempty_c = mapa_shared(smem_u32(tempty), 0);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      tma_prefetch_desc(&tmap);
      mbar_expect_tx(wbar, Cfg::W_BYTES);
      for (int t = 0; t < 9; ++t)
        bulk_load(smem_w + t * Cfg::TAP_BYTES, p.weights + (size_t)t * Cfg::TAP_BYTES_FULL + rank * Cfg::TAP_BYTES,
                  Cfg::TAP_BYTES, wbar);
      if (rank != 0) {
        mbar_wait(wbar, 0);
        mbar_arrive_cluster(mapa_shared(smem_u32(wready), 0));
      }
      griddep_wait();
      uint32_t L = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
          const uint32_t slot = L % TS2_NSTAGE;
          mbar_wait(&empty[slot], ((L / TS2_NSTAGE) & 1) ^ 1);
          mbar_expect_tx(&full[slot], Cfg::BOX_BYTES);
          tma_load_4d(ring + slot * Cfg::SLOT_BYTES, &tmap, &full[slot], 0, c.x0 - 1, y, c.b);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------------------------------------------------------- MMA issuer of the pair
      constexpr uint32_t idesc = make_idesc_bf16(2 * TILE_M, NOUT);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
      const uint32_t w_lo = (smem_u32(smem_w) >> 4) | 0x10000u;
      mbar_wait(wbar, 0);
      mbar_wait_cluster(wready, 0);
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
            mbar_wait(&afull[q % TS_NA], (q / TS_NA) & 1);
            ++waited;
          }
          const uint32_t acc = T % TS_NACC;
          mbar_wait(&tempty[acc], ((T / TS_NACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * NOUT;
          if (elect_one()) {
            uint32_t accumulate = 0;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int yy = y + dy - 1;
              if (yy < 0 || yy >= p.H) continue;
              const uint32_t q = L0 + (uint32_t)(yy - c.ylo);
              const uint32_t a_t = tmem_base + TS_A_COL0 + (q % TS_NA) * 96u;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t bl = w_lo + (uint32_t)(((dy * 3 + dx) * Cfg::TAP_BYTES + k * 32) >> 4);
                  umma_bf16_ts2(d_tmem, a_t + dx * 32 + k * 8, ((uint64_t)DESC_HI << 32) | bl, idesc, accumulate);
                  accumulate = 1;
                }
              }
              if (dy == 0 && y - 1 >= c.ylo) umma_commit2(&aempty[(L0 + (uint32_t)(y - 1 - c.ylo)) % TS_NA], 3);
            }
            umma_commit2(&tfull[acc], 3);
            if (y == ylast)
              for (int yy = y; yy <= c.yhi; ++yy) umma_commit2(&aempty[(L0 + (uint32_t)(yy - c.ylo)) % TS_NA], 3);
          }
          __syncwarp();
        }
        L0 += (uint32_t)(c.yhi - c.ylo + 1);
      }
      if (elect_one()) umma_commit2(done, 3);
      __syncwarp();
    }
    mbar_wait(done, 0);  // both CTAs: no MMA still reads this CTA's shared / tensor memory, no commit is still in flight
  } else if (warp < 6) {
    // ---------------------------------------------------------------- loaders: staging ring -> registers -> TMEM
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t ring_addr = smem_u32(ring);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + TS_A_COL0;
    uint32_t L = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
        const uint32_t slot = L % TS2_NSTAGE, as = L % TS_NA;
        mbar_wait(&full[slot], (L / TS2_NSTAGE) & 1);
        mbar_wait(&aempty[as], ((L / TS_NA) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tile = ring_addr + slot * Cfg::SLOT_BYTES;
        // (reading the three shifted rows into registers before the TMEM slot is awaited was tried: 113 -> 122 us per layer,
        // 96 live registers per loader thread and shared-memory reads bunched against the operand fetch)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          uint32_t v[32];
          ld_swizzled_row128(tile, m + dx, v);
          tmem_st_32x32b_x32(lane_taddr + as * 96u + dx * 32u, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&empty[slot]);
          mbar_arrive_remote(afull_c + as * 8u);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 2 groups x 4 warps
    const int ew = warp - 6;
    uint32_t T = 0;
    if (RES)
      epilogue_hidden_tmares<NOUT, TS_NACC>(p, &tmap_out, &tmap_res, smem + Cfg::OFF_STAGE + ew * Cfg::STAGE_BYTES,
                                            rbar + 2 * ew, bias_s, tfull, tempty, tmem_base, ew >> 2, warp & 3, lane, tempty_c);
    else
      epilogue_hidden<NOUT, TS_NACC>(p, &tmap_out, smem + Cfg::OFF_STAGE + ew * Cfg::STAGE_BYTES, bias_s, tfull, tempty,
                                     tmem_base, ew >> 2, warp & 3, lane, T, tempty_c, Cfg::STAGE_BUFS);
  }
  tc_fence_before();
  cluster_sync();  // neither CTA may exit (or free tensor memory) while its partner can still signal or read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ layer-chain kernel
// With few chains a layer is ~3 us of MMAs wrapped in ~5 us of launch, prologue, first-load latency and drain, so the 18
// hidden layers of DnCNN are also available as ONE persistent launch: every CTA walks the layers, ping-ponging between
// the two activation buffers, reloading the 73.7 KB of weights per layer (prefetched as soon as the previous layer's MMAs
// retire) and meeting the other CTAs at a grid-wide barrier between layers (layer l+1 needs halo rows and neighbouring
// strips produced by other CTAs).  grid <= #SMs with one CTA per SM, so all CTAs are co-resident and the barrier cannot
// deadlock.  Same roles and pipelines as conv3x3_ts_kernel; counters and mbarrier phases simply run on across layers.
constexpr int HIDDEN_LAYER_STRIDE = 9 * 64 * 128 + 1024;  // packed weights (73 728 B) + bias, rounded to 1 KB

struct ChainParams {
  int n_layers;
  const uint8_t* weights0;   // packed weights of the first layer of the chain; layer l at + l * HIDDEN_LAYER_STRIDE
  unsigned int* barrier;     // zero-initialised counter in global memory
};

__device__ __forceinline__ void grid_barrier_arrive_wait(unsigned int* counter, unsigned int target) {
  __threadfence();  // publish this CTA's completed stores (already awaited by their issuers) at gpu scope
  atomicAdd(counter, 1u);
  unsigned int v;
  long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v < target && clock64() - t0 > 20000000000LL) {  // ~10 s: a protocol bug must not hang the GPU
      printf("psgla_b200: grid barrier timed out (block %d, %u of %u)\n", (int)blockIdx.x, v, target);
      __trap();
    }
  } while (v < target);
  fence_proxy_async_global();  // order the TMA (async proxy) loads that follow after the acquire
}

__global__ void __launch_bounds__(TS_THREADS, 1)
conv3x3_ts_chain_kernel(const __grid_constant__ CUtensorMap map_ld0, const __grid_constant__ CUtensorMap map_ld1,
                        const __grid_constant__ CUtensorMap map_st0, const __grid_constant__ CUtensorMap map_st1,
                        const ConvParams p, const ChainParams cp) {
  // layer l reads buffer (l & 1) through map_ld{l&1} and writes buffer ((l + 1) & 1) through map_st{(l+1)&1}
  constexpr int NOUT = 64;
  using Cfg = ConvTsCfg<NOUT, EPI_HIDDEN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + TS_NSTAGE;
  uint64_t* afull = empty + TS_NSTAGE;
  uint64_t* aempty = afull + TS_NA;
  uint64_t* tfull = aempty + TS_NA;
  uint64_t* tempty = tfull + TS_NACC;
  uint64_t* wbar = tempty + TS_NACC;   // weights of the current layer have landed
  uint64_t* wfree = wbar + 1;          // every MMA of the layer that used them has completed
  uint64_t* ldone = wfree + 1;         // the eight epilogue warps have finished (and flushed) their rows of the layer
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ldone + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < TS_NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 4);
    }
    for (int i = 0; i < TS_NA; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < TS_NACC; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(wbar, 1);
    mbar_init(wfree, 1);
    mbar_init(ldone, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer (+ the CTA's voice at the grid barrier)
      tma_prefetch_desc(&map_ld0);
      tma_prefetch_desc(&map_ld1);
      uint32_t L = 0;
      for (int l = 0; l < cp.n_layers; ++l) {
        if (l > 0) mbar_wait(wfree, (l - 1) & 1);  // the previous layer's MMAs no longer read the weight buffer
        mbar_expect_tx(wbar, Cfg::W_BYTES);
        bulk_load(smem_w, cp.weights0 + (size_t)l * HIDDEN_LAYER_STRIDE, Cfg::W_BYTES, wbar);
        if (l == 0) {
          griddep_wait();
        } else {
          mbar_wait(ldone, (l - 1) & 1);  // this CTA's outputs of layer l-1 are complete in global memory
          grid_barrier_arrive_wait(cp.barrier, (unsigned)l * gridDim.x);
        }
        const CUtensorMap* map = (l & 1) ? &map_ld1 : &map_ld0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          const ItemCoord c = decode_item(p, item);
          for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
            const uint32_t slot = L % TS_NSTAGE;
            mbar_wait(&empty[slot], ((L / TS_NSTAGE) & 1) ^ 1);
            mbar_expect_tx(&full[slot], Cfg::BOX_BYTES);
            tma_load_4d(ring + slot * Cfg::SLOT_BYTES, map, &full[slot], 0, c.x0 - 1, y, c.b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, NOUT);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
    const uint32_t w_lo = (smem_u32(smem_w) >> 4) | 0x10000u;
    uint32_t L0 = 0, T = 0;
    for (int l = 0; l < cp.n_layers; ++l) {
      mbar_wait(wbar, l & 1);
      tc_fence_after();
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        int waited = 0;
        const int ylast = c.y0 + c.rcur - 1;
        for (int y = c.y0; y <= ylast; ++y, ++T) {
          const int need = min(y + 1, c.yhi) - c.ylo + 1;
          while (waited < need) {
            const uint32_t q = L0 + waited;
            mbar_wait(&afull[q % TS_NA], (q / TS_NA) & 1);
            ++waited;
          }
          const uint32_t acc = T % TS_NACC;
          mbar_wait(&tempty[acc], ((T / TS_NACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * NOUT;
          if (elect_one()) {
            uint32_t accumulate = 0;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int yy = y + dy - 1;
              if (yy < 0 || yy >= p.H) continue;
              const uint32_t q = L0 + (uint32_t)(yy - c.ylo);
              const uint32_t a_t = tmem_base + TS_A_COL0 + (q % TS_NA) * 96u;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t bl = w_lo + (uint32_t)(((dy * 3 + dx) * Cfg::TAP_BYTES + k * 32) >> 4);
                  umma_bf16_ts(d_tmem, a_t + dx * 32 + k * 8, ((uint64_t)DESC_HI << 32) | bl, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            umma_commit(&tfull[acc]);
            if (y - 1 >= c.ylo) umma_commit(&aempty[(L0 + (uint32_t)(y - 1 - c.ylo)) % TS_NA]);
            if (y == ylast)
              for (int yy = y; yy <= c.yhi; ++yy) umma_commit(&aempty[(L0 + (uint32_t)(yy - c.ylo)) % TS_NA]);
          }
          __syncwarp();
        }
        L0 += (uint32_t)(c.yhi - c.ylo + 1);
      }
      if (elect_one()) umma_commit(wfree);  // arrives once every MMA issued so far has completed
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- loaders: staging ring -> registers -> TMEM
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const uint32_t ring_addr = smem_u32(ring);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + TS_A_COL0;
    uint32_t L = 0;
    for (int l = 0; l < cp.n_layers; ++l) {
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        for (int y = c.ylo; y <= c.yhi; ++y, ++L) {
          const uint32_t slot = L % TS_NSTAGE, as = L % TS_NA;
          mbar_wait(&full[slot], (L / TS_NSTAGE) & 1);
          mbar_wait(&aempty[as], ((L / TS_NA) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tile = ring_addr + slot * Cfg::SLOT_BYTES;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            uint32_t v[32];
            ld_swizzled_row128(tile, m + dx, v);
            tmem_st_32x32b_x32(lane_taddr + as * 96u + dx * 32u, v);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&empty[slot]);
            mbar_arrive(&afull[as]);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 2 groups x 4 warps
    const int ew = warp - 6;
    uint32_t T = 0;
    for (int l = 0; l < cp.n_layers; ++l) {
      const float* bias = reinterpret_cast<const float*>(cp.weights0 + (size_t)l * HIDDEN_LAYER_STRIDE + Cfg::W_BYTES);
      epilogue_hidden<NOUT, TS_NACC>(p, ((l + 1) & 1) ? &map_st1 : &map_st0, smem + Cfg::OFF_STAGE + ew * Cfg::STAGE_BYTES, bias,
                                     tfull, tempty, tmem_base, ew >> 2, warp & 3, lane, T, 0, Cfg::STAGE_BUFS);
      // epilogue_hidden ends with cp.async.bulk.wait_group 0 on the issuing lane: this warp's stores are complete
      __syncwarp();
      if (lane == 0) mbar_arrive(ldone);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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

// Activation tensor maps (bf16 NHWC, dims {C, W, H, B}).  box_w = BOX_W: input rows with halo (OOB zero fill = the
// convolution's zero padding); box_w = 32: one epilogue warp's output box.  Encoding costs microseconds on the host,
// so the few (pointer, shape) combinations of a run are cached per thread.
struct MapKey {
  const void* ptr;
  int B, H, W, C, box_w;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && B == o.B && H == o.H && W == o.W && C == o.C && box_w == o.box_w;
  }
};
static int get_act_tensor_map(CUtensorMap* map, const void* ptr, int B, int H, int W, int C, int box_w) {
  static thread_local std::vector<std::pair<MapKey, CUtensorMap>> cache;
  const MapKey key{ptr, B, H, W, C, box_w};
  for (const auto& e : cache)
    if (e.first == key) {
      *map = e.second;
      return PSGLA_OK;
    }
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return set_error(PSGLA_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_w, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = (C == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  if (cache.size() >= 64) cache.erase(cache.begin());
  cache.emplace_back(key, *map);
  return PSGLA_OK;
}

// Work items = (chain, 128-pixel strip, block of R output rows), dealt round-robin to one persistent CTA per SM.
// R trades the 2 halo rows an item re-reads against the tail when items do not divide by the CTA count:
// pick the R that minimises (items per CTA) * (R + 1).
static void plan_items(ConvParams* p) {
  p->strips = (p->W + TILE_M - 1) / TILE_M;
  const int sms = num_sms();
  int best = 1;
  long long best_cost = -1;
  const int cands[] = {32, 16, 8, 4, 2, 1};
  for (int R : cands) {
    const long long items = (long long)p->B * p->strips * ((p->H + R - 1) / R);
    const long long cost = ((items + sms - 1) / sms) * (std::min(R, p->H) + 1);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = R;
    }
  }
  const char* e = getenv("PSGLA_CONV_ROWS");
  if (e && atoi(e) > 0) best = atoi(e);
  p->R = best;
  p->row_blocks = (p->H + best - 1) / best;
  p->n_items = p->B * p->strips * p->row_blocks;
}

template <int CIN, int NOUT, int EPI>
static int launch_conv(const void* in, void* out_bf16, ConvParams p, cudaStream_t st) {
  using Cfg = ConvCfg<CIN, NOUT, EPI>;
  CUtensorMap map, map_out;
  int rc = get_act_tensor_map(&map, in, p.B, p.H, p.W, CIN, BOX_W);
  if (rc) return rc;
  if (EPI == EPI_HIDDEN) {
    rc = get_act_tensor_map(&map_out, out_bf16, p.B, p.H, p.W, NOUT, 32);
    if (rc) return rc;
  } else {
    map_out = map;  // unused by the fused-post epilogue
  }
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(conv3x3_kernel<CIN, NOUT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  plan_items(&p);
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CONV_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // pairs with griddepcontrol.* in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_kernel<CIN, NOUT, EPI>, map, map_out, p));
  return PSGLA_OK;
}

template <int NOUT, int EPI>
static int launch_conv_ts(const void* in, void* out_bf16, ConvParams p, cudaStream_t st) {
  using Cfg = ConvTsCfg<NOUT, EPI>;
  CUtensorMap map, map_out;
  int rc = get_act_tensor_map(&map, in, p.B, p.H, p.W, 64, BOX_W);
  if (rc) return rc;
  if (EPI == EPI_HIDDEN) {
    rc = get_act_tensor_map(&map_out, out_bf16, p.B, p.H, p.W, NOUT, 32);
    if (rc) return rc;
  } else {
    map_out = map;
  }
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(conv3x3_ts_kernel<NOUT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  plan_items(&p);
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_ts_kernel<NOUT, EPI>, map, map_out, p));
  return PSGLA_OK;
}

// A-operand source of the 64-input-channel layers: tensor memory (default) or shared memory (PSGLA_CONV_SS=1, kept for
// A/B measurements and as the path of the 3-channel first layer).
// Work items of the pair kernel: (chain, row block, strip) with the strip count padded to even and the strip index
// fastest, so that items 2i and 2i + 1 -- the two CTAs of a cluster -- share chain and rows.
static void plan_items_pair(ConvParams* p, int n_clusters) {
  p->strips = ((p->W + TILE_M - 1) / TILE_M + 1) & ~1;
  int best = 1;
  long long best_cost = -1;
  const int cands[] = {32, 16, 8, 4, 2, 1};
  for (int R : cands) {
    const long long pairs = (long long)p->B * (p->strips / 2) * ((p->H + R - 1) / R);
    const long long cost = ((pairs + n_clusters - 1) / n_clusters) * (std::min(R, p->H) + 1);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = R;
    }
  }
  const char* e = getenv("PSGLA_CONV_ROWS");
  if (e && atoi(e) > 0) best = atoi(e);
  p->R = best;
  p->row_blocks = (p->H + best - 1) / best;
  p->n_items = p->B * p->strips * p->row_blocks;
}

template <int NOUT, bool RES>
static int launch_conv_ts2_t(const void* in, void* out_bf16, ConvParams p, cudaStream_t st) {
  using Cfg = ConvTs2Cfg<NOUT>;
  CUtensorMap map, map_out;
  int rc = get_act_tensor_map(&map, in, p.B, p.H, p.W, 64, BOX_W);
  if (rc) return rc;
  rc = get_act_tensor_map(&map_out, out_bf16, p.B, p.H, p.W, NOUT, 32);
  if (rc) return rc;
  CUtensorMap map_res = map_out;  // unused without a residual input
  if (p.res1) {
    rc = get_act_tensor_map(&map_res, p.res1, p.B, p.H, p.W, NOUT, 32);
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static int max_clusters = 0;  // CTA pairs the device holds at once (one CTA per SM)
  if (!max_clusters) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(conv3x3_ts2_kernel<NOUT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    cfg.gridDim = dim3((unsigned)(num_sms() & ~1));
    int n = 0;
    PSGLA_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, conv3x3_ts2_kernel<NOUT, RES>, &cfg));
    max_clusters = n > 0 ? std::min(n, num_sms() / 2) : num_sms() / 2;
    if (getenv("PSGLA_VERBOSE")) fprintf(stderr, "psgla_b200: conv3x3_ts2_kernel: %d co-resident CTA pairs (occupancy query %d)\n", max_clusters, n);
  }
  plan_items_pair(&p, max_clusters);
  const int pairs = p.n_items / 2;
  cfg.gridDim = dim3((unsigned)(2 * std::min(pairs, max_clusters)));
  if (getenv("PSGLA_VERBOSE")) fprintf(stderr, "psgla_b200: pair conv B=%d H=%d W=%d: R=%d items=%d grid=%u\n", p.B, p.H, p.W, p.R, p.n_items, cfg.gridDim.x);
  PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_ts2_kernel<NOUT, RES>, map, map_out, map_res, p));
  return PSGLA_OK;
}

template <int NOUT>
static int launch_conv_ts2(const void* in, void* out_bf16, const ConvParams& p, cudaStream_t st) {
  return p.res1 ? launch_conv_ts2_t<NOUT, true>(in, out_bf16, p, st) : launch_conv_ts2_t<NOUT, false>(in, out_bf16, p, st);
}

// PSGLA_CONV_PAIR=0 falls back to the single-CTA TS kernel (A/B runs)
static bool conv_use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PSGLA_CONV_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static bool conv_use_ts() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PSGLA_CONV_SS");
    v = (e && atoi(e) != 0) ? 0 : 1;
  }
  return v == 1;
}
static int launch_hidden64(const void* in, void* out, const ConvParams& p, cudaStream_t st) {
  if (!conv_use_ts()) return launch_conv<64, 64, EPI_HIDDEN>(in, out, p, st);
  return conv_use_pair() ? launch_conv_ts2<64>(in, out, p, st) : launch_conv_ts<64, EPI_HIDDEN>(in, out, p, st);
}
static int launch_last(const void* in, const ConvParams& p, cudaStream_t st) {
  return conv_use_ts() ? launch_conv_ts<16, EPI_POST>(in, nullptr, p, st) : launch_conv<64, 16, EPI_POST>(in, nullptr, p, st);
}

// The 18 hidden layers as one persistent launch (conv3x3_ts_chain_kernel).  buf0 holds the input of the first layer of the
// chain; layers alternate buf0 -> buf1 -> buf0 ...; the result is in buf[n_layers & 1].
static int launch_hidden_chain(void* buf0, void* buf1, int n_layers, const uint8_t* weights0, unsigned int* barrier,
                               ConvParams p, cudaStream_t st) {
  using Cfg = ConvTsCfg<64, EPI_HIDDEN>;
  CUtensorMap ld0, ld1, st0, st1;
  int rc = get_act_tensor_map(&ld0, buf0, p.B, p.H, p.W, 64, BOX_W);
  if (!rc) rc = get_act_tensor_map(&ld1, buf1, p.B, p.H, p.W, 64, BOX_W);
  if (!rc) rc = get_act_tensor_map(&st0, buf0, p.B, p.H, p.W, 64, 32);
  if (!rc) rc = get_act_tensor_map(&st1, buf1, p.B, p.H, p.W, 64, 32);
  if (rc) return rc;
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(conv3x3_ts_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  plan_items(&p);
  p.relu = 1;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  PSGLA_CUDA_TRY(cudaMemsetAsync(barrier, 0, sizeof(unsigned int), st));
  ChainParams cp{n_layers, weights0, barrier};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = nullptr;  // plain stream order: the memset above must be complete, and every CTA must be free to start
  cfg.numAttrs = 0;
  PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_ts_chain_kernel, ld0, ld1, st0, st1, p, cp));
  return PSGLA_OK;
}

// Measured on B200 (1-16 chains of 256 x 256): the chain is NOT faster than the per-layer launches (181.7 vs 187.8 us per
// DnCNN application at one chain in round 1; 174.6 vs 176.9 us in round 2, and 8-20 % SLOWER at 4-16 chains).  Round 2 also
// tried replacing the grid barrier by per-item flags (an item waits only for the <= 9 items whose rows / halo pixels it
// reads): 224 us -- the acquire-polls of the neighbours' flags cost more than the barrier's single counter.  A layer's ~8 us at that size is pipeline fill and drain inside the CTA -- TMA load
// latency, three rows through the loader warps before the first MMA, the last row's epilogue and store -- not launch
// overhead, which PDL already overlaps; the grid barrier costs what the kernel boundary cost.  The kernel stays as a
// tested alternative (PSGLA_CHAIN=1) and as the base for keeping a CTA's own rows on chip between layers.
static bool use_chain(const ConvParams&) {
  const char* e = getenv("PSGLA_CHAIN");  // read per call: tests and A/B scripts switch it inside one process
  return conv_use_ts() && e && atoi(e) == 1;
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

void pack_conv3x3_swizzled(const float* w, int nout_real, int cin_real, int nout_pad, int cin_pad, uint8_t* dst);

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
    const float* w = weights_host[l];  // OIHW [nout_real][cin_real][3][3]
    PSGLA_REQUIRE(w != nullptr, "layer %d: null weight pointer", l);
    pack_conv3x3_swizzled(w, nout_real, cin_real, li.nout, li.cin, host.data() + li.w_off);
    if (biases_host && biases_host[l])
      std::memcpy(&host[li.b_off], biases_host[l], (size_t)nout_real * 4);
  }
  PSGLA_CUDA_TRY(cudaMemcpyAsync(packed_dev, host.data(), host.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  PSGLA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));  // `host` dies at return
  return PSGLA_OK;
}

extern "C" size_t psgla_dncnn_workspace_bytes(psgla_img_shape s) {
  // two ping-pong activation buffers (each rounded up to 1 KB) + 4 KB holding the layer-chain kernel's barrier / item flags
  return 2 * (((size_t)s.B * s.H * s.W * 64 * 2 + 1023) / 1024 * 1024) + 4096 + 1024;
}

static int check_shape(const psgla_img_shape& s) {
  PSGLA_REQUIRE(s.B > 0 && s.H > 0 && s.W > 0 && s.C == 3, "image shape must be [B>0][3][H>0][W>0], got [%d][%d][%d][%d]",
                s.B, s.C, s.H, s.W);
  return PSGLA_OK;
}

// PSGLA_CONV_ALTERNATE=0: every layer walks its items front to back (A/B runs)
static bool alternate_items() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PSGLA_CONV_ALTERNATE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
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

// The last layer (64 -> 3) with the Langevin post (and optionally the next iteration's pre) in its epilogue.
static int last_layer_post(int depth, const uint8_t* packed, const psgla_img_shape& shape, const void* hidden, const float* base_dev,
                           const psgla_post_params* post, float* x_out_dev, float* sample_dev, float* mean_dev,
                           float* mean2_dev, const psgla_next_pre* next, cudaStream_t st) {
  const LayerInfo li = layer_info(depth, depth - 1);
  ConvParams p = base_params(shape, packed, li);
  p.base = base_dev;
  p.x_out = x_out_dev;
  p.sample = sample_dev;
  p.mean = mean_dev;
  p.mean2 = mean2_dev;
  p.gain = post->gain;
  p.base_scale = 1.0f;  // DnCNN is a residual denoiser: X+ = base + gain * R
  p.w_old = post->w_old;
  p.w_new = post->w_new;
  p.reverse = alternate_items() ? ((depth - 1) & 1) : 0;
  int rc = set_next_pre(&p, next);
  if (rc) return rc;
  return launch_last(hidden, p, st);
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
    return launch_last(in_dev, p, st);
  }
  return layer == 0 ? launch_conv<16, 64, EPI_HIDDEN>(in_dev, out_dev, p, st) : launch_hidden64(in_dev, out_dev, p, st);
}

extern "C" int psgla_dncnn_residual_post(int depth, const void* packed_dev, psgla_img_shape shape,
                                         const void* den_in_dev, void* workspace_dev, size_t workspace_bytes,
                                         const float* base_dev, const psgla_post_params* post, float* x_out_dev,
                                         float* sample_dev, float* mean_dev, float* mean2_dev, void* stream) {
  return psgla_dncnn_residual_post_next(depth, packed_dev, shape, den_in_dev, workspace_dev, workspace_bytes, base_dev, post,
                                        x_out_dev, sample_dev, mean_dev, mean2_dev, nullptr, stream);
}

extern "C" int psgla_dncnn_residual_post_next(int depth, const void* packed_dev, psgla_img_shape shape,
                                              const void* den_in_dev, void* workspace_dev, size_t workspace_bytes,
                                              const float* base_dev, const psgla_post_params* post, float* x_out_dev,
                                              float* sample_dev, float* mean_dev, float* mean2_dev,
                                              const psgla_next_pre* next, void* stream) {
  PSGLA_REQUIRE(packed_dev && den_in_dev && workspace_dev && post && x_out_dev && depth >= 2,
                "psgla_dncnn_residual_post: bad argument");
  {
    const int rcn = check_next_pre(next, shape);
    if (rcn) return rcn;
  }
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
  {
    ConvParams p = base_params(shape, packed, layer_info(depth, 0));
    if (depth > 3 && use_chain(p)) {
      // first layer, then all hidden layers in one persistent launch
      p.relu = 1;
      rc = launch_conv<16, 64, EPI_HIDDEN>(cur, ws[0], p, st);
      if (rc) return rc;
      const LayerInfo l1 = layer_info(depth, 1);
      unsigned int* barrier = reinterpret_cast<unsigned int*>((uint8_t*)workspace_dev + 2 * half);
      rc = launch_hidden_chain(ws[0], ws[1], depth - 2, packed + l1.w_off, barrier, base_params(shape, packed, l1), st);
      if (rc) return rc;
      cur = ws[(depth - 2) & 1];
    } else {
      for (int l = 0; l < depth - 1; ++l) {
        const LayerInfo li = layer_info(depth, l);
        ConvParams pl = base_params(shape, packed, li);
        pl.relu = 1;
        pl.reverse = alternate_items() ? (l & 1) : 0;
        rc = (l == 0) ? launch_conv<16, 64, EPI_HIDDEN>(cur, ws[l & 1], pl, st) : launch_hidden64(cur, ws[l & 1], pl, st);
        if (rc) return rc;
        cur = ws[l & 1];
      }
    }
  }
  return last_layer_post(depth, packed, shape, cur, base_dev, post, x_out_dev, sample_dev, mean_dev, mean2_dev, next, st);
}

extern "C" int psgla_dncnn_last_layer_post_next(int depth, const void* packed_dev, psgla_img_shape shape,
                                                const void* hidden_dev, const float* base_dev,
                                                const psgla_post_params* post, float* x_out_dev, float* sample_dev,
                                                float* mean_dev, float* mean2_dev, const psgla_next_pre* next, void* stream) {
  PSGLA_REQUIRE(packed_dev && hidden_dev && base_dev && post && x_out_dev && depth >= 2,
                "psgla_dncnn_last_layer_post_next: bad argument");
  PSGLA_REQUIRE((mean_dev == nullptr) == (mean2_dev == nullptr), "mean and mean2 must be given together");
  int rc = check_next_pre(next, shape);
  if (rc) return rc;
  rc = check_shape(shape);
  if (rc) return rc;
  return last_layer_post(depth, (const uint8_t*)packed_dev, shape, hidden_dev, base_dev, post, x_out_dev, sample_dev, mean_dev,
                         mean2_dev, next, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ internal API (drunet.cu)
namespace psgla {

// fp32 OIHW [nout_real][cin_real][3][3] -> bf16 [tap][nout_pad][cin_pad], 128B- (cin_pad = 64) or 32B- (16) swizzled as the
// resident-weight kernels expect; dst must be zero-initialised (padding rows / channels stay zero).
void pack_conv3x3_swizzled(const float* w, int nout_real, int cin_real, int nout_pad, int cin_pad, uint8_t* dst) {
  const int row_bytes = cin_pad * 2;
  for (int tap = 0; tap < 9; ++tap)
    for (int n = 0; n < nout_real; ++n)
      for (int k = 0; k < cin_real; ++k) {
        const float val = w[((size_t)n * cin_real + k) * 9 + tap];
        const int kbyte = k * 2;
        int chunk = kbyte >> 4;
        chunk ^= (row_bytes == 128) ? (n & 7) : ((n >> 2) & 1);
        const size_t phys = (size_t)tap * nout_pad * row_bytes + (size_t)n * row_bytes + chunk * 16 + (kbyte & 15);
        const uint16_t h = f32_to_bf16_rn(val);
        std::memcpy(dst + phys, &h, 2);
      }
}

int next_layer_direction() {
  static thread_local int dir = 0;
  if (!alternate_items()) return 0;
  dir ^= 1;
  return dir;
}

int conv64_hidden(const void* in, void* out, const uint8_t* w, const float* bias, int B, int H, int W, int relu,
                  const void* res1, const void* res2, cudaStream_t st) {
  ConvParams p{};
  p.B = B, p.H = H, p.W = W;
  p.reverse = next_layer_direction();
  p.weights = w;
  p.bias = bias;
  p.relu = relu;
  p.res1 = (const __nv_bfloat16*)res1;
  p.res2 = (const __nv_bfloat16*)res2;
  return launch_hidden64(in, out, p, st);
}

int conv_first16(const void* in16, void* out, const uint8_t* w, const float* bias, int B, int H, int W, int relu,
                 cudaStream_t st) {
  ConvParams p{};
  p.B = B, p.H = H, p.W = W;
  p.weights = w;
  p.bias = bias;
  p.relu = relu;
  return launch_conv<16, 64, EPI_HIDDEN>(in16, out, p, st);
}

int check_next_pre(const psgla_next_pre* next, const psgla_img_shape& s) {
  if (!next) return PSGLA_OK;
  PSGLA_REQUIRE(next->pre && next->mask_dev && next->y_dev && next->base_dev && next->den_in_dev,
                "psgla_next_pre: null pointer");
  PSGLA_REQUIRE((next->mask_B == 1 || next->mask_B == s.B) && (next->y_B == 1 || next->y_B == s.B),
                "psgla_next_pre: mask_B / y_B must be 1 or B");
  PreArgs a;
  return fill_pre(next->pre, &a);
}

static int set_next_pre(ConvParams* p, const psgla_next_pre* next) {
  p->nx_enable = 0;
  if (!next) return PSGLA_OK;
  int rc = fill_pre(next->pre, &p->nx);
  if (rc) return rc;
  p->nx.chw = 3LL * p->H * p->W;
  p->nx_enable = 1;
  p->nx_mask = next->mask_dev;
  p->nx_y = next->y_dev;
  p->nx_mask_B = next->mask_B;
  p->nx_y_B = next->y_B;
  p->nx_base = next->base_dev;
  p->nx_den_in = (__nv_bfloat16*)next->den_in_dev;
  return PSGLA_OK;
}

int conv_last_post(const void* in, const uint8_t* w, const float* bias, int B, int H, int W, const float* base,
                   float base_scale, float gain, float w_old, float w_new, float* x_out, float* sample, float* mean,
                   float* mean2, const psgla_next_pre* next, cudaStream_t st) {
  ConvParams p{};
  p.B = B, p.H = H, p.W = W;
  p.weights = w;
  p.bias = bias;
  p.base = base;
  p.base_scale = base_scale;
  p.gain = gain;
  p.w_old = w_old;
  p.w_new = w_new;
  p.x_out = x_out;
  p.sample = sample;
  p.mean = mean;
  p.mean2 = mean2;
  int rc = set_next_pre(&p, next);
  if (rc) return rc;
  return launch_last(in, p, st);
}

}  // namespace psgla

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
    tmem_alloc(tptr, 128);
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
  }
  if (mode == 2) {
    // A through tensor memory: every thread copies its (shifted) row into TMEM columns [64, 96), then TS MMAs
    mbar_wait(&bar[0], 0);
    uint32_t v[32];
    ld_swizzled_row128(smem_u32(sa), (int)threadIdx.x + row_shift, v);
    tmem_st_32x32b_x32(tbase + ((uint32_t)(warp * 32) << 16) + 64, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, 64);
      for (int k = 0; k < 4; ++k)
        umma_bf16_ts(tbase, tbase + 64 + k * 8, make_smem_desc(smem_u32(sb) + k * 32, 1024, LAYOUT_SW128, 0), idesc, k > 0);
      umma_commit(&bar[1]);
    }
  } else if (threadIdx.x == 0) {
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
    tmem_dealloc(tbase, 128);
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
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, d_dev, row_shift, mode);
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}

// ------------------------------------------------------------------------------------------------ CTA-pair self-test
// D[256 x 64] = A[256 x 64] B[64 x 64]^T with one cta_group::2 MMA chain: CTA r of the pair holds A rows [128 r, 128 r + 128)
// in tensor memory (copied there by its own threads) and B rows [32 r, 32 r + 32) in shared memory; the leader issues,
// both read their 128 accumulator lanes back.  Pins down the operand split, the multicast commit and the remote arrive
// the conv kernel relies on.  mode 0: A from TMEM (TS); mode 1: A from shared memory (SS).
namespace psgla {
__global__ void __launch_bounds__(128, 1)
selftest_umma2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float* __restrict__ d, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                // 128 rows x 128 B
  uint8_t* sb = smem + 16 * 1024;    // 32 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20 * 1024);  // 0: TMA landed, 1: MMAs done, 2 (leader): operands ready
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 2);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tptr, 128);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], 128 * 128 + 32 * 128);
    tma_load_2d(sa, &map_a, &bar[0], 0, (int)rank * 128);
    tma_load_2d(sb, &map_b, &bar[0], 0, (int)rank * 32);
  }
  mbar_wait(&bar[0], 0);
  if (mode == 0) {
    uint32_t v[32];
    ld_swizzled_row128(smem_u32(sa), (int)threadIdx.x, v);
    tmem_st_32x32b_x32(tbase + ((uint32_t)(warp * 32) << 16) + 64, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bar[2]), 0));  // this CTA's operands are in place
  if (rank == 0 && warp == 0) {
    mbar_wait_cluster(&bar[2], 0);
    tc_fence_after();
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 64);
      for (int k = 0; k < 4; ++k) {
        const uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 1024, LAYOUT_SW128, 0);
        if (mode == 0)
          umma_bf16_ts2(tbase, tbase + 64 + k * 8, bd, idesc, k > 0);
        else
          umma_bf16_ss2(tbase, make_smem_desc(smem_u32(sa) + k * 32, 1024, LAYOUT_SW128, 0), bd, idesc, k > 0);
      }
      umma_commit2(&bar[1], 3);
    }
    __syncwarp();
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tbase + ((uint32_t)(warp * 32) << 16) + half * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      d[(size_t)(rank * 128 + warp * 32 + lane) * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tbase, 128);
  }
}
}  // namespace psgla

extern "C" int psgla_selftest_umma2(const void* a_dev, const void* b_dev, float* d_dev, int mode, void* stream) {
  PSGLA_REQUIRE(a_dev && b_dev && d_dev && (mode == 0 || mode == 1), "psgla_selftest_umma2: bad argument");
  PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
  if (!enc) return set_error(PSGLA_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  CUtensorMap ma, mb;
  const cuuint32_t estr[2] = {1, 1};
  const cuuint64_t strides[1] = {128};
  {
    const cuuint64_t dims[2] = {64, 256};
    const cuuint32_t box[2] = {64, 128};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_dev), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "tensor map A: CUresult %d", (int)r);
  }
  {
    const cuuint64_t dims[2] = {64, 64};
    const cuuint32_t box[2] = {64, 32};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_dev), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PSGLA_E_BADARG, "tensor map B: CUresult %d", (int)r);
  }
  const int smem = 22 * 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, selftest_umma2_kernel, ma, mb, d_dev, mode));
  return PSGLA_OK;
}

// ------------------------------------------------------------------------------------------------ MMA rate probe
namespace psgla {
// One CTA per block issues `iters` x 4 K-steps of M128 x N x K16 bf16 MMAs back to back on zeroed operands and reports
// the cycles one MMA took.  mode 0: A and B from shared memory (SS); 1: SS with the A start address shifted by one
// 128-byte row (the conv kernel's dx tap shift); 2: A from tensor memory (TS).
template <int mode>  // compile-time so that the issue loop is nothing but the MMAs
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 136 rows x 128 B
  uint8_t* sb = smem + 18 * 1024;     // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 52 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 16);
  for (int i = threadIdx.x; i < 52 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
    const uint32_t a_lo = ((smem_u32(sa) + (mode == 1 ? 128u : 0u)) >> 4) | 0x10000u;
    const uint32_t b_lo = (smem_u32(sb) >> 4) | 0x10000u;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < (mode >= 5 ? 0 : iters); ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 2)
            umma_bf16_ts(tbase, tbase + 256 + k * 8, ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
          else if (mode >= 5)
            ;  // handled below
          else if (mode == 3)  // TS, consecutive MMAs alternate between two accumulators (no back-to-back dependency)
            umma_bf16_ts(tbase + ((it * 4 + k) & 1) * 128, tbase + 256 + k * 8, ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
          else if (mode == 4)  // SS, alternating accumulators
            umma_bf16(tbase + ((it * 4 + k) & 1) * 128, ((uint64_t)DESC_HI << 32) | (a_lo + k * 2),
                      ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
          else
            umma_bf16(tbase, ((uint64_t)DESC_HI << 32) | (a_lo + k * 2), ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
        }
      }
      if (mode >= 5) {
        // the conv kernel's issue pattern: per "row" 9 taps x 4 K-steps into one of two accumulators, first MMA overwrites,
        // B walks the 72 KB weight array (8 KB per tap), A walks a 4-slot ring of 96 columns; one commit per row.
        // mode 5: commit to a second barrier every row; mode 6: no per-row commit; mode 7: as 5 with A always at slot 0
        for (int it = 0; it < iters; ++it) {
          const uint32_t d = tbase + (it & 1) * 64;
          uint32_t accumulate = 0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t a_t = tbase + 128 + (mode == 7 ? 0u : (uint32_t)((it + dy) & 3) * 96u);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t bl = b_lo + (uint32_t)((((dy * 3 + dx) * 8192) % 32768 + k * 32) >> 4);
                umma_bf16_ts(d, a_t + dx * 32 + k * 8, ((uint64_t)DESC_HI << 32) | bl, idesc, accumulate);
                accumulate = 1;
              }
          }
          if (mode != 6) umma_commit(bar + 8);  // a dummy barrier nobody waits on (initialised below)
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if (elect_one()) cycles[blockIdx.x] = t0 ? (t1 - t0) : 0;
    // only the elected lane took t0
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}
}  // namespace psgla

namespace psgla {
// The same probe for a CTA pair: the leader issues `iters` x 4 K-steps of M256 x N x K16 MMAs (cta_group::2) on zeroed
// operands, A from tensor memory (mode 0) or shared memory (mode 1), B split between the two shared memories.
template <int mode>
__global__ void __launch_bounds__(128, 1) mma_rate2_kernel(int n, int iters, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + 18 * 1024;     // up to 128 rows x 128 B (this CTA's half of N)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 52 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 16);
  for (int i = threadIdx.x; i < 52 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(tptr, 512);
    tmem_relinquish2();
  }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tbase = *tptr;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(256, n);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
      const uint32_t a_lo = (smem_u32(sa) >> 4) | 0x10000u;
      const uint32_t b_lo = (smem_u32(sb) >> 4) | 0x10000u;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 0)
            umma_bf16_ts2(tbase, tbase + 256 + k * 8, ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
          else
            umma_bf16_ss2(tbase, ((uint64_t)DESC_HI << 32) | (a_lo + k * 2), ((uint64_t)DESC_HI << 32) | (b_lo + k * 2), idesc, 1);
        }
      }
      umma_commit2(bar, 3);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if (rank == 0 && elect_one()) cycles[blockIdx.x >> 1] = t0 ? (t1 - t0) : 0;
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tbase, 512);
  }
}
}  // namespace psgla

extern "C" int psgla_selftest_mma_rate2(int mode, int n, int iters, int n_pairs, long long* cycles_dev, void* stream) {
  PSGLA_REQUIRE(cycles_dev && (mode == 0 || mode == 1) && n >= 32 && n <= 256 && n % 32 == 0 && iters > 0 && n_pairs > 0,
                "psgla_selftest_mma_rate2: bad argument");
  const int smem = 54 * 1024;
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * n_pairs));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (mode == 0)
    PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, mma_rate2_kernel<0>, n, iters, cycles_dev));
  else
    PSGLA_CUDA_TRY(cudaLaunchKernelEx(&cfg, mma_rate2_kernel<1>, n, iters, cycles_dev));
  return PSGLA_OK;
}

extern "C" int psgla_selftest_mma_rate(int mode, int n, int iters, int grid, long long* cycles_dev, void* stream) {
  PSGLA_REQUIRE(cycles_dev && mode >= 0 && mode <= 7 && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && grid > 0,
                "psgla_selftest_mma_rate: bad argument");
  PSGLA_REQUIRE(mode < 3 || n <= 128, "alternating-accumulator modes need n <= 128");
  const int smem = 54 * 1024;
  static std::atomic<unsigned long long> attr_done{0};  // bit d: opted in on device d (a per-device function attribute)
  const unsigned long long dev_bit = 1ull << (current_device() & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PSGLA_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: mma_rate_kernel<0><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 1: mma_rate_kernel<1><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 2: mma_rate_kernel<2><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 3: mma_rate_kernel<3><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 5: mma_rate_kernel<5><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 6: mma_rate_kernel<6><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    case 7: mma_rate_kernel<7><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
    default: mma_rate_kernel<4><<<grid, 128, smem, st>>>(n, iters, cycles_dev); break;
  }
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}
