// Measured FP32 issue peak: the denominator of the 2D chain kernel's roofline (MEASURED_PEAKS.json, written by the driver,
// holds a copy bandwidth and a cuBLAS bf16 figure but no FP32 one).  Every thread runs `iters` rounds of 8 independent
// dependent-FMA chains -- nothing but FFMA (mode 0) or the packed FFMA2 (mode 1) issues in the loop.
#include "common.cuh"

namespace psgla {

template <int MODE>
__global__ void __launch_bounds__(256) fp32_rate_kernel(int iters, float seed, float* __restrict__ out) {
  const float a = 1.0f + seed * 1e-7f, b = seed * 1e-3f;
  if (MODE == 0) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = seed + (float)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a, b);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the loop alive, never true in practice
  } else {
    unsigned long long v[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = seed + (float)(threadIdx.x + j);
      asm("mov.b64 %0, {%1, %2};" : "=l"(v[j]) : "f"(x), "f"(x + 0.5f));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[j]));
      s += lo + hi;
    }
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}


// Issue-rate probe of single instruction classes (what bounds the chain kernel's Philox + Box-Muller + sigmoid mix):
// every thread runs iters x 32 instructions of ONE class in 8 independent dependency chains.
//   0 IMAD.WIDE.U32 (mul.wide.u32, the Philox round multiply)   1 IMAD.HI.U32 (mul.hi.u32)   2 IMAD (mul.lo.u32)
//   3 LOP3 (3-input xor)   4 MUFU.EX2   5 I2FP.F32.U32   6 FFMA   7 MUFU.SIN   8 IMAD.WIDE + FFMA interleaved 1:1
//   9 MUFU.EX2 + IMAD.WIDE 1:1   10 MUFU.EX2 + 4 FFMA
//   11 / 100 + NF: the chain kernel's per-step MIX as independent chains (6 MUFU, 9 IMAD.WIDE, NF FFMA, 10 LOP3, 2 I2FP per
//      "step", 4 steps per loop round; NF = 19 for mode 11, else 21 / 23 / 26 / 28 / 30 = the FP32 pipe slots per step of the six
//      (algorithm, constants-structure) specialisations of gmm2d_lean_kernel): what the SM sustains for this instruction mix
//      when no instruction waits for another -- the measured ceiling the chain kernel is held against
template <int MODE, int NF = 19>
__global__ void __launch_bounds__(256) pipe_rate_kernel(int iters, unsigned seed, unsigned* __restrict__ out) {
  unsigned v[8];
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = seed * 2654435761u + threadIdx.x * 40503u + j;
    f[j] = (float)(v[j] & 1023) * 1e-3f;
  }
  const unsigned m = 0xD2511F53u + seed;
  const float a = 1.0f + seed * 1e-7f, b = seed * 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) {
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(v[j]), "r"(m));
          v[j] = (unsigned)(p >> 32) + (unsigned)p;  // IADD on the ALU pipe keeps both halves live
        } else if (MODE == 1) {
          asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(m));
        } else if (MODE == 2) {
          asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(v[j]) : "r"(m));
        } else if (MODE == 3) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j]) : "r"(m), "r"(v[(j + 1) & 7]));
        } else if (MODE == 4) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
        } else if (MODE == 5) {
          asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[j]) : "r"(v[j]));
          v[j] = __float_as_uint(f[j]);
        } else if (MODE == 6) {
          f[j] = fmaf(f[j], a, b);
        } else if (MODE == 7) {
          asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
        } else if (MODE == 8) {
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(v[j]), "r"(m));
          v[j] = (unsigned)(p >> 32) ^ (unsigned)p;
          f[j] = fmaf(f[j], a, b);
        } else if (MODE == 9) {
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(v[j]), "r"(m));
          v[j] = (unsigned)(p >> 32) ^ (unsigned)p;
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
        } else if (MODE == 10) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
          float g = f[(j + 1) & 7];
          g = fmaf(g, a, b), g = fmaf(g, a, b), g = fmaf(g, a, b), g = fmaf(g, a, b);
          f[(j + 1) & 7] = g;
        }
      }
    if (MODE == 11) {
      float e[6];  // MUFU chains
      unsigned w[3];
#pragma unroll
      for (int q = 0; q < 6; ++q) e[q] = f[q];
#pragma unroll
      for (int q = 0; q < 3; ++q) w[q] = v[q];
#pragma unroll
      for (int st = 0; st < 4; ++st) {  // 4 "steps"
#pragma unroll
        for (int q = 0; q < 6; ++q) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[q]));
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(w[q % 3]), "r"(m));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(w[q % 3]) : "r"((unsigned)(p >> 32)), "r"((unsigned)p), "r"(m));
        }
#pragma unroll
        for (int q = 0; q < NF; ++q) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[6 + (q & 1)]) : "f"(a), "f"(b));
        float c0, c1;
        asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(c0) : "r"(w[0]));
        asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(c1) : "r"(w[1]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[7]) : "r"(__float_as_uint(c0)), "r"(__float_as_uint(c1)));  // the 10th LOP3
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) f[q] = e[q];
#pragma unroll
      for (int q = 0; q < 3; ++q) v[q] = w[q];
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j] + __float_as_uint(f[j]);
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace psgla

using namespace psgla;

extern "C" int psgla_selftest_fp32_rate(int mode, int iters, int blocks_per_sm, float* out_dev, double* flop_out,
                                        void* stream) {
  PSGLA_REQUIRE((mode == 0 || mode == 1) && iters > 0 && blocks_per_sm > 0 && blocks_per_sm <= 8 && out_dev != nullptr,
                "psgla_selftest_fp32_rate: bad argument");
  const int grid = num_sms() * blocks_per_sm;
  if (mode == 0)
    fp32_rate_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, 0.25f, out_dev);
  else
    fp32_rate_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, 0.25f, out_dev);
  PSGLA_CUDA_TRY(cudaGetLastError());
  if (flop_out) *flop_out = (double)grid * 256.0 * (double)iters * 32.0 * 2.0 * (mode == 1 ? 2.0 : 1.0);
  return PSGLA_OK;
}

extern "C" int psgla_selftest_pipe_rate(int mode, int iters, int blocks_per_sm, void* out_dev, double* ops_out, void* stream) {
  PSGLA_REQUIRE(((mode >= 0 && mode <= 11) || mode == 121 || mode == 123 || mode == 126 || mode == 128 || mode == 130) && iters > 0 && blocks_per_sm > 0 && blocks_per_sm <= 8 && out_dev != nullptr,
                "psgla_selftest_pipe_rate: bad argument");
  const int grid = num_sms() * blocks_per_sm;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* o = (unsigned*)out_dev;
  switch (mode) {
    case 0: pipe_rate_kernel<0><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 1: pipe_rate_kernel<1><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 2: pipe_rate_kernel<2><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 3: pipe_rate_kernel<3><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 4: pipe_rate_kernel<4><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 5: pipe_rate_kernel<5><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 6: pipe_rate_kernel<6><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 7: pipe_rate_kernel<7><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 8: pipe_rate_kernel<8><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 9: pipe_rate_kernel<9><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 10: pipe_rate_kernel<10><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 121: pipe_rate_kernel<11, 21><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 123: pipe_rate_kernel<11, 23><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 126: pipe_rate_kernel<11, 26><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 128: pipe_rate_kernel<11, 28><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    case 130: pipe_rate_kernel<11, 30><<<grid, 256, 0, st>>>(iters, 3u, o); break;
    default: pipe_rate_kernel<11><<<grid, 256, 0, st>>>(iters, 3u, o); break;
  }
  PSGLA_CUDA_TRY(cudaGetLastError());
  // thread-level instructions of the probed class; mode 11: thread-level "steps" (4 per loop round)
  if (ops_out) *ops_out = (double)grid * 256.0 * (double)iters * (mode >= 11 ? 4.0 : 32.0);
  return PSGLA_OK;
}
