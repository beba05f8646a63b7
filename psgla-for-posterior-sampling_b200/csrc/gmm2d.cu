// 2D Gaussian-mixture posterior sampling: PSGLA (SnoPnP_ULA) and PnP-ULA as one register-resident kernel.
//
// Replaces the Python loops sampling_2D.py:21-45 / :48-72, the data-fidelity score sampling_2D.py:30-31 and the
// closed-form MMSE denoiser utils_2D.py:209-233 of the reference.  One thread owns CPT chains for all n_steps:
// state, Philox counter and step constants live in registers / the constant bank; HBM sees x_0 once, the final
// state once, and (optionally) a thinned trajectory.  Bound: FP32 + MUFU issue, not memory ("2D kernel").
//
// Algebra (all folding done on the host in double, see fold_problem):
//   score(x) = A^T (y - A x) / sigma^2 = bb - G x
//   tau      = sqrt(eps)           (the reference feeds sqrt(epsilon) where a variance belongs, utils_2D.py:223-226;
//                                   PSGLA calls the denoiser with eps = delta, sampling_2D.py:63)
//   D(v)     = sum_i w_i(v) (M_i v + b_i),  w = softmax_i( kappa_i - 1/2 (v-mu_i)^T S_i (v-mu_i) )
//              S_i = (tau I + Sigma_i)^-1, kappa_i = log pi_i - 1/2 log det(tau I + Sigma_i),
//              M_i = (I/tau + Sigma_i^-1)^-1 / tau, b_i = (I/tau + Sigma_i^-1)^-1 Sigma_i^-1 mu_i
//   PSGLA  : x+ = D(P x + q + cn z),            P = I - (delta/alpha) G, q = (delta/alpha) bb, cn = sqrt(2 delta)
//   PnP-ULA: x+ = P x + q + cn z + cp D(x),     P = (1 - cp) I - delta G, q = delta bb, cp = alpha delta / eps
// The reference evaluates the softmax with plain exp() and can hit 0/0 far from every mode; here the logits are
// carried in log2 units and normalised (r == 2: a single sigmoid of the logit *difference*, which is itself a
// quadratic in v; r > 2: online log-sum-exp over the components held in shared memory).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace psgla {

constexpr int RMAX = PSGLA_GMM_MAX_COMPONENTS;

// default launch geometry of the fp32, r == 2 Philox path: cpt, nb, block, pack, dynamic (see launch_run)
#ifndef PSGLA_GMM_DEFAULT_GEOM
#define PSGLA_GMM_DEFAULT_GEOM 4, 64, 16, 1, 2
#endif

template <typename T>
struct Gmm2dConsts {
  T P[4], q[2];
  T cn, cp;
  // r == 2 fast path: t(v) = l_1(v) - l_0(v) in log2 units, D = m_1 + sigmoid * (m_0 - m_1)
  T tq[6];
  T M1[4], b1[2], dM[4], db[2];
  int r;
};

template <typename T>
struct Gmm2dComponents {  // general-r path, staged into shared memory
  T mu[RMAX][2];
  T s[RMAX][3];  // 1/2 log2(e) S00, log2(e) S01, 1/2 log2(e) S11
  T kappa2[RMAX];
  T M[RMAX][4];
  T b[RMAX][2];
};

// ------------------------------------------------------------------------------------------------ math helpers
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_exp2(double x) { return exp2(x); }
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
  using type = float2;
};
template <>
struct Vec2<double> {
  using type = double2;
};

// D(v) for r == 2: everything comes from the constant bank as FFMA operands.
template <typename T>
__device__ __forceinline__ void denoise_r2(const Gmm2dConsts<T>& c, T v0, T v1, T& d0, T& d1) {
  const T t = fma_(v0, fma_(c.tq[3], v0, fma_(c.tq[4], v1, c.tq[1])), fma_(v1, fma_(c.tq[5], v1, c.tq[2]), c.tq[0]));
  const T w0 = fast_rcp(T(1) + fast_exp2(t));  // weight of component 0
  const T m0 = fma_(c.M1[0], v0, fma_(c.M1[1], v1, c.b1[0]));
  const T m1 = fma_(c.M1[2], v0, fma_(c.M1[3], v1, c.b1[1]));
  const T e0 = fma_(c.dM[0], v0, fma_(c.dM[1], v1, c.db[0]));
  const T e1 = fma_(c.dM[2], v0, fma_(c.dM[3], v1, c.db[1]));
  d0 = fma_(w0, e0, m0);
  d1 = fma_(w0, e1, m1);
}

// D(v) for general r: online log-sum-exp over components in shared memory (broadcast LDS).
template <typename T>
__device__ __forceinline__ void denoise_general(const Gmm2dComponents<T>* __restrict__ k, int r, T v0, T v1, T& d0,
                                                T& d1) {
  T mx = -INFINITY, den = 0, a0 = 0, a1 = 0;
  for (int i = 0; i < r; ++i) {
    const T e0 = v0 - k->mu[i][0], e1 = v1 - k->mu[i][1];
    const T l = k->kappa2[i] - fma_(k->s[i][0] * e0, e0, fma_(k->s[i][1] * e0, e1, k->s[i][2] * e1 * e1));
    const T m0 = fma_(k->M[i][0], v0, fma_(k->M[i][1], v1, k->b[i][0]));
    const T m1 = fma_(k->M[i][2], v0, fma_(k->M[i][3], v1, k->b[i][1]));
    const T nmx = l > mx ? l : mx;
    const T scale = fast_exp2(mx - nmx);  // exp2(-inf) = 0 on the first component
    const T w = fast_exp2(l - nmx);
    den = fma_(den, scale, w);
    a0 = fma_(a0, scale, w * m0);
    a1 = fma_(a1, scale, w * m1);
    mx = nmx;
  }
  const T inv = T(1) / den;
  d0 = a0 * inv;
  d1 = a1 * inv;
}

template <typename T, int ALG, bool R2>
__device__ __forceinline__ void langevin_step(const Gmm2dConsts<T>& c, const Gmm2dComponents<T>* __restrict__ comps,
                                              T& x0, T& x1, T z0, T z1) {
  const T l0 = fma_(c.P[0], x0, fma_(c.P[1], x1, fma_(c.cn, z0, c.q[0])));
  const T l1 = fma_(c.P[2], x0, fma_(c.P[3], x1, fma_(c.cn, z1, c.q[1])));
  const T v0 = (ALG == PSGLA_ALG_PSGLA) ? l0 : x0;
  const T v1 = (ALG == PSGLA_ALG_PSGLA) ? l1 : x1;
  T d0, d1;
  if (R2)
    denoise_r2(c, v0, v1, d0, d1);
  else
    denoise_general(comps, c.r, v0, v1, d0, d1);
  if (ALG == PSGLA_ALG_PSGLA) {
    x0 = d0;
    x1 = d1;
  } else {
    x0 = fma_(c.cp, d0, l0);
    x1 = fma_(c.cp, d1, l1);
  }
}

// ------------------------------------------------------------------------------------------------ packed fp32 pairs
// sm_100 issues two fp32 FMAs per lane as ONE instruction (FFMA2, PTX fma.rn.f32x2).  The chain kernel is issue-bound, so
// the r == 2 fp32 path carries chains in PAIRS: every state / noise / intermediate register pair holds the same quantity of
// chains 2p and 2p+1, the constants are duplicated into both halves on the host (Gmm2dConstsPk), and the affine maps, the
// quadratic logit, the component means and the Box-Muller arithmetic cost one instruction per pair instead of two.  The
// operations and their order per chain are those of langevin_step<float>: results are bit-identical to the scalar path.
typedef unsigned long long pk2;  // {lo: chain 2p, hi: chain 2p+1}
__device__ __forceinline__ pk2 pk(float lo, float hi) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 fma2(pk2 a, pk2 b, pk2 c) {
  pk2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ pk2 mul2(pk2 a, pk2 b) {
  pk2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk2 add2(pk2 a, pk2 b) {
  pk2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

struct Gmm2dConstsPk {  // Gmm2dConsts<float> with every constant duplicated into both halves of a 64-bit word
  pk2 P[4], q[2], cn, cp, tq[6], M1[4], b1[2], dM[4], db[2], one;
};

__device__ __forceinline__ void denoise_r2_pk(const Gmm2dConstsPk& c, pk2 v0, pk2 v1, pk2& d0, pk2& d1) {
  const pk2 t = fma2(v0, fma2(c.tq[3], v0, fma2(c.tq[4], v1, c.tq[1])), fma2(v1, fma2(c.tq[5], v1, c.tq[2]), c.tq[0]));
  float ta, tb;
  unpk(t, ta, tb);
  float sa, sb;
  unpk(add2(c.one, pk(fast_exp2(ta), fast_exp2(tb))), sa, sb);
  const pk2 w0 = pk(fast_rcp(sa), fast_rcp(sb));
  const pk2 m0 = fma2(c.M1[0], v0, fma2(c.M1[1], v1, c.b1[0]));
  const pk2 m1 = fma2(c.M1[2], v0, fma2(c.M1[3], v1, c.b1[1]));
  const pk2 e0 = fma2(c.dM[0], v0, fma2(c.dM[1], v1, c.db[0]));
  const pk2 e1 = fma2(c.dM[2], v0, fma2(c.dM[3], v1, c.db[1]));
  d0 = fma2(w0, e0, m0);
  d1 = fma2(w0, e1, m1);
}

template <int ALG>
__device__ __forceinline__ void langevin_step_pk(const Gmm2dConstsPk& c, pk2& x0, pk2& x1, pk2 z0, pk2 z1) {
  const pk2 l0 = fma2(c.P[0], x0, fma2(c.P[1], x1, fma2(c.cn, z0, c.q[0])));
  const pk2 l1 = fma2(c.P[2], x0, fma2(c.P[3], x1, fma2(c.cn, z1, c.q[1])));
  pk2 d0, d1;
  if (ALG == PSGLA_ALG_PSGLA) {
    denoise_r2_pk(c, l0, l1, d0, d1);
    x0 = d0;
    x1 = d1;
  } else {
    denoise_r2_pk(c, x0, x1, d0, d1);
    x0 = fma2(c.cp, d0, l0);
    x1 = fma2(c.cp, d1, l1);
  }
}

// box_muller (common.cuh) on the same word of two chains at once: (a, b) of chain 2p and of chain 2p+1.
__device__ __forceinline__ void box_muller_pk(uint32_t aA, uint32_t bA, uint32_t aB, uint32_t bB, pk2& z0, pk2& z1) {
  const pk2 u1 = fma2(pk((float)aA, (float)aB), pk(2.3283064365386963e-10f, 2.3283064365386963e-10f),
                      pk(1.1641532182693481e-10f, 1.1641532182693481e-10f));
  const pk2 th = fma2(pk((float)bA, (float)bB), pk(1.4629180792671596e-9f, 1.4629180792671596e-9f),
                      pk(7.314590396335798e-10f, 7.314590396335798e-10f));
  float ua, ub, la, lb, ra, rb, ta, tb;
  unpk(u1, ua, ub);
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(ua));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(ub));
  unpk(mul2(pk(la, lb), pk(-1.3862943611198906f, -1.3862943611198906f)), ua, ub);
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(ua));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(ub));
  unpk(th, ta, tb);
  float sa, ca, sb, cb;
  __sincosf(ta, &sa, &ca);
  __sincosf(tb, &sb, &cb);
  const pk2 r = pk(ra, rb);
  z0 = mul2(r, pk(ca, cb));
  z1 = mul2(r, pk(sa, sb));
}

// ------------------------------------------------------------------------------------------------ the kernel
// Chain j of thread g is chain index g + j * (gridDim.x * blockDim.x): every global access is coalesced.
// Philox mode draws NB counter blocks (= 2 NB steps) of every owned chain at once: the 10-round integer pipeline of the NB x CPT
// independent blocks is where the instruction-level parallelism comes from, so a thread that owns ONE chain (small populations,
// and the 1-warp blocks of the dynamically scheduled launch) still keeps the pipes busy between the dependent Langevin steps.
template <typename T, int ALG, bool R2, int CPT, int NB, int BLOCK, bool PACK>
__global__ void __launch_bounds__(BLOCK, (sizeof(T) == 4 && R2) ? 1024 / BLOCK : 1)
gmm2d_kernel(const Gmm2dConsts<T> c, const Gmm2dConstsPk cpk, const Gmm2dComponents<T>* __restrict__ comps_gmem,
             T* __restrict__ x, long long n_chains, long long chain_lo, long long n_launch, unsigned long long chain_id0,
             long long n_steps, long long step0, const PhiloxKeys keys, const T* __restrict__ noise, T* __restrict__ traj,
             long long thin) {
  // This launch owns chains [chain_lo, chain_lo + n_launch) of the call's n_chains (the stride of noise / traj rows).
  using V2 = typename Vec2<T>::type;
  static_assert(!PACK || (sizeof(T) == 4 && R2 && CPT % 2 == 0), "the packed path is fp32, r == 2, chains in pairs");
  __shared__ Gmm2dComponents<T> comps_smem;
  const Gmm2dComponents<T>* comps = nullptr;
  if (!R2) {
    const int nwords = sizeof(Gmm2dComponents<T>) / 4;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x)
      reinterpret_cast<uint32_t*>(&comps_smem)[i] = reinterpret_cast<const uint32_t*>(comps_gmem)[i];
    __syncthreads();
    comps = &comps_smem;
  }
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  T x0[CPT], x1[CPT];
  bool live[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const long long ch = chain_lo + g + j * nthreads;
    live[j] = g + j * nthreads < n_launch;
    V2 v = live[j] ? reinterpret_cast<const V2*>(x)[ch] : V2{0, 0};
    x0[j] = v.x;
    x1[j] = v.y;
  }

  long long keep_in = thin;  // steps until the next trajectory row
  long long row = 0;
  auto after_step = [&](void) {
    if (traj != nullptr) {
      if (--keep_in == 0) {
        keep_in = thin;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const long long ch = chain_lo + g + j * nthreads;
          if (live[j]) reinterpret_cast<V2*>(traj)[row * n_chains + ch] = V2{x0[j], x1[j]};
        }
        ++row;
      }
    }
  };

  if (noise != nullptr) {
    // replay: the caller's N(0,1) draws (the reference's np.random.randn(2) per step, sampling_2D.py:35,62)
    for (long long k = 0; k < n_steps; ++k) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const long long ch = chain_lo + g + j * nthreads;
        if (live[j]) {
          const V2 z = reinterpret_cast<const V2*>(noise)[k * n_chains + ch];
          langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], z.x, z.y);
        }
      }
      after_step();
    }
  } else {
    long long t = step0;
    const long long t_end = step0 + n_steps;
    unsigned long long sub[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) sub[j] = chain_id0 + (unsigned long long)(chain_lo + g + j * nthreads);
    // one counter block = the normals of steps 2 * pair and 2 * pair + 1; `first` / `second`: which of the two are taken
    auto one_block = [&](bool first, bool second) {
      float z[CPT][4];
      const unsigned long long pair = (unsigned long long)t >> 1;
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        philox_normal4_keyed(keys, sub[j], (uint32_t)pair, (uint32_t)(pair >> 32), z[j][0], z[j][1], z[j][2], z[j][3]);
      if (first) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], T(z[j][0]), T(z[j][1]));
        after_step();
        ++t;
      }
      if (second) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], T(z[j][2]), T(z[j][3]));
        after_step();
        ++t;
      }
    };
    if (t < t_end && (t & 1)) one_block(false, true);  // a segment that starts on the second half of a counter block
    if constexpr (PACK) {
      pk2 X0[CPT / 2], X1[CPT / 2];
#pragma unroll
      for (int p = 0; p < CPT / 2; ++p) {
        X0[p] = pk(x0[2 * p], x0[2 * p + 1]);
        X1[p] = pk(x1[2 * p], x1[2 * p + 1]);
      }
      auto sync_scalars = [&](void) {
#pragma unroll
        for (int p = 0; p < CPT / 2; ++p) {
          unpk(X0[p], x0[2 * p], x0[2 * p + 1]);
          unpk(X1[p], x1[2 * p], x1[2 * p + 1]);
        }
      };
      while (t + 2 * NB <= t_end) {
        pk2 z[NB][CPT / 2][4];
        const unsigned long long pair = (unsigned long long)t >> 1;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int p = 0; p < CPT / 2; ++p) {
            const unsigned long long pr = pair + b;
            uint32_t a0 = (uint32_t)pr, a1 = (uint32_t)(pr >> 32), a2 = (uint32_t)sub[2 * p], a3 = (uint32_t)(sub[2 * p] >> 32);
            uint32_t b0 = a0, b1 = a1, b2 = (uint32_t)sub[2 * p + 1], b3 = (uint32_t)(sub[2 * p + 1] >> 32);
            philox4x32_10_keyed(a0, a1, a2, a3, keys);
            philox4x32_10_keyed(b0, b1, b2, b3, keys);
            box_muller_pk(a0, a1, b0, b1, z[b][p][0], z[b][p][1]);
            box_muller_pk(a2, a3, b2, b3, z[b][p][2], z[b][p][3]);
          }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int p = 0; p < CPT / 2; ++p) langevin_step_pk<ALG>(cpk, X0[p], X1[p], z[b][p][2 * h], z[b][p][2 * h + 1]);
            if (traj != nullptr) {
              sync_scalars();
              after_step();
            }
          }
        }
        t += 2 * NB;
      }
      sync_scalars();
    } else {
      while (t + 2 * NB <= t_end) {
        float z[NB][CPT][4];
        const unsigned long long pair = (unsigned long long)t >> 1;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const unsigned long long pr = pair + b;
            philox_normal4_keyed(keys, sub[j], (uint32_t)pr, (uint32_t)(pr >> 32), z[b][j][0], z[b][j][1], z[b][j][2], z[b][j][3]);
          }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < CPT; ++j)
              langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], T(z[b][j][2 * h]), T(z[b][j][2 * h + 1]));
            after_step();
          }
        }
        t += 2 * NB;
      }
    }
    while (t < t_end) one_block(true, t + 1 < t_end);  // fewer than 2 NB steps left
  }

#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const long long ch = chain_lo + g + j * nthreads;
    if (live[j]) reinterpret_cast<V2*>(x)[ch] = V2{x0[j], x1[j]};
  }
}

// Structure of the folded constants.  The kernel is bound by the FMA pipe (scripts/pipe_rates.py: an FFMA costs the pipe 1
// cycle per warp, each of Philox's ten IMAD.WIDE.U32 pairs 4), so FMAs whose coefficient is exactly zero are worth removing:
//   STRUCT 1: P diagonal (A^T A diagonal -- every cell of the reference's experiment has A = I, sampling_2D.py:84);
//   STRUCT 2: additionally M_i, M_0 - M_1 diagonal and no v0 v1 term in the logit (isotropic / axis-aligned covariances: the
//             "symetric_gaussians" and "disymmetric_gaussians" priors, utils_2D.py:28-33): 15 FP32 instructions per step
//             instead of 24.
// fma(0, a, b) == b for finite a, so dropping those terms leaves every result bit-identical to the general path.
template <int ALG, int STRUCT>
__device__ __forceinline__ void langevin_step_s(const Gmm2dConsts<float>& c, float& x0, float& x1, float z0, float z1) {
  float l0, l1;
  if (STRUCT >= 1) {
    l0 = fmaf(c.P[0], x0, fmaf(c.cn, z0, c.q[0]));
    l1 = fmaf(c.P[3], x1, fmaf(c.cn, z1, c.q[1]));
  } else {
    l0 = fmaf(c.P[0], x0, fmaf(c.P[1], x1, fmaf(c.cn, z0, c.q[0])));
    l1 = fmaf(c.P[2], x0, fmaf(c.P[3], x1, fmaf(c.cn, z1, c.q[1])));
  }
  const float v0 = (ALG == PSGLA_ALG_PSGLA) ? l0 : x0;
  const float v1 = (ALG == PSGLA_ALG_PSGLA) ? l1 : x1;
  float d0, d1;
  if (STRUCT >= 2) {
    const float t = fmaf(v0, fmaf(c.tq[3], v0, c.tq[1]), fmaf(v1, fmaf(c.tq[5], v1, c.tq[2]), c.tq[0]));
    const float w0 = fast_rcp(1.0f + fast_exp2(t));
    const float m0 = fmaf(c.M1[0], v0, c.b1[0]);
    const float m1 = fmaf(c.M1[3], v1, c.b1[1]);
    const float e0 = fmaf(c.dM[0], v0, c.db[0]);
    const float e1 = fmaf(c.dM[3], v1, c.db[1]);
    d0 = fmaf(w0, e0, m0);
    d1 = fmaf(w0, e1, m1);
  } else {
    denoise_r2(c, v0, v1, d0, d1);
  }
  if (ALG == PSGLA_ALG_PSGLA) {
    x0 = d0;
    x1 = d1;
  } else {
    x0 = fmaf(c.cp, d0, l0);
    x1 = fmaf(c.cp, d1, l1);
  }
}

static int constants_structure(const Gmm2dConsts<float>& c) {
  if (c.P[1] != 0.0f || c.P[2] != 0.0f) return 0;
  if (c.M1[1] != 0.0f || c.M1[2] != 0.0f || c.dM[1] != 0.0f || c.dM[2] != 0.0f || c.tq[4] != 0.0f) return 1;
  return 2;
}

// ------------------------------------------------------------------------------------------------ the lean kernel
// fp32, r == 2, in-kernel Philox, no trajectory: the throughput path and nothing else, so that it fits a register budget
// that keeps more warps resident than the general kernel's 64 registers allow (the general kernel spends ~20 registers on
// 64-bit bookkeeping of the optional replay / trajectory paths).  One chain per thread; each round draws NB counter blocks
// (2 NB steps), with the Box-Muller arithmetic of two blocks packed into FFMA2 / FMUL2 when PK.  Same Philox counters, same
// arithmetic per chain as gmm2d_kernel: bit-identical results.
// MODE bit 0: Box-Muller arithmetic of two counter blocks packed (FFMA2 / FMUL2); bit 1: software pipelining -- the Philox
// words of round r + 1 are computed in the same loop body as the Box-Muller transform and the Langevin steps of round r, so
// that every warp offers the scheduler IMAD.WIDE (FMA pipe), MUFU (XU pipe) and FFMA work at the same time instead of in
// three phases.
template <int ALG, int STRUCT, int NB, int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB)
gmm2d_lean_kernel(const Gmm2dConsts<float> c, float2* __restrict__ x, unsigned n_launch, unsigned long long sub0,
                  unsigned long long pair0, int n_rounds, const PhiloxKeys keys) {
  constexpr bool PK = (MODE & 1) != 0 && NB % 2 == 0, PIPE = (MODE & 2) != 0;
  const unsigned g = blockIdx.x * BLOCK + threadIdx.x;
  if (g >= n_launch) return;
  float2 v = x[g];
  float x0 = v.x, x1 = v.y;
  const unsigned long long sub = sub0 + g;
  const uint32_t s_lo = (uint32_t)sub, s_hi = (uint32_t)(sub >> 32);
  unsigned long long pair = pair0;
  auto words = [&](unsigned long long pr0, uint32_t (&w)[NB][4]) {  // Philox4x32-10 of counter blocks pr0 .. pr0 + NB - 1
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const unsigned long long pr = pr0 + b;
      w[b][0] = (uint32_t)pr, w[b][1] = (uint32_t)(pr >> 32), w[b][2] = s_lo, w[b][3] = s_hi;
      philox4x32_10_keyed(w[b][0], w[b][1], w[b][2], w[b][3], keys);
    }
  };
  auto normals = [&](const uint32_t (&w)[NB][4], float (&z)[NB][4]) {
    if constexpr (PK) {
#pragma unroll
      for (int b = 0; b < NB; b += 2) {
        pk2 q0, q1, q2, q3;
        box_muller_pk(w[b][0], w[b][1], w[b + 1][0], w[b + 1][1], q0, q1);
        box_muller_pk(w[b][2], w[b][3], w[b + 1][2], w[b + 1][3], q2, q3);
        unpk(q0, z[b][0], z[b + 1][0]);
        unpk(q1, z[b][1], z[b + 1][1]);
        unpk(q2, z[b][2], z[b + 1][2]);
        unpk(q3, z[b][3], z[b + 1][3]);
      }
    } else {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        box_muller(w[b][0], w[b][1], z[b][0], z[b][1]);
        box_muller(w[b][2], w[b][3], z[b][2], z[b][3]);
      }
    }
  };
  uint32_t w[NB][4];
  if (PIPE) words(pair, w);
  for (int r = 0; r < n_rounds; ++r, pair += NB) {
    float z[NB][4];
    if (PIPE) {
      uint32_t wn[NB][4];
      words(pair + NB, wn);  // one round ahead (the last iteration's are unused)
      normals(w, z);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        langevin_step_s<ALG, STRUCT>(c, x0, x1, z[b][0], z[b][1]);
        langevin_step_s<ALG, STRUCT>(c, x0, x1, z[b][2], z[b][3]);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) w[b][i] = wn[b][i];
    } else {
      words(pair, w);
      normals(w, z);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        langevin_step_s<ALG, STRUCT>(c, x0, x1, z[b][0], z[b][1]);
        langevin_step_s<ALG, STRUCT>(c, x0, x1, z[b][2], z[b][3]);
      }
    }
  }
  x[g] = float2{x0, x1};
}

template <typename T, bool R2>
__global__ void gmm2d_denoise_kernel(const Gmm2dConsts<T> c, const Gmm2dComponents<T>* __restrict__ comps_gmem,
                                     const T* __restrict__ x, T* __restrict__ out, long long n) {
  using V2 = typename Vec2<T>::type;
  __shared__ Gmm2dComponents<T> comps_smem;
  const int nwords = sizeof(Gmm2dComponents<T>) / 4;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&comps_smem)[i] = reinterpret_cast<const uint32_t*>(comps_gmem)[i];
  __syncthreads();
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const V2 v = reinterpret_cast<const V2*>(x)[g];
  T d0, d1;
  if (R2)
    denoise_r2(c, v.x, v.y, d0, d1);
  else
    denoise_general(&comps_smem, c.r, v.x, v.y, d0, d1);
  reinterpret_cast<V2*>(out)[g] = V2{d0, d1};
}

__global__ void gmm2d_noise_kernel(float* __restrict__ out, long long n_chains, unsigned long long chain_id0,
                                   long long n_steps, long long step0, unsigned long long seed) {
  const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= n_chains) return;
  for (long long k = 0; k < n_steps; ++k) {
    const unsigned long long t = (unsigned long long)(step0 + k), pair = t >> 1;
    float z0, z1, z2, z3;
    philox_normal4(seed, chain_id0 + (unsigned long long)ch, (uint32_t)pair, (uint32_t)(pair >> 32), z0, z1, z2, z3);
    reinterpret_cast<float2*>(out)[k * n_chains + ch] = (t & 1) ? float2{z2, z3} : float2{z0, z1};
  }
}

// ------------------------------------------------------------------------------------------------ host folding
struct Mat2 {
  double a, b, c, d;  // [[a b][c d]]
};
static inline Mat2 inv2(const Mat2& m) {
  const double det = m.a * m.d - m.b * m.c;
  return {m.d / det, -m.b / det, -m.c / det, m.a / det};
}
static inline double det2(const Mat2& m) { return m.a * m.d - m.b * m.c; }

struct Folded {
  Gmm2dConsts<double> c;
  Gmm2dComponents<double> k;
};

static int fold_problem(const psgla_gmm2d_problem* p, double eps, bool linear_part, Folded* out) {
  const int r = p->n_components;
  if (r < 1 || r > RMAX) return set_error(PSGLA_E_UNSUPPORTED, "n_components=%d outside 1..%d", r, RMAX);
  if (!(eps > 0)) return set_error(PSGLA_E_BADARG, "denoiser level must be > 0 (got %g)", eps);
  std::memset(out, 0, sizeof(*out));
  const double LOG2E = 1.4426950408889634;
  const double tau = std::sqrt(eps);
  Gmm2dConsts<double>& c = out->c;
  Gmm2dComponents<double>& k = out->k;
  c.r = r;
  double lq[RMAX][6];
  for (int i = 0; i < r; ++i) {
    const Mat2 Sig{p->Sigma[i][0], p->Sigma[i][1], p->Sigma[i][2], p->Sigma[i][3]};
    if (!(det2(Sig) > 0) || !(Sig.a > 0) || !(p->pi[i] > 0))
      return set_error(PSGLA_E_BADARG, "component %d: Sigma must be SPD and pi > 0", i);
    const Mat2 Sinv = inv2(Sig);
    const Mat2 St = inv2({tau + Sig.a, Sig.b, Sig.c, tau + Sig.d});
    const double kappa = std::log(p->pi[i]) - 0.5 * std::log(det2({tau + Sig.a, Sig.b, Sig.c, tau + Sig.d}));
    const Mat2 Pm = inv2({1.0 / tau + Sinv.a, Sinv.b, Sinv.c, 1.0 / tau + Sinv.d});
    const double m0 = p->mu[i][0], m1 = p->mu[i][1];
    const double sm0 = Sinv.a * m0 + Sinv.b * m1, sm1 = Sinv.c * m0 + Sinv.d * m1;
    k.mu[i][0] = m0;
    k.mu[i][1] = m1;
    k.s[i][0] = 0.5 * LOG2E * St.a;
    k.s[i][1] = 0.5 * LOG2E * (St.b + St.c);
    k.s[i][2] = 0.5 * LOG2E * St.d;
    k.kappa2[i] = kappa * LOG2E;
    k.M[i][0] = Pm.a / tau;
    k.M[i][1] = Pm.b / tau;
    k.M[i][2] = Pm.c / tau;
    k.M[i][3] = Pm.d / tau;
    k.b[i][0] = Pm.a * sm0 + Pm.b * sm1;
    k.b[i][1] = Pm.c * sm0 + Pm.d * sm1;
    // l_i(v) = lq0 + lq1 v0 + lq2 v1 + lq3 v0^2 + lq4 v0 v1 + lq5 v1^2
    const double s00 = k.s[i][0], s01 = k.s[i][1], s11 = k.s[i][2];
    lq[i][0] = k.kappa2[i] - (s00 * m0 * m0 + s01 * m0 * m1 + s11 * m1 * m1);
    lq[i][1] = 2 * s00 * m0 + s01 * m1;
    lq[i][2] = 2 * s11 * m1 + s01 * m0;
    lq[i][3] = -s00;
    lq[i][4] = -s01;
    lq[i][5] = -s11;
  }
  if (r == 2) {
    for (int j = 0; j < 6; ++j) c.tq[j] = lq[1][j] - lq[0][j];
    for (int j = 0; j < 4; ++j) {
      c.M1[j] = k.M[1][j];
      c.dM[j] = k.M[0][j] - k.M[1][j];
    }
    for (int j = 0; j < 2; ++j) {
      c.b1[j] = k.b[1][j];
      c.db[j] = k.b[0][j] - k.b[1][j];
    }
  }
  if (linear_part) {
    if (!(p->delta > 0) || !(p->sigma != 0) || !(p->alpha != 0))
      return set_error(PSGLA_E_BADARG, "delta must be > 0, sigma and alpha non-zero");
    const double s2 = p->sigma * p->sigma;
    const double* A = p->A;
    const Mat2 G{(A[0] * A[0] + A[2] * A[2]) / s2, (A[0] * A[1] + A[2] * A[3]) / s2, (A[1] * A[0] + A[3] * A[2]) / s2,
                 (A[1] * A[1] + A[3] * A[3]) / s2};
    const double bb0 = (A[0] * p->y[0] + A[2] * p->y[1]) / s2, bb1 = (A[1] * p->y[0] + A[3] * p->y[1]) / s2;
    c.cn = std::sqrt(2 * p->delta);
    if (p->alg == PSGLA_ALG_PSGLA) {
      const double cs = p->delta / p->alpha;
      c.P[0] = 1 - cs * G.a;
      c.P[1] = -cs * G.b;
      c.P[2] = -cs * G.c;
      c.P[3] = 1 - cs * G.d;
      c.q[0] = cs * bb0;
      c.q[1] = cs * bb1;
      c.cp = 0;
    } else if (p->alg == PSGLA_ALG_PNPULA) {
      c.cp = p->alpha * p->delta / p->epsilon;
      c.P[0] = 1 - c.cp - p->delta * G.a;
      c.P[1] = -p->delta * G.b;
      c.P[2] = -p->delta * G.c;
      c.P[3] = 1 - c.cp - p->delta * G.d;
      c.q[0] = p->delta * bb0;
      c.q[1] = p->delta * bb1;
    } else {
      return set_error(PSGLA_E_BADARG, "alg=%d is neither PSGLA_ALG_PSGLA nor PSGLA_ALG_PNPULA", p->alg);
    }
  }
  return PSGLA_OK;
}

template <typename T>
static void narrow(const Folded& f, Gmm2dConsts<T>* c, Gmm2dComponents<T>* k) {
  const double* src = reinterpret_cast<const double*>(&f.c);
  // Gmm2dConsts<double> is all doubles followed by one int: copy field by field
  (void)src;
  for (int i = 0; i < 4; ++i) c->P[i] = (T)f.c.P[i];
  for (int i = 0; i < 2; ++i) c->q[i] = (T)f.c.q[i];
  c->cn = (T)f.c.cn;
  c->cp = (T)f.c.cp;
  for (int i = 0; i < 6; ++i) c->tq[i] = (T)f.c.tq[i];
  for (int i = 0; i < 4; ++i) {
    c->M1[i] = (T)f.c.M1[i];
    c->dM[i] = (T)f.c.dM[i];
  }
  for (int i = 0; i < 2; ++i) {
    c->b1[i] = (T)f.c.b1[i];
    c->db[i] = (T)f.c.db[i];
  }
  c->r = f.c.r;
  for (int i = 0; i < RMAX; ++i) {
    for (int j = 0; j < 2; ++j) k->mu[i][j] = (T)f.k.mu[i][j];
    for (int j = 0; j < 3; ++j) k->s[i][j] = (T)f.k.s[i][j];
    k->kappa2[i] = (T)f.k.kappa2[i];
    for (int j = 0; j < 4; ++j) k->M[i][j] = (T)f.k.M[i][j];
    for (int j = 0; j < 2; ++j) k->b[i][j] = (T)f.k.b[i][j];
  }
}

// Per-stream-ordered staging of the general-r component table: a small device buffer per call, freed stream-ordered.
template <typename T>
static int upload_components(const Gmm2dComponents<T>& k, cudaStream_t st, Gmm2dComponents<T>** dev) {
  PSGLA_CUDA_TRY(cudaMallocAsync((void**)dev, sizeof(k), st));
  PSGLA_CUDA_TRY(cudaMemcpyAsync(*dev, &k, sizeof(k), cudaMemcpyHostToDevice, st));
  return PSGLA_OK;
}

static Gmm2dConstsPk pack_consts(const Gmm2dConsts<float>& c) {
  Gmm2dConstsPk r;
  auto dup = [](float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    return (pk2)u | ((pk2)u << 32);
  };
  for (int i = 0; i < 4; ++i) r.P[i] = dup(c.P[i]), r.M1[i] = dup(c.M1[i]), r.dM[i] = dup(c.dM[i]);
  for (int i = 0; i < 2; ++i) r.q[i] = dup(c.q[i]), r.b1[i] = dup(c.b1[i]), r.db[i] = dup(c.db[i]);
  for (int i = 0; i < 6; ++i) r.tq[i] = dup(c.tq[i]);
  r.cn = dup(c.cn);
  r.cp = dup(c.cp);
  r.one = dup(1.0f);
  return r;
}

// Launch geometry.  The kernel is issue-bound and its warps do identical work, but the SM's warp arbiter is not fair: of the
// warps resident on a scheduler some finish long before others (ncu: 5.4 of 8 warps active on average over a launch whose
// grid is exactly one resident wave), and what they leave behind is an under-filled tail.  Two ways to launch:
//   waves   (the round-1 policy) full waves of (resident threads) x CPT chains + one remainder wave of 1..CPT chains per thread;
//   dynamic ONE launch of small blocks (1 warp), many more blocks than fit: the hardware block scheduler back-fills every slot
//           a finished warp frees, so all SMs stay full until the population is exhausted and the tail is one short block.
struct Geom {
  int cpt, nb, block, pack, dynamic;
};

struct Launch {  // everything a launch needs besides its chain range
  const void* c;
  const Gmm2dConstsPk* cpk;
  const void* kdev;
  void* x;
  long long n_chains;
  unsigned long long chain_id0;
  long long n_steps, step0;
  PhiloxKeys keys;
  const void* noise;
  void* traj;
  long long thin;
  cudaStream_t st;
};

template <typename T, int ALG, bool R2, int CPT, int NB, int BLOCK, bool PACK>
static int launch_one(const Launch& a, long long lo, long long n_launch, int* blocks_per_sm) {
  auto kern = gmm2d_kernel<T, ALG, R2, CPT, NB, BLOCK, PACK>;
  if (blocks_per_sm) {
    PSGLA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, BLOCK, 0));
    return PSGLA_OK;
  }
  const long long threads = (n_launch + CPT - 1) / CPT;
  const unsigned grid = (unsigned)((threads + BLOCK - 1) / BLOCK);
  kern<<<grid, BLOCK, 0, a.st>>>(*(const Gmm2dConsts<T>*)a.c, *a.cpk, (const Gmm2dComponents<T>*)a.kdev, (T*)a.x, a.n_chains, lo,
                                 n_launch, a.chain_id0, a.n_steps, a.step0, a.keys, (const T*)a.noise, (T*)a.traj, a.thin);
  return PSGLA_OK;
}

// The instantiated geometries.  fp64 and r != 2 keep the plain 128-thread kernels (parity paths, not throughput paths).
template <typename T, int ALG, bool R2>
static int launch_geom(const Geom& g, const Launch& a, long long lo, long long n, int* bps) {
#define PSGLA_GEOM(C_, N_, B_, P_) \
  if (g.cpt == C_ && g.nb == N_ && g.block == B_ && g.pack == P_) return launch_one<T, ALG, R2, C_, N_, B_, (P_ != 0)>(a, lo, n, bps);
  PSGLA_GEOM(1, 1, 128, 0)
  PSGLA_GEOM(2, 1, 128, 0)
  PSGLA_GEOM(3, 1, 128, 0)
  PSGLA_GEOM(4, 1, 128, 0)
  if constexpr (sizeof(T) == 4 && R2) {
    PSGLA_GEOM(1, 2, 128, 0)
    PSGLA_GEOM(1, 4, 128, 0)
    PSGLA_GEOM(2, 2, 128, 0)
    PSGLA_GEOM(1, 2, 32, 0)
    PSGLA_GEOM(1, 4, 32, 0)
    PSGLA_GEOM(2, 2, 32, 0)
    PSGLA_GEOM(2, 1, 32, 0)
    PSGLA_GEOM(4, 1, 32, 0)
    PSGLA_GEOM(1, 4, 64, 0)
    PSGLA_GEOM(2, 2, 64, 0)
    PSGLA_GEOM(2, 1, 128, 1)
    PSGLA_GEOM(4, 1, 128, 1)
    PSGLA_GEOM(2, 2, 128, 1)
    PSGLA_GEOM(2, 1, 32, 1)
    PSGLA_GEOM(2, 2, 32, 1)
    PSGLA_GEOM(4, 1, 32, 1)
    PSGLA_GEOM(2, 2, 64, 1)
    PSGLA_GEOM(4, 1, 64, 1)
  }
#undef PSGLA_GEOM
  return set_error(PSGLA_E_UNSUPPORTED, "gmm2d geometry cpt=%d nb=%d block=%d pack=%d is not instantiated", g.cpt, g.nb, g.block,
                   g.pack);
}

// Lean-kernel geometries: "nb,block,minb,pk" (dynamic == 2 in PSGLA_GMM_GEOM: cpt = nb, nb = block, block = minb, pack = pk).
template <int ALG, int STRUCT>
static int launch_lean_s(int nb, int block, int minb, int pk, const Launch& a, long long lo, long long n, long long n_rounds,
                         long long step_lo) {
  const Gmm2dConsts<float>& c = *(const Gmm2dConsts<float>*)a.c;
  float2* x = (float2*)a.x + lo;
  const unsigned long long sub0 = a.chain_id0 + (unsigned long long)lo, pair0 = (unsigned long long)step_lo >> 1;
#define PSGLA_LEAN(N_, B_, M_, P_)                                                                                          \
  if (nb == N_ && block == B_ && minb == M_ && pk == P_) {                                                                  \
    gmm2d_lean_kernel<ALG, STRUCT, N_, B_, M_, P_><<<(unsigned)((n + B_ - 1) / B_), B_, 0, a.st>>>(                           \
        c, x, (unsigned)n, sub0, pair0, (int)n_rounds, a.keys);                                                             \
    return PSGLA_OK;                                                                                                        \
  }
  PSGLA_LEAN(4, 64, 16, 1)
  PSGLA_LEAN(4, 64, 16, 0)
  PSGLA_LEAN(2, 64, 16, 0)
  PSGLA_LEAN(4, 64, 20, 1)
  PSGLA_LEAN(4, 64, 24, 0)
  PSGLA_LEAN(4, 32, 32, 1)
  PSGLA_LEAN(8, 32, 32, 0)
  PSGLA_LEAN(1, 64, 16, 2)
  PSGLA_LEAN(2, 64, 16, 2)
  PSGLA_LEAN(2, 64, 16, 3)
  PSGLA_LEAN(4, 64, 16, 2)
  PSGLA_LEAN(4, 64, 16, 3)
  PSGLA_LEAN(1, 64, 24, 2)
  PSGLA_LEAN(2, 64, 24, 2)
  PSGLA_LEAN(2, 64, 20, 3)
  PSGLA_LEAN(4, 64, 12, 3)
  PSGLA_LEAN(4, 64, 12, 2)
#undef PSGLA_LEAN
  return set_error(PSGLA_E_UNSUPPORTED, "lean gmm2d geometry nb=%d block=%d minb=%d pk=%d is not instantiated", nb, block, minb, pk);
}

template <int ALG>
static int launch_lean(int nb, int block, int minb, int pk, const Launch& a, long long lo, long long n, long long n_rounds,
                       long long step_lo) {
  int structure = constants_structure(*(const Gmm2dConsts<float>*)a.c);
  if (const char* e = std::getenv("PSGLA_GMM_STRUCT")) structure = std::min(structure, std::atoi(e));  // A/B: cap the specialisation
  if (structure == 2) return launch_lean_s<ALG, 2>(nb, block, minb, pk, a, lo, n, n_rounds, step_lo);
  if (structure == 1) return launch_lean_s<ALG, 1>(nb, block, minb, pk, a, lo, n, n_rounds, step_lo);
  return launch_lean_s<ALG, 0>(nb, block, minb, pk, a, lo, n, n_rounds, step_lo);
}

// PSGLA_GMM_GEOM="cpt,nb,block,pack,dynamic" overrides the default policy (A/B runs, scripts/gmm2d_sweep.py).
static bool geom_from_env(Geom* g) {
  const char* e = std::getenv("PSGLA_GMM_GEOM");
  if (!e || !*e) return false;
  Geom t{4, 1, 128, 0, 0};
  if (std::sscanf(e, "%d,%d,%d,%d,%d", &t.cpt, &t.nb, &t.block, &t.pack, &t.dynamic) < 1) return false;
  *g = t;
  return true;
}

static thread_local int g_last_launches = 0;  // launches of this thread's most recent psgla_gmm2d_run (bench bookkeeping)

template <typename T, int ALG, bool R2>
static int launch_run(const Folded& f, T* x, long long n_chains, unsigned long long chain_id0, long long n_steps,
                      long long step0, unsigned long long seed, const T* noise, T* traj, long long thin,
                      cudaStream_t st) {
  Gmm2dConsts<T> c;
  static thread_local Gmm2dComponents<T> k;  // pageable source of the async copy must outlive the call: keep it TLS
  narrow<T>(f, &c, &k);
  Gmm2dConstsPk cpk;
  std::memset(&cpk, 0, sizeof(cpk));
  if constexpr (sizeof(T) == 4) cpk = pack_consts(c);
  Gmm2dComponents<T>* kdev = nullptr;
  if (!R2) {
    int rc = upload_components<T>(k, st, &kdev);
    if (rc) return rc;
  }
  const Launch a{&c, &cpk, kdev, x, n_chains, chain_id0, n_steps, step0, philox_round_keys(seed), noise, traj, thin, st};
  constexpr bool fast = sizeof(T) == 4 && R2;
  Geom g{4, 1, 128, 0, 0};
  if (fast && noise == nullptr) {
    g = Geom{PSGLA_GMM_DEFAULT_GEOM};
    geom_from_env(&g);
    if (g.dynamic == 2 && traj != nullptr) g = Geom{1, 4, 32, 0, 1};  // the lean kernel keeps no trajectory
  }
  int launches = 0, rc = PSGLA_OK;
  if (g.dynamic == 2) {
    if constexpr (fast) {
      // lean kernel on the whole rounds of 2 nb steps that start on an even step; the general kernel finishes the rest
      const int nb = g.cpt;
      long long lead = (step0 & 1) ? 1 : 0;
      if (lead > n_steps) lead = n_steps;
      const long long rounds = (n_steps - lead) / (2 * nb), body = rounds * 2 * nb;
      const Geom rest{1, 1, 128, 0, 1};
      if (lead) {
        Launch b = a;
        b.n_steps = lead;
        rc = launch_geom<T, ALG, R2>(rest, b, 0, n_chains, nullptr);
        ++launches;
      }
      for (long long lo = 0; lo < n_chains && rc == PSGLA_OK && rounds > 0; lo += (1ll << 30)) {
        const long long n = n_chains - lo < (1ll << 30) ? n_chains - lo : (1ll << 30);
        rc = launch_lean<ALG>(nb, g.nb, g.block, g.pack, a, lo, n, rounds, step0 + lead);
        ++launches;
      }
      if (rc == PSGLA_OK && lead + body < n_steps) {
        Launch b = a;
        b.step0 = step0 + lead + body;
        b.n_steps = n_steps - lead - body;
        rc = launch_geom<T, ALG, R2>(rest, b, 0, n_chains, nullptr);
        ++launches;
      }
    }
  } else if (g.dynamic) {
    rc = launch_geom<T, ALG, R2>(g, a, 0, n_chains, nullptr);
    launches = 1;
  } else {
    int per_sm = 0;
    rc = launch_geom<T, ALG, R2>(g, a, 0, 0, &per_sm);
    if (rc) return rc;
    const long long resident = (long long)num_sms() * (per_sm > 0 ? per_sm : 8) * g.block;  // threads one wave holds
    long long lo = 0;
    while (lo < n_chains && rc == PSGLA_OK) {
      const long long left = n_chains - lo;
      Geom w = g;
      long long n = (long long)g.cpt * resident;
      if (left < n) {  // remainder wave: as few chains per thread as still fit in one wave (pairs stay pairs)
        const int step = g.pack ? 2 : 1;
        w.cpt = (int)((left + resident - 1) / resident);
        w.cpt = (w.cpt + step - 1) / step * step;
        if (w.cpt < g.cpt && w.nb * w.cpt < g.nb * g.cpt && !g.pack) w.nb = 1;
        n = left;
        if (launch_geom<T, ALG, R2>(w, a, 0, 0, &per_sm) != PSGLA_OK) w = g;  // not instantiated: the full geometry, partly idle
      }
      rc = launch_geom<T, ALG, R2>(w, a, lo, n, nullptr);
      lo += n;
      ++launches;
    }
  }
  if (rc) return rc;
  g_last_launches = launches;
  PSGLA_CUDA_TRY(cudaGetLastError());
  if (kdev) PSGLA_CUDA_TRY(cudaFreeAsync(kdev, st));
  return PSGLA_OK;
}

template <typename T>
static int dispatch_run(const psgla_gmm2d_problem* p, const Folded& f, void* x, long long n_chains,
                        unsigned long long chain_id0, long long n_steps, long long step0, unsigned long long seed,
                        const void* noise, void* traj, long long thin, cudaStream_t st) {
  const bool r2 = p->n_components == 2;
  T* xx = (T*)x;
  const T* nz = (const T*)noise;
  T* tj = (T*)traj;
  if (p->alg == PSGLA_ALG_PSGLA)
    return r2 ? launch_run<T, PSGLA_ALG_PSGLA, true>(f, xx, n_chains, chain_id0, n_steps, step0, seed, nz, tj, thin, st)
              : launch_run<T, PSGLA_ALG_PSGLA, false>(f, xx, n_chains, chain_id0, n_steps, step0, seed, nz, tj, thin, st);
  return r2 ? launch_run<T, PSGLA_ALG_PNPULA, true>(f, xx, n_chains, chain_id0, n_steps, step0, seed, nz, tj, thin, st)
            : launch_run<T, PSGLA_ALG_PNPULA, false>(f, xx, n_chains, chain_id0, n_steps, step0, seed, nz, tj, thin, st);
}

}  // namespace psgla

using namespace psgla;

extern "C" int psgla_gmm2d_run(const psgla_gmm2d_problem* problem, int precision, void* x_dev, int64_t n_chains,
                               int64_t chain_id0, int64_t n_steps, int64_t step0, uint64_t seed,
                               const void* noise_dev, void* traj_dev, int64_t thin, void* stream) {
  PSGLA_REQUIRE(problem != nullptr && x_dev != nullptr, "psgla_gmm2d_run: null problem or state pointer");
  PSGLA_REQUIRE(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (fp64), got %d", precision);
  PSGLA_REQUIRE(n_chains > 0 && n_steps >= 0 && step0 >= 0 && chain_id0 >= 0, "negative size or offset");
  PSGLA_REQUIRE(traj_dev == nullptr || thin >= 1, "thin must be >= 1 when a trajectory buffer is given");
  if (n_steps == 0) return PSGLA_OK;
  Folded f;
  const double eps = problem->alg == PSGLA_ALG_PSGLA ? problem->delta : problem->epsilon;
  int rc = fold_problem(problem, eps, true, &f);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == 0)
    return dispatch_run<float>(problem, f, x_dev, n_chains, (unsigned long long)chain_id0, n_steps, step0, seed,
                               noise_dev, traj_dev, thin, st);
  return dispatch_run<double>(problem, f, x_dev, n_chains, (unsigned long long)chain_id0, n_steps, step0, seed,
                              noise_dev, traj_dev, thin, st);
}

template <typename T>
static int denoise_impl(const psgla_gmm2d_problem* p, const Folded& f, const void* x, void* out, long long n,
                        cudaStream_t st) {
  Gmm2dConsts<T> c;
  static thread_local Gmm2dComponents<T> k;
  narrow<T>(f, &c, &k);
  Gmm2dComponents<T>* kdev = nullptr;
  int rc = upload_components<T>(k, st, &kdev);
  if (rc) return rc;
  const int block = 128;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  if (p->n_components == 2)
    gmm2d_denoise_kernel<T, true><<<grid, block, 0, st>>>(c, kdev, (const T*)x, (T*)out, n);
  else
    gmm2d_denoise_kernel<T, false><<<grid, block, 0, st>>>(c, kdev, (const T*)x, (T*)out, n);
  PSGLA_CUDA_TRY(cudaGetLastError());
  PSGLA_CUDA_TRY(cudaFreeAsync(kdev, st));
  return PSGLA_OK;
}

extern "C" int psgla_gmm2d_denoise(const psgla_gmm2d_problem* problem, double epsilon, int precision,
                                   const void* x_dev, void* out_dev, int64_t n, void* stream) {
  PSGLA_REQUIRE(problem != nullptr && x_dev != nullptr && out_dev != nullptr, "psgla_gmm2d_denoise: null pointer");
  PSGLA_REQUIRE(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (fp64), got %d", precision);
  PSGLA_REQUIRE(n >= 0, "negative size");
  if (n == 0) return PSGLA_OK;
  Folded f;
  int rc = fold_problem(problem, epsilon, false, &f);
  if (rc) return rc;
  return precision == 0 ? denoise_impl<float>(problem, f, x_dev, out_dev, n, (cudaStream_t)stream)
                        : denoise_impl<double>(problem, f, x_dev, out_dev, n, (cudaStream_t)stream);
}

extern "C" int psgla_gmm2d_last_launches(void) { return g_last_launches; }

extern "C" int psgla_gmm2d_noise(float* out_dev, int64_t n_chains, int64_t chain_id0, int64_t n_steps, int64_t step0,
                                 uint64_t seed, void* stream) {
  PSGLA_REQUIRE(out_dev != nullptr && n_chains > 0 && n_steps >= 0 && step0 >= 0 && chain_id0 >= 0,
                "psgla_gmm2d_noise: bad argument");
  if (n_steps == 0) return PSGLA_OK;
  const int block = 128;
  gmm2d_noise_kernel<<<(unsigned)((n_chains + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
      out_dev, n_chains, (unsigned long long)chain_id0, n_steps, step0, seed);
  PSGLA_CUDA_TRY(cudaGetLastError());
  return PSGLA_OK;
}
