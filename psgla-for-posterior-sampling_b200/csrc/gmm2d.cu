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
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace psgla {

constexpr int RMAX = PSGLA_GMM_MAX_COMPONENTS;

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

// ------------------------------------------------------------------------------------------------ the kernel
// Chain j of thread g is chain index g + j * (gridDim.x * blockDim.x): every global access is coalesced.
template <typename T, int ALG, bool R2, int CPT>
__global__ void __launch_bounds__(128, (sizeof(T) == 4 && R2) ? 8 : 1)
gmm2d_kernel(const Gmm2dConsts<T> c, const Gmm2dComponents<T>* __restrict__ comps_gmem, T* __restrict__ x,
             long long n_chains, long long chain_lo, long long n_launch, unsigned long long chain_id0, long long n_steps,
             long long step0, const PhiloxKeys keys, const T* __restrict__ noise, T* __restrict__ traj, long long thin) {
  // This launch owns chains [chain_lo, chain_lo + n_launch) of the call's n_chains (the stride of noise / traj rows).
  using V2 = typename Vec2<T>::type;
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
    while (t < t_end) {
      float z[CPT][4];
      const unsigned long long pair = (unsigned long long)t >> 1;
#pragma unroll
      for (int j = 0; j < CPT; ++j)
        philox_normal4_keyed(keys, chain_id0 + (unsigned long long)(chain_lo + g + j * nthreads), (uint32_t)pair,
                             (uint32_t)(pair >> 32), z[j][0], z[j][1], z[j][2], z[j][3]);
      if ((t & 1) == 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], T(z[j][0]), T(z[j][1]));
        after_step();
        ++t;
        if (t >= t_end) break;
      }
#pragma unroll
      for (int j = 0; j < CPT; ++j) langevin_step<T, ALG, R2>(c, comps, x0[j], x1[j], T(z[j][2]), T(z[j][3]));
      after_step();
      ++t;
    }
  }

#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const long long ch = chain_lo + g + j * nthreads;
    if (live[j]) reinterpret_cast<V2*>(x)[ch] = V2{x0[j], x1[j]};
  }
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

// Launch geometry.  The kernel is issue-bound, so what matters is that every SM holds the same number of warps for
// the whole launch: the population is cut into full waves of (resident threads) x 4 chains, and the remainder runs as
// one more wave with 1..4 chains per thread.  (A single launch of 10^6 chains at 4 per thread is 1.65 waves: the second,
// two-thirds-empty wave costs as much as the first.)
template <typename T, int ALG, bool R2, int CPT>
static void launch_wave(const Gmm2dConsts<T>& c, const Gmm2dComponents<T>* kdev, T* x, long long n_chains, long long lo,
                        long long n_launch, unsigned long long chain_id0, long long n_steps, long long step0,
                        unsigned long long seed, const T* noise, T* traj, long long thin, cudaStream_t st) {
  const int block = 128;
  const long long threads = (n_launch + CPT - 1) / CPT;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  gmm2d_kernel<T, ALG, R2, CPT><<<grid, block, 0, st>>>(c, kdev, x, n_chains, lo, n_launch, chain_id0, n_steps, step0,
                                                      philox_round_keys(seed), noise, traj, thin);
}

// Number of kernel launches launch_run makes for n_chains (full 4-chain waves + one remainder wave).
static int count_waves(long long n_chains, long long resident) {
  const long long full = n_chains / (4 * resident);
  return (int)full + ((n_chains - full * 4 * resident) > 0 ? 1 : 0);
}
static int g_last_launches = 0;  // launches of the most recent psgla_gmm2d_run on this thread's behalf (bench bookkeeping)

template <typename T, int ALG, bool R2>
static int launch_run(const Folded& f, T* x, long long n_chains, unsigned long long chain_id0, long long n_steps,
                      long long step0, unsigned long long seed, const T* noise, T* traj, long long thin,
                      cudaStream_t st) {
  Gmm2dConsts<T> c;
  static thread_local Gmm2dComponents<T> k;  // pageable source of the async copy must outlive the call: keep it TLS
  narrow<T>(f, &c, &k);
  Gmm2dComponents<T>* kdev = nullptr;
  if (!R2) {
    int rc = upload_components<T>(k, st, &kdev);
    if (rc) return rc;
  }
  static long long resident = 0;  // threads of the 4-chain kernel one wave holds
  if (!resident) {
    int per_sm = 0;
    PSGLA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmm2d_kernel<T, ALG, R2, 4>, 128, 0));
    resident = (long long)num_sms() * (per_sm > 0 ? per_sm : 8) * 128;
  }
  g_last_launches = count_waves(n_chains, resident);
  long long lo = 0;
  while (lo < n_chains) {
    const long long left = n_chains - lo;
    if (left >= 4 * resident) {
      launch_wave<T, ALG, R2, 4>(c, kdev, x, n_chains, lo, 4 * resident, chain_id0, n_steps, step0, seed, noise, traj, thin, st);
      lo += 4 * resident;
      continue;
    }
    const int cpt = (int)((left + resident - 1) / resident);  // 1..4
    if (cpt <= 1)
      launch_wave<T, ALG, R2, 1>(c, kdev, x, n_chains, lo, left, chain_id0, n_steps, step0, seed, noise, traj, thin, st);
    else if (cpt == 2)
      launch_wave<T, ALG, R2, 2>(c, kdev, x, n_chains, lo, left, chain_id0, n_steps, step0, seed, noise, traj, thin, st);
    else if (cpt == 3)
      launch_wave<T, ALG, R2, 3>(c, kdev, x, n_chains, lo, left, chain_id0, n_steps, step0, seed, noise, traj, thin, st);
    else
      launch_wave<T, ALG, R2, 4>(c, kdev, x, n_chains, lo, left, chain_id0, n_steps, step0, seed, noise, traj, thin, st);
    lo = n_chains;
  }
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
