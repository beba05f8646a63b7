/* TEST INFRASTRUCTURE ONLY -- plain-C (float64) restatement of the reference's 2D path, operation for operation:
 *   Theorical_MMSE   utils_2D.py:209-233   (sqrt(epsilon) where a variance belongs, plain exp, no log-sum-exp: kept)
 *   PnP_ULA          sampling_2D.py:21-45
 *   SnoPnP_ULA       sampling_2D.py:48-72  (the paper's PSGLA)
 *   score            sampling_2D.py:30-31,57-58
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs (as the checker and as a
 * compiled CPU baseline).  The product never links or loads it.
 * Parity status: PINNED through tests/test_oracle_c.py -- against oracle/gmm2d_oracle.py (itself pinned to the unmodified
 * reference) on seeded noise, and against the golden vectors tests/golden/gmm2d_golden.json made by running the reference.
 * Built by __graft_entry__.build() / oracle/c_oracle.py:  gcc -O2 -shared -fPIC -o oracle/_build/libgmm2d_oracle.so
 * (-O2 without -ffast-math: IEEE evaluation order as written). */
#include <math.h>
#include <stddef.h>

#define RMAX 16

typedef struct {
  double a, b, c, d; /* [[a b][c d]] */
} mat2;

static mat2 inv2(mat2 m) { /* np.linalg.inv of a 2x2 */
  const double det = m.a * m.d - m.b * m.c;
  mat2 r = {m.d / det, -m.b / det, -m.c / det, m.a / det};
  return r;
}
static double det2(mat2 m) { return m.a * m.d - m.b * m.c; }

/* D(x, epsilon), utils_2D.py:219-232.  mu [r][2], Sigma [r][4] row-major, pi [r]. */
void gmm2d_denoise(int r, const double* mu, const double* Sigma, const double* pi, const double* x, double epsilon,
                   double* out) {
  const double tau = sqrt(epsilon); /* :223-226 use np.sqrt(epsilon) * Id */
  double A0 = 0.0, A1 = 0.0, B = 0.0;
  for (int i = 0; i < r && i < RMAX; ++i) {
    const mat2 Sig = {Sigma[4 * i], Sigma[4 * i + 1], Sigma[4 * i + 2], Sigma[4 * i + 3]};
    const mat2 Sinv = inv2(Sig);                                   /* :215-217 */
    const mat2 T = {tau + Sig.a, Sig.b, Sig.c, tau + Sig.d};       /* sqrt(eps) Id + Sigma_i */
    const mat2 Ti = inv2(T);
    const double d0 = x[0] - mu[2 * i], d1 = x[1] - mu[2 * i + 1];
    const double q = d0 * (Ti.a * d0 + Ti.b * d1) + d1 * (Ti.c * d0 + Ti.d * d1);
    double c = exp(-0.5 * q);                                       /* :223 */
    c = c / sqrt(det2(T));                                          /* :224 */
    const mat2 P = inv2((mat2){1.0 / tau + Sinv.a, Sinv.b, Sinv.c, 1.0 / tau + Sinv.d}); /* (Id/tau + Sigma_i^-1)^-1 */
    const double v0 = x[0] / tau + (Sinv.a * mu[2 * i] + Sinv.b * mu[2 * i + 1]);
    const double v1 = x[1] / tau + (Sinv.c * mu[2 * i] + Sinv.d * mu[2 * i + 1]);
    const double m0 = P.a * v0 + P.b * v1, m1 = P.c * v0 + P.d * v1; /* :226 */
    A0 += c * pi[i] * m0;                                           /* :229-231 */
    A1 += c * pi[i] * m1;
    B += c * pi[i];
  }
  out[0] = A0 / B;
  out[1] = A1 / B;
}

static void score(const double* A, const double* y, const double* x, double sigma, double* g) {
  /* A^T (y - A x) / sigma^2, sampling_2D.py:30-31; A row-major 2x2 */
  const double r0 = y[0] - (A[0] * x[0] + A[1] * x[1]);
  const double r1 = y[1] - (A[2] * x[0] + A[3] * x[1]);
  g[0] = (A[0] * r0 + A[2] * r1) / (sigma * sigma);
  g[1] = (A[1] * r0 + A[3] * r1) / (sigma * sigma);
}

/* traj [N][2] (row 0 = x_0), noise [N-1][2] standard normals replacing np.random.randn(2) (sampling_2D.py:35). */
void gmm2d_pnp_ula(long N, const double* x0, const double* y, double delta, const double* A, double sigma, int r,
                   const double* mu, const double* Sigma, const double* pi, double epsilon, double alpha,
                   const double* noise, double* traj) {
  traj[0] = x0[0];
  traj[1] = x0[1];
  const double sn = sqrt(2.0 * delta);
  for (long i = 0; i + 1 < N; ++i) {
    const double* x = traj + 2 * i;
    double g[2], D[2];
    score(A, y, x, sigma, g);
    gmm2d_denoise(r, mu, Sigma, pi, x, epsilon, D);
    /* x + delta * score + alpha * delta * (1 / epsilon) * (D - x) + sqrt(2 delta) z        sampling_2D.py:36 */
    traj[2 * i + 2] = x[0] + delta * g[0] + alpha * delta * (1.0 / epsilon) * (D[0] - x[0]) + sn * noise[2 * i];
    traj[2 * i + 3] = x[1] + delta * g[1] + alpha * delta * (1.0 / epsilon) * (D[1] - x[1]) + sn * noise[2 * i + 1];
  }
}

void gmm2d_snopnp_ula(long N, const double* x0, const double* y, double delta, const double* A, double sigma, int r,
                      const double* mu, const double* Sigma, const double* pi, double alpha, const double* noise,
                      double* traj) {
  traj[0] = x0[0];
  traj[1] = x0[1];
  const double sn = sqrt(2.0 * delta);
  for (long i = 0; i + 1 < N; ++i) {
    const double* x = traj + 2 * i;
    double g[2], u[2];
    score(A, y, x, sigma, g);
    /* D(x + (delta / alpha) * score + sqrt(2 delta) z, delta)                                sampling_2D.py:63 */
    u[0] = x[0] + (delta / alpha) * g[0] + sn * noise[2 * i];
    u[1] = x[1] + (delta / alpha) * g[1] + sn * noise[2 * i + 1];
    gmm2d_denoise(r, mu, Sigma, pi, u, delta, traj + 2 * i + 2);
  }
}

/* Compiled-CPU baseline: n_steps of one chain from x0 with a private xorshift64* / Box-Muller stream (the noise source is not
 * part of the parity claim; the arithmetic per step is the functions above).  alg 0 = PSGLA, 1 = PnP-ULA.  out = final state. */
void gmm2d_run_chain(int alg, long n_steps, const double* x0, const double* y, double delta, const double* A, double sigma,
                     int r, const double* mu, const double* Sigma, const double* pi, double epsilon, double alpha,
                     unsigned long long seed, double* out) {
  unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  double x[2] = {x0[0], x0[1]};
  const double sn = sqrt(2.0 * delta);
  for (long i = 0; i < n_steps; ++i) {
    double z[2];
    for (int k = 0; k < 2; ++k) { /* two uniforms -> one normal pair would do; kept simple and branch-free */
      s ^= s >> 12, s ^= s << 25, s ^= s >> 27;
      const double u1 = ((double)((s * 0x2545F4914F6CDD1Dull) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      s ^= s >> 12, s ^= s << 25, s ^= s >> 27;
      const double u2 = ((double)((s * 0x2545F4914F6CDD1Dull) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      z[k] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
    double g[2], D[2];
    score(A, y, x, sigma, g);
    if (alg == 0) {
      double u[2] = {x[0] + (delta / alpha) * g[0] + sn * z[0], x[1] + (delta / alpha) * g[1] + sn * z[1]};
      gmm2d_denoise(r, mu, Sigma, pi, u, delta, x);
    } else {
      gmm2d_denoise(r, mu, Sigma, pi, x, epsilon, D);
      x[0] = x[0] + delta * g[0] + alpha * delta * (1.0 / epsilon) * (D[0] - x[0]) + sn * z[0];
      x[1] = x[1] + delta * g[1] + alpha * delta * (1.0 / epsilon) * (D[1] - x[1]) + sn * z[1];
    }
  }
  out[0] = x[0];
  out[1] = x[1];
}
