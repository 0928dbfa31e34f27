// Polarization factors J = (I, Q, U) of every geodesic sample (setup, once per (spin, inclination)): the chain of
// alma.image_plane_model (bhnerf/alma.py:47-60) = kgeo.azimuthal_velocity_vector (bhnerf/kgeo.py:199-223) ->
// doppler_factor (:225-248) -> magnetic_field_fluid_frame (:274-313, fluid_frame_tetrad :315-350) -> normalisation by the
// mean field strength inside the recovery domain (alma.py:55-57) -> parallel_transport (:438-519, V_frac = 0) -> NaN -> 0.
// float64 arithmetic like the reference's numpy; one thread per sample, neighbours along the ray give the turning-point
// signs of the wave vector (np.gradient, kgeo.py:108-109).  For the purely azimuthal 4-velocity the reference uses
// (u^r = u^theta = 0) the tetrad collapses to e_t = -u, e_r = (0,1,0,0)/sqrt(g_rr), e_th = (0,0,1,0)/sqrt(g_thth),
// e_ph = (u_ph, 0, 0, -u_t)/(sqrt(Delta) sin th); the restatement is pinned on the reference's own functions
// (oracle/ref_shim.reference_polarization_factors, tests/golden/make_golden_pol.py).
#include "common.cuh"

namespace {

struct PolConsts {
  double a, inc, omega_sign, arad, avert, ator, Q_frac, rmin, rmax, z_width;
  int spectral_index;
};

__device__ __forceinline__ double sgn(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : (x == 0.0 ? 0.0 : x)); }   // NaN stays NaN
__device__ __forceinline__ double grad_at(const double* f, int k, int G) {      // np.gradient along the ray, unit spacing
  if (G < 2) return 0.0;
  if (k == 0) return f[1] - f[0];
  if (k == G - 1) return f[G - 1] - f[G - 2];
  return 0.5 * (f[k + 1] - f[k - 1]);
}

// J_raw [3][n]: factors with the UN-normalised field; acc[0] += |b| over the domain, acc[1] += 1
__global__ void __launch_bounds__(256)
pol_factors_kernel(const double* __restrict__ r_, const double* __restrict__ th_, const double* __restrict__ aff_,
                   const double* __restrict__ lam_, const double* __restrict__ eta_, const double* __restrict__ alpha_,
                   const double* __restrict__ beta_, const double* __restrict__ omega_in, long long P, int G, PolConsts c,
                   double* __restrict__ J_raw, double* __restrict__ acc) {
  const long long n = P * G;
  double loc_sum = 0.0, loc_cnt = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / G;
    const int k = (int)(i - p * G);
    const double r = r_[i], th = th_[i], a = c.a, lam = lam_[p], eta = eta_[p];
    const double sth = sin(th), cth = cos(th);
    const double Delta = r * r + a * a - 2.0 * r, Sigma = r * r + a * a * cth * cth;
    const double Xi = (r * r + a * a) * (r * r + a * a) - a * a * Delta * sth * sth;
    double Rpot = (r * r + a * a - a * lam) * (r * r + a * a - a * lam) - Delta * (eta + (lam - a) * (lam - a));
    if (!(fabs(Rpot) > 1e-10)) Rpot = 0.0;                                        // kerr_raytracing_utils.py:270
    const double tn = tan(th);
    const double Thpot = eta + a * a * cth * cth - lam * lam / (tn * tn);
    const double ga = grad_at(aff_ + p * G, k, G);
    const double pm_r = sgn(grad_at(r_ + p * G, k, G) / ga), pm_th = sgn(grad_at(th_ + p * G, k, G) / ga);
    const double k_t = -1.0, k_r = sqrt(fmax(Rpot, 0.0)) * pm_r / Delta, k_th = sqrt(fmax(Thpot, 0.0)) * pm_th, k_ph = lam;
    const double g_tt = -(1.0 - 2.0 * r / Sigma), g_rr = Sigma / Delta, g_thth = Sigma;
    const double g_phph = Xi * sth * sth / Sigma, g_tph = -2.0 * a * r * sth * sth / Sigma;
    const double Om = omega_in ? omega_in[i] : c.omega_sign / (r * sqrt(r) + a);
    const double ut = 1.0 / sqrt(-(g_tt + 2.0 * Om * g_tph + g_phph * Om * Om)), uph = ut * Om;
    const double u_t = g_tt * ut + g_tph * uph, u_ph = g_phph * uph + g_tph * ut;
    const double s = u_t * ut + u_ph * uph;
    const double N_r = sqrt(-g_rr * s), N_th = sqrt(g_thth), N_ph = sqrt(-s * Delta * sth * sth);
    double gdop = 1.0 / -(k_t * ut + k_ph * uph);
    if (gdop != gdop) gdop = 0.0;
    const double kp0 = -s * k_r / N_r, kp1 = k_th / N_th, kp2 = (u_ph * k_t - u_t * k_ph) / N_ph;
    const double Br = c.arad * sth + c.avert * cth, Bth = -c.avert * sth, Bph = c.ator;
    const double b0 = Bph * u_ph, b1 = Br / u_t, b2 = Bth / u_t, b3 = (Bph + b0 * u_ph) / u_t;
    const double bl0 = g_tt * b0 + g_tph * b3, bl1 = g_rr * b1, bl2 = g_thth * b2, bl3 = g_phph * b3 + g_tph * b0;
    const double bp0 = -s * bl1 / N_r, bp1 = bl2 / N_th, bp2 = (u_ph * bl0 - u_t * bl3) / N_ph;
    const double b_mag = sqrt(bp0 * bp0 + bp1 * bp1 + bp2 * bp2);
    const double z = r * cth;
    if (fabs(z) < c.z_width && r > c.rmin && r < c.rmax) { loc_sum += b_mag; loc_cnt += 1.0; }   // NaN propagates like the reference's mean
    const double k_mag = sqrt(kp0 * kp0 + kp1 * kp1 + kp2 * kp2);
    const double f0 = (kp1 * bp2 - kp2 * bp1) / k_mag, f1 = (kp2 * bp0 - kp0 * bp2) / k_mag, f2 = (kp0 * bp1 - kp1 * bp0) / k_mag;
    const double f_t = u_ph / N_ph * f2, f_r = -s / N_r * f0, f_th = f1 / N_th, f_ph = -u_t / N_ph * f2;
    const double sin_b = sqrt(f0 * f0 + f1 * f1 + f2 * f2) / k_mag;               // (un-normalised b: scales out below)
    const int si = c.spectral_index;
    // I = g^si * b_mag^(si+1) * sin_b^(si+1); sin_b carries one factor 1/b_mean through f, b_mag the other powers: the
    // normalisation kernel multiplies by (1/b_mean)^(2 si + 2)
    const double I = pow(gdop, (double)si) * pow(b_mag, (double)(si + 1)) * pow(sin_b, (double)(si + 1));
    const double gi_tt = -Xi / (Delta * Sigma), gi_rr = Delta / Sigma, gi_thth = 1.0 / Sigma;
    const double gi_phph = (Delta - a * a * sth * sth) / (Delta * Sigma * sth * sth), gi_tph = -2.0 * a * r / (Delta * Sigma);
    const double ku0 = gi_tt * k_t + gi_tph * k_ph, ku1 = gi_rr * k_r, ku2 = gi_thth * k_th, ku3 = gi_phph * k_ph + gi_tph * k_t;
    const double A = (ku0 * f_r - ku1 * f_t) + a * sth * sth * (ku1 * f_ph - ku3 * f_r);
    const double B = ((r * r + a * a) * (ku3 * f_th - ku2 * f_ph) - a * (ku0 * f_th - ku2 * f_t)) * sth;
    // kappa = (r - i a cos th)(A - i B); chi2 = angle( (beta + i mu) conj(kappa) / ((beta - i mu) kappa) )
    const double kr = r * A - a * cth * B, ki = -(r * B + a * cth * A);
    const double mu = -(alpha_[p] + a * sin(c.inc)), be = beta_[p];
    // numerator n = (be + i mu)(kr - i ki), denominator d = (be - i mu)(kr + i ki) = conj(n)  ->  n/d = n^2/|n|^2
    const double nr = be * kr + mu * ki, ni = mu * kr - be * ki;
    const double chi2 = atan2(2.0 * nr * ni, nr * nr - ni * ni);
    const double Q = c.Q_frac * I;
    J_raw[i] = I; J_raw[n + i] = cos(chi2) * Q; J_raw[2 * n + i] = sin(chi2) * Q;
  }
  // block reduction of the domain statistics
  __shared__ double sh_s[256], sh_c[256];
  sh_s[threadIdx.x] = loc_sum; sh_c[threadIdx.x] = loc_cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh_s[threadIdx.x] += sh_s[threadIdx.x + o]; sh_c[threadIdx.x] += sh_c[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0 && sh_c[0] > 0.0) { atomicAdd(acc, sh_s[0]); atomicAdd(acc + 1, sh_c[0]); }
}

__global__ void __launch_bounds__(256)
pol_normalise_kernel(const double* __restrict__ J_raw, const double* __restrict__ acc, long long n3, int spectral_index,
                     float* __restrict__ J) {
  const double b_mean = acc[0] / acc[1];
  // the reference normalises b BEFORE parallel_transport; there both b_mag and its "sin_th_b" = |k' x b| / |k'|^2 are
  // linear in b, so I, Q, U scale as b^(2 si + 2) and the rotation angle not at all
  const double scale = pow(1.0 / b_mean, (double)(2 * spectral_index + 2));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
    double v = J_raw[i] * scale;
    if (v != v) v = 0.0;                                                          // np.nan_to_num (alma.py:60)
    v = fmin(fmax(v, -3.0e38), 3.0e38);
    J[i] = (float)v;
  }
}

}  // namespace

extern "C" size_t bhnerf_polarization_workspace_bytes(int64_t P, int32_t G) {
  return (size_t)(3 * P * G + 2) * sizeof(double) + 256;
}

extern "C" int bhnerf_polarization_factors(const double* r, const double* theta, const double* affine, const double* lam,
                                           const double* eta, const double* alpha, const double* beta,
                                           const double* Omega_in, int64_t P, int32_t G, double spin, double inclination,
                                           double omega_sign, double arad, double avert, double ator, double Q_frac,
                                           double rmin, double rmax, double z_width, int32_t spectral_index, float* J,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BH_REQUIRE(r && theta && affine && lam && eta && alpha && beta && J && workspace, "polarization_factors: NULL argument");
  BH_REQUIRE(P > 0 && G > 1, "polarization_factors: P must be > 0 and G > 1");
  BH_REQUIRE(Q_frac >= 0.0 && Q_frac <= 1.0, "Q_frac should be in [0,1]");                       // kgeo.py:472
  BH_REQUIRE(workspace_bytes >= bhnerf_polarization_workspace_bytes(P, G), "polarization_factors: workspace too small");
  PolConsts c{spin, inclination, omega_sign, arad, avert, ator, Q_frac, rmin, rmax, z_width, spectral_index};
  double* acc = (double*)workspace;
  double* J_raw = (double*)((char*)workspace + 256);
  BH_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  const long long n = (long long)P * G;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  BhProfScope ps(BH_CAT_MISC, 2, st);
  pol_factors_kernel<<<blocks, 256, 0, st>>>(r, theta, affine, lam, eta, alpha, beta, Omega_in, (long long)P, G, c, J_raw, acc);
  pol_normalise_kernel<<<blocks, 256, 0, st>>>(J_raw, acc, 3 * n, spectral_index, J);
  BH_CHECK_CUDA(cudaGetLastError());
  return 0;
}
