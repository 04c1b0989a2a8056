/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the compiled logp/dlogp callable of the Radon benchmark model
 * (benchmarks/benchmarks/benchmarks.py:34-45), the same formulas as oracle/logp_numpy.py::RadonLogp, each citing the
 * reference line it follows.  The reference evaluates this function as a PyTensor C thunk (model/core.py:232-267); PyTensor
 * is absent here, so the CPU baseline of bench.py (cpu_baseline / --impl reference) calls this gcc -O3 build instead of the
 * NumPy form: the per-call cost is then that of compiled code, like the reference's, and the Python NUTS around it is the
 * reference's own structure.  Checked against the NumPy form to 1e-12 (tests/test_oracle_logp.py).  Never linked into
 * libb200nuts.so; nothing under pymc_b200/ loads it.
 *
 * q = [mu_a, log sigma_a, mu_b, log sigma_b, a[J], b[J], log eps]
 */
#include <math.h>
#include <stddef.h>

#define HALF_LOG_2PI 0.91893853320467274178
#define LOG_2 0.69314718055994530942
#define LOG_PI 1.1447298858494001741

/* HalfCauchy(beta) on x = exp(z) plus the log-transform Jacobian: Cauchy.logp continuous.py:2287-2288, HalfCauchy.logp
 * :2383-2385, LogTransform logprob/transforms.py:880-891 */
static double halfcauchy_log(double z, double beta, double* dz) {
    const double x = exp(z);
    const double t = x / beta;
    const double u = t * t;
    *dz = 1.0 - 2.0 * u / (1.0 + u);
    return LOG_2 - LOG_PI - log(beta) - log1p(u) + z;
}

/* Normal.logp continuous.py:526-527 */
static double normal_logp(double x, double mu, double sigma) {
    const double z = (x - mu) / sigma;
    return -0.5 * z * z - HALF_LOG_2PI - log(sigma);
}

/* returns logp, writes grad[n]; scratch: Ga[J], Gb[J] */
double radon_logp_dlogp(const double* q, int J, long n_obs, const long* idx, const double* x, const double* y, double* grad,
                        double* Ga, double* Gb) {
    const double mu_a = q[0], lsa = q[1], mu_b = q[2], lsb = q[3];
    const double* a = q + 4;
    const double* b = q + 4 + J;
    const double leps = q[4 + 2 * J];
    const double sa = exp(lsa), sb = exp(lsb), eps = exp(leps);
    const double S = 100.0 * 100.0; /* Normal(0, sigma=100**2), benchmarks.py:36-37 */
    double dhca, dhcb, dhce;
    double logp = normal_logp(mu_a, 0.0, S) + normal_logp(mu_b, 0.0, S) + halfcauchy_log(lsa, 5.0, &dhca) +
                  halfcauchy_log(lsb, 5.0, &dhcb) + halfcauchy_log(leps, 5.0, &dhce);
    double sa2 = 0.0, sb2 = 0.0;
    for (int j = 0; j < J; ++j) {
        sa2 += a[j] * a[j];
        sb2 += b[j] * b[j];
        Ga[j] = 0.0;
        Gb[j] = 0.0;
    }
    logp += -0.5 * sa2 - J * HALF_LOG_2PI - 0.5 * sb2 - J * HALF_LOG_2PI; /* a, b ~ Normal(0, 1) (non-centred) */
    const double inv_e2 = 1.0 / (eps * eps);
    double ssq = 0.0;
    for (long i = 0; i < n_obs; ++i) { /* y ~ Normal(alpha[county] + beta[county] * floor, eps) */
        const long c = idx[i];
        const double m = (mu_a + sa * a[c]) + (mu_b + sb * b[c]) * x[i];
        const double d = y[i] - m;
        const double r = d * inv_e2;
        ssq += d * d;
        Ga[c] += r;
        Gb[c] += r * x[i];
    }
    logp += -0.5 * ssq * inv_e2 - n_obs * (HALF_LOG_2PI + leps);
    double sGa = 0.0, sGb = 0.0, aGa = 0.0, bGb = 0.0;
    for (int j = 0; j < J; ++j) {
        sGa += Ga[j];
        sGb += Gb[j];
        aGa += a[j] * Ga[j];
        bGb += b[j] * Gb[j];
        grad[4 + j] = -a[j] + sa * Ga[j];
        grad[4 + J + j] = -b[j] + sb * Gb[j];
    }
    grad[0] = -mu_a / (S * S) + sGa;
    grad[1] = dhca + sa * aGa;
    grad[2] = -mu_b / (S * S) + sGb;
    grad[3] = dhcb + sb * bGb;
    grad[4 + 2 * J] = dhce + ssq * inv_e2 - (double)n_obs;
    return logp;
}
