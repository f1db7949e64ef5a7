/* C restatement of the MaxStyle layer -- TEST INFRASTRUCTURE ONLY (second, independent checker next to the numpy oracle).
 *
 * Plain C99, double precision, scalar loops; follows the reference's src/advanced/maxstyle.py line by line:
 *   forward  :140-189  (instance statistics :157-159, normalise :161, batch std :165-168, mixing :172-179, noise + affine :181-185)
 *   backward : the closed form autograd derives from those lines with mu / sig detached at :160
 * Built by oracle/build_c.py into oracle/_build/libmaxstyle_oracle.so and checked against the reference-generated goldens by
 * tests/test_oracle_golden.py.  Nothing in the product package links or loads it.
 *
 * Layout: x, y, dy, dx are [N, C, M] row-major doubles (M = H*W); tables [N, C]; lmda [N]; gamma_std / beta_std [C]; perm int64 [N].
 * flags: bit 0 mix_style, bit 1 no_noise, bit 2 compute gamma_std / beta_std (first forward), bit 3 do not clamp lmda (MixStyle).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

/* mu = mean over M, sig = sqrt(unbiased var + eps)   (maxstyle.py:157-159) */
void ms_oracle_stats(const double* x, int N, int C, int64_t M, double eps, double* mu, double* sig) {
    for (int64_t p = 0; p < (int64_t)N * C; ++p) {
        const double* v = x + p * M;
        double s = 0.0;
        for (int64_t i = 0; i < M; ++i) s += v[i];
        const double m = s / (double)M;
        double q = 0.0;
        for (int64_t i = 0; i < M; ++i) q += (v[i] - m) * (v[i] - m);
        mu[p] = m;
        sig[p] = sqrt(q / (double)(M - 1) + eps);
    }
}

/* unbiased std over the batch dimension of an [N, C] table   (torch.std(t, dim=0), maxstyle.py:166,168) */
void ms_oracle_batch_std(const double* table, int N, int C, double* out) {
    for (int c = 0; c < C; ++c) {
        double s = 0.0;
        for (int n = 0; n < N; ++n) s += table[(int64_t)n * C + c];
        const double m = s / (double)N;
        double q = 0.0;
        for (int n = 0; n < N; ++n) q += (table[(int64_t)n * C + c] - m) * (table[(int64_t)n * C + c] - m);
        out[c] = sqrt(q / (double)(N - 1));
    }
}

/* A = sig_mix + gamma_noise*gamma_std, B = mu_mix + beta_noise*beta_std   (maxstyle.py:172-185) */
static void coeffs(const double* mu, const double* sig, const int64_t* perm, const double* lmda, const double* gamma_noise,
                   const double* beta_noise, const double* gamma_std, const double* beta_std, int N, int C, int flags, double* A,
                   double* B) {
    const int mix = flags & 1, no_noise = flags & 2, no_clamp = flags & 8;
    for (int n = 0; n < N; ++n) {
        for (int c = 0; c < C; ++c) {
            const int64_t i = (int64_t)n * C + c;
            double sg = sig[i], m = mu[i];
            if (mix) {
                const double l = no_clamp ? lmda[n] : clamp01(lmda[n]);                 /* :173 */
                const int64_t j = perm[n] * C + c;                                       /* :174 */
                sg = sig[i] * (1.0 - l) + sig[j] * l;                                    /* :175 */
                m = mu[i] * (1.0 - l) + mu[j] * l;                                       /* :176 */
            }
            if (!no_noise) {
                sg += gamma_noise[i] * gamma_std[c];                                     /* :184 */
                m += beta_noise[i] * beta_std[c];                                        /* :185 */
            }
            A[i] = sg;
            B[i] = m;
        }
    }
}

/* y = A * (x - mu)/sig + B.  gamma_std / beta_std are outputs when (flags & 4), inputs otherwise.  Returns 0, or 1 on bad sizes. */
int ms_oracle_forward(const double* x, int N, int C, int64_t M, double eps, const int64_t* perm, const double* lmda,
                      const double* gamma_noise, const double* beta_noise, double* gamma_std, double* beta_std, int flags,
                      double* y, double* mu, double* sig, double* A) {
    if (N < 2 || C < 1 || M < 2) return 1;
    double* B = (double*)malloc(sizeof(double) * (size_t)N * C);
    if (!B) return 1;
    ms_oracle_stats(x, N, C, M, eps, mu, sig);
    if (flags & 4) {
        ms_oracle_batch_std(sig, N, C, gamma_std);
        ms_oracle_batch_std(mu, N, C, beta_std);
    }
    coeffs(mu, sig, perm, lmda, gamma_noise, beta_noise, gamma_std, beta_std, N, C, flags, A, B);
    for (int64_t p = 0; p < (int64_t)N * C; ++p)
        for (int64_t i = 0; i < M; ++i) y[p * M + i] = A[p] * ((x[p * M + i] - mu[p]) / sig[p]) + B[p];   /* :161, :184-185 */
    free(B);
    return 0;
}

/* dx = dy*A/sig;  dA = sum dy*xhat, dB = sum dy;  d_gamma = dA*gamma_std, d_beta = dB*beta_std;
 * d_lmda = [0 <= lmda <= 1] * sum_c (dA*(sig[perm]-sig) + dB*(mu[perm]-mu)) */
int ms_oracle_backward(const double* dy, const double* x, int N, int C, int64_t M, const int64_t* perm, const double* lmda,
                       const double* gamma_std, const double* beta_std, int flags, const double* mu, const double* sig,
                       const double* A, double* dx, double* d_gamma, double* d_beta, double* d_lmda) {
    if (N < 2 || C < 1 || M < 2) return 1;
    const int mix = flags & 1, no_noise = flags & 2, no_clamp = flags & 8;
    for (int n = 0; n < N; ++n) {
        double dl = 0.0;
        for (int c = 0; c < C; ++c) {
            const int64_t p = (int64_t)n * C + c;
            double dA = 0.0, dB = 0.0;
            for (int64_t i = 0; i < M; ++i) {
                const double g = dy[p * M + i];
                dA += g * ((x[p * M + i] - mu[p]) / sig[p]);
                dB += g;
                dx[p * M + i] = g * (A[p] / sig[p]);
            }
            d_gamma[p] = no_noise ? 0.0 : dA * gamma_std[c];
            d_beta[p] = no_noise ? 0.0 : dB * beta_std[c];
            if (mix) {
                const int64_t j = perm[n] * C + c;
                dl += dA * (sig[j] - sig[p]) + dB * (mu[j] - mu[p]);
            }
        }
        const int inside = no_clamp || (lmda[n] >= 0.0 && lmda[n] <= 1.0);
        d_lmda[n] = (mix && inside) ? dl : 0.0;
    }
    return 0;
}
