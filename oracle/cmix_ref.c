/*
 * cmix_ref.c — C restatement of the reference's stage 2+3 in the REFERENCE'S OWN OPERATION ORDER.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py): used by tests/ as the checker at sizes
 * where the numpy oracle is too slow, and by bench.py as the `cpu_baseline` / `--impl reference` arm
 * (kind "port": Julia is not installed in this image, so the reference itself cannot run).
 * Nothing in the product package links or calls this file.
 *
 * Follows hsgg/SphericalFourierBesselDecompositions.jl v0.5.19:
 *   sfbo_wigner3j000        src/windows.jl:421-431   closed form through lgamma, recomputed per (element, L1)
 *   sfbo_calc_wrl_wrl       src/windows.jl:682-696   scalar loops over L1, j, i, M1 (unthreaded in the reference)
 *   sfbo_cmixlnnLNN         src/windows.jl:613-627   gg1, gg2, then per L1: w3j^2 * (gg1' * W[:,:,L1] * gg2)
 *   sfbo_calc_cmix_rows     src/windows.jl:700-746   (i,i') loop with the N<->N' partner, div2Lp1, interchange;
 *                                                    dynamic self-scheduling over work batches like mybroadcast
 *                                                    (src/MyBroadcast.jl:76-147), here OpenMP schedule(dynamic)
 * Layouts are Julia's: W[i + nr*(j + nr*L1)], G[r + nr*((n-1) + nmax*l)], lnn[3*i + {0,1,2}] (1-based n).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

double sfbo_wigner3j000(int64_t l, int64_t lp, int64_t L) {
    if (!(llabs(l - lp) <= L && L <= l + lp)) return 0.0;
    const int64_t J = l + lp + L;
    if (J % 2 != 0) return 0.0;
    const double sign = ((J / 2) % 2) ? -1.0 : 1.0;
    return sign * exp(0.5 * lgamma(1.0 + J - 2 * l) + 0.5 * lgamma(1.0 + J - 2 * lp) + 0.5 * lgamma(1.0 + J - 2 * L) -
                      0.5 * lgamma(1.0 + J + 1) + lgamma(1.0 + J / 2) - lgamma(1.0 + J / 2 - l) -
                      lgamma(1.0 + J / 2 - lp) - lgamma(1.0 + J / 2 - L));
}

/* W1, W2: nr x lmsize complex, column-major, m-fast column order (idx = m + l(l+1)/2). */
void sfbo_calc_wrl_wrl(const double complex* W1, const double complex* W2, int64_t nr, int64_t LMAX, double* W) {
    for (int64_t L1 = 0; L1 <= LMAX; ++L1)
        for (int64_t j = 0; j < nr; ++j)
            for (int64_t i = 0; i < nr; ++i) {
                int64_t lm = L1 * (L1 + 1) / 2;
                double s = creal(W1[i + nr * lm] * conj(W2[j + nr * lm]));
                for (int64_t M1 = 1; M1 <= L1; ++M1) {
                    lm = M1 + L1 * (L1 + 1) / 2;
                    s += 2 * creal(W1[i + nr * lm] * conj(W2[j + nr * lm]));
                }
                W[i + nr * (j + nr * L1)] = s;
            }
}

static double cmixlnnLNN(int64_t l, int64_t n, int64_t n_, int64_t L, int64_t N, int64_t N_, const double* W,
                         const double* G, int64_t nr, int64_t nmax, double* gg1, double* gg2) {
    const double* g_nl = G + nr * ((n - 1) + nmax * l);
    const double* g_NL = G + nr * ((N - 1) + nmax * L);
    const double* g_n_l = G + nr * ((n_ - 1) + nmax * l);
    const double* g_N_L = G + nr * ((N_ - 1) + nmax * L);
    for (int64_t r = 0; r < nr; ++r) {
        gg1[r] = g_nl[r] * g_NL[r];
        gg2[r] = g_n_l[r] * g_N_L[r];
    }
    double mix = 0.0;
    const int64_t lo = llabs(l - L);
    for (int64_t L1 = lo; L1 <= l + L; L1 += 2) {
        const double w3j = sfbo_wigner3j000(l, L, L1);
        const double* WL = W + nr * nr * L1;
        double q = 0.0; /* gg1' * W * gg2 */
        for (int64_t j = 0; j < nr; ++j) {
            const double* col = WL + nr * j;
            double t = 0.0;
            for (int64_t i = 0; i < nr; ++i) t += gg1[i] * col[i];
            q += t * gg2[j];
        }
        mix += w3j * w3j * q;
    }
    return mix * (2 * L + 1) / (4 * M_PI);
}

double sfbo_cmixlnnLNN(int64_t l, int64_t n, int64_t n_, int64_t L, int64_t N, int64_t N_, const double* W,
                       const double* G, int64_t nr, int64_t nmax) {
    double* gg = (double*)malloc(2 * nr * sizeof(double));
    const double v = cmixlnnLNN(l, n, n_, L, N, N_, W, G, nr, nmax, gg, gg + nr);
    free(gg);
    return v;
}

/* out[k + nrows*(c)] for rows[k] (1-based indices into lnn) and columns col_lo..col_hi (1-based, inclusive). */
void sfbo_calc_cmix_rows(const int64_t* lnn, const int64_t* rows, int64_t nrows, int64_t col_lo, int64_t col_hi,
                         const double* G, int64_t nr, int64_t nmax, const double* W, int32_t div2Lp1,
                         int32_t interchange, int32_t nthreads, double* out) {
    const int64_t ncols = col_hi - col_lo + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double* gg = (double*)malloc(2 * nr * sizeof(double));
#pragma omp for schedule(dynamic, 16)
        for (int64_t idx = 0; idx < nrows * ncols; ++idx) {
            const int64_t k = idx % nrows, c = idx / nrows;
            const int64_t i = rows[k] - 1, ip = col_lo - 1 + c;
            const int64_t l = lnn[3 * i], n = lnn[3 * i + 1], n_ = lnn[3 * i + 2];
            const int64_t L = lnn[3 * ip];
            int64_t N = lnn[3 * ip + 1], N_ = lnn[3 * ip + 2];
            if (interchange) {
                const int64_t t = N;
                N = N_;
                N_ = t;
            }
            double mix = cmixlnnLNN(l, n, n_, L, N, N_, W, G, nr, nmax, gg, gg + nr);
            if (!interchange && N != N_) mix += cmixlnnLNN(l, n, n_, L, N_, N, W, G, nr, nmax, gg, gg + nr);
            if (div2Lp1) mix /= (2 * L + 1);
            out[k + nrows * c] = mix;
        }
        free(gg);
    }
}

int32_t sfbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
