// Minimal LAPACK symbols needed to link IPX without a system LAPACK.
//
// The reference declares dpotrf_/dpotrs_/dtrcon_ in src/lapack.cc:16-23 and
// uses them only in the dense-column branch of DiagonalPrecond
// (src/diagonal_precond.cc:48-102,133-149). Only the argument combinations the
// reference issues are supported: uplo == 'L', column-major, 32-bit ints.

#include <cmath>
#include <vector>

// Hidden visibility: these three cover only what IPX calls; an application that also loads a
// real LAPACK (HiGHS, numpy ...) must neither pick them up nor have its own interposed here.
#define IPX_LAPACK_LOCAL __attribute__((visibility("hidden")))

extern "C" {

IPX_LAPACK_LOCAL void dpotrf_(const char* uplo, const int* n, double* a, const int* lda,
             int* info) {
    const int N = *n, LDA = *lda;
    *info = 0;
    if (*uplo != 'L' && *uplo != 'l') { *info = -1; return; }
    for (int j = 0; j < N; j++) {
        double d = a[j + (long)j * LDA];
        for (int k = 0; k < j; k++) {
            const double l = a[j + (long)k * LDA];
            d -= l * l;
        }
        if (!(d > 0.0)) { *info = j + 1; return; }
        d = std::sqrt(d);
        a[j + (long)j * LDA] = d;
        for (int i = j + 1; i < N; i++) {
            double s = a[i + (long)j * LDA];
            for (int k = 0; k < j; k++)
                s -= a[i + (long)k * LDA] * a[j + (long)k * LDA];
            a[i + (long)j * LDA] = s / d;
        }
    }
}

IPX_LAPACK_LOCAL void dpotrs_(const char* uplo, const int* n, const int* nrhs, const double* a,
             const int* lda, double* b, const int* ldb, int* info) {
    const int N = *n, LDA = *lda, LDB = *ldb;
    *info = 0;
    if (*uplo != 'L' && *uplo != 'l') { *info = -1; return; }
    for (int r = 0; r < *nrhs; r++) {
        double* x = b + (long)r * LDB;
        for (int i = 0; i < N; i++) {  // L z = b
            double s = x[i];
            for (int k = 0; k < i; k++) s -= a[i + (long)k * LDA] * x[k];
            x[i] = s / a[i + (long)i * LDA];
        }
        for (int i = N - 1; i >= 0; i--) {  // L' x = z
            double s = x[i];
            for (int k = i + 1; k < N; k++) s -= a[k + (long)i * LDA] * x[k];
            x[i] = s / a[i + (long)i * LDA];
        }
    }
}

// Reciprocal condition number of a lower triangular matrix in the 1-norm,
// computed from the explicit inverse (O(n^3); n <= 1000 dense columns).
IPX_LAPACK_LOCAL void dtrcon_(const char* norm, const char* uplo, const char* diag, const int* n,
             const double* a, const int* lda, double* rcond, double* work,
             int* iwork, int* info) {
    (void)norm; (void)work; (void)iwork;
    const int N = *n, LDA = *lda;
    *info = 0;
    *rcond = 0.0;
    if (*uplo != 'L' && *uplo != 'l') { *info = -2; return; }
    const bool unit = (*diag == 'U' || *diag == 'u');
    double anorm = 0.0, inorm = 0.0;
    std::vector<double> x(N);
    for (int j = 0; j < N; j++) {
        double s = 0.0;
        for (int i = j; i < N; i++)
            s += std::abs(i == j && unit ? 1.0 : a[i + (long)j * LDA]);
        anorm = std::max(anorm, s);
        for (int i = 0; i < N; i++) x[i] = (i == j) ? 1.0 : 0.0;
        double colsum = 0.0;
        for (int i = j; i < N; i++) {
            double v = x[i];
            for (int k = j; k < i; k++) v -= a[i + (long)k * LDA] * x[k];
            if (!unit) v /= a[i + (long)i * LDA];
            x[i] = v;
            colsum += std::abs(v);
        }
        inorm = std::max(inorm, colsum);
    }
    if (anorm > 0.0 && inorm > 0.0) *rcond = 1.0 / (anorm * inorm);
}

}  // extern "C"
