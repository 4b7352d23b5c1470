// vfk_nullspace.cuh -- orthonormal basis of null(J) for the reference's nullspace interface.
//
// scripts/nullspace:75-107 builds B = I - pinv(J) J (LAPACK SVD) and takes the left singular vectors of B^T whose
// singular value is >= 1e-8: an orthonormal basis of null(J), k = N - rank(J) vectors.  Forming B through the normal
// equations J^T (J J^T)^-1 J loses cond(J)^2 * eps; a Householder QR of J^T (N x 6) loses only cond(J) * eps and needs
// no damping: J^T = Q R, Q = H_0 ... H_5 orthogonal N x N, and columns 6 .. N-1 of Q are orthogonal to range(J^T),
// i.e. they span null(J) (exactly N - 6 of them when J has full row rank; at a rank-deficient posture they are still
// null vectors, the reference would find one more).  The reflectors follow LAPACK's dgeqr2 / dlarfg convention
// (beta = -sign(alpha) |x|, v_0 = 1, H = I - tau v v^T), so numpy.linalg.qr(J.T, mode="complete") gives the same
// vectors up to rounding -- that is what oracle/batch.py:ns_basis pins this against.  The projector needs no basis at
// all: (I - pinv(J) J) x = Q diag(0, .., 0, 1, .., 1) Q^T x.
//
// Rare path (control mode / undamped projector), any N <= 17: rolled loops over thread-local arrays, compiled once per
// translation unit and called -- it costs the throughput instantiations neither registers nor instruction-cache lines.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace vfk {

// Factor a = J^T, n x 6 row-major (a[r * 6 + c] = J[c][r]), in place: on return the reflector vectors sit below the
// diagonal (v_j[j] = 1 implied) and tau[j] holds their scales; tau[j] = 0 means H_j = I (dlarfg's zero-tail case, and every
// j >= n).
static __device__ __noinline__ void ns_qr_factor(double* a, int n, double* tau) {
    for (int j = 0; j < 6; ++j) {
        tau[j] = 0.0;
        if (j >= n) continue;
        const double alpha = a[j * 6 + j];
        double xn2 = 0.0;
        for (int r = j + 1; r < n; ++r) xn2 = fma(a[r * 6 + j], a[r * 6 + j], xn2);
        if (xn2 == 0.0) continue;
        const double beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
        tau[j] = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        for (int r = j + 1; r < n; ++r) a[r * 6 + j] *= sc;
        a[j * 6 + j] = beta;
        for (int c = j + 1; c < 6; ++c) {                          // apply H_j to the remaining columns
            double w = a[j * 6 + c];
            for (int r = j + 1; r < n; ++r) w = fma(a[r * 6 + j], a[r * 6 + c], w);
            w *= tau[j];
            a[j * 6 + c] -= w;
            for (int r = j + 1; r < n; ++r) a[r * 6 + c] = fma(-a[r * 6 + j], w, a[r * 6 + c]);
        }
    }
}

__device__ __forceinline__ void ns_reflect(const double* a, double tau_j, int j, int n, double* w) {
    if (tau_j == 0.0) return;
    double d = w[j];
    for (int r = j + 1; r < n; ++r) d = fma(a[r * 6 + j], w[r], d);
    d *= tau_j;
    w[j] -= d;
    for (int r = j + 1; r < n; ++r) w[r] = fma(-a[r * 6 + j], d, w[r]);
}

// w <- Q e_col = H_0 ( ... (H_5 e_col)): column `col` of Q; col >= 6 is a unit vector of null(J).
static __device__ __noinline__ void ns_qr_column(const double* a, const double* tau, int n, int col, double* w) {
    for (int r = 0; r < n; ++r) w[r] = (r == col) ? 1.0 : 0.0;
    for (int j = 5; j >= 0; --j) ns_reflect(a, tau[j], j, n, w);
}

// w <- (I - pinv(J) J) w = Q diag(0 x 6, 1 ...) Q^T w: the orthogonal projector onto null(J), without forming a basis.
static __device__ __noinline__ void ns_qr_project(const double* a, const double* tau, int n, double* w) {
    for (int j = 0; j < 6; ++j) ns_reflect(a, tau[j], j, n, w);          // Q^T w = H_5 ( ... (H_0 w))
    for (int r = 0; r < 6 && r < n; ++r) w[r] = 0.0;
    for (int j = 5; j >= 0; --j) ns_reflect(a, tau[j], j, n, w);
}

// vectors of the basis the four-float control interface can address: min(4, N - 6) (scripts/nullspace:113)
__host__ __device__ constexpr int ns_ctrl_vectors(int n) { return n > 6 ? (n - 6 < 4 ? n - 6 : 4) : 0; }
__host__ __device__ constexpr int ns_null_vectors(int n) { return n > 6 ? n - 6 : 0; }

}  // namespace vfk
