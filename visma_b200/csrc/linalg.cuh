// linalg.cuh — the small dense solves the reference delegates to Eigen, as device code so the whole ICP
// loop stays on the GPU (no per-iteration host round trip):
//   6x6 / 4x4 SPD solve with determinant guard  <- SolveLinearSystem (O3D/src/Core/Utility/Eigen.cpp:35-56)
//   Euler ZYX compose                           <- TransformVector6dToMatrix4d (Utility/Eigen.cpp:58-68)
//   3x3 SVD + Kabsch/Umeyama                    <- Eigen::umeyama (O3D/3rdparty/Eigen/Eigen/src/Geometry/Umeyama.h:93-162)
// All double precision, executed by one thread per ICP problem.
#pragma once

#include "common.cuh"

namespace vb {

#ifdef __CUDACC__

__device__ inline void mat4_identity(double *T) {
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// C = A * B (row-major 4x4); C may alias A or B
__device__ inline void mat4_mul(const double *A, const double *B, double *C) {
    double R[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
            R[4 * r + c] = s;
        }
    for (int i = 0; i < 16; i++) C[i] = R[i];
}

// Both factorizations below pivot, but are written with compile-time indices only (every loop unrolled, a
// row / column exchange expressed as predicated element swaps over the static candidates): the matrices stay
// in registers instead of local memory, which is what made the one-thread solve the longest part of k_solve.
__device__ __forceinline__ void cswap(bool c, double &a, double &b) {
    const double t = a;
    a = c ? b : a;
    b = c ? t : b;
}

// x = A^-1 b for symmetric A by LDL^T with diagonal pivoting (what Eigen's A.ldlt().solve(b) computes), and
// det A = the product of D (a symmetric exchange of rows and columns leaves the determinant's sign alone): the
// reference's guard |det| < 1e-6 (Utility/Eigen.cpp:41-52) needs no second factorization.  One reciprocal per
// pivot (6 divisions in all: a division is ~30 dependent instructions, and this runs on ONE thread between two
// correspondence passes); a zero pivot means singular: det = 0, and its column is left alone.  (A warp-cooperative
// version — a row per lane, pivot search and row exchanges through shuffles — was measured slower: 11.8 vs 10.5 us for
// the whole k_solve; the shuffle round trips cost more than the selects they replace.)
template <int N>
__device__ __forceinline__ double ldlt_solve(const double *A, const double *b, double *x) {
    double M[N * N], y[N], inv[N];
    int piv[N];
    double det = 1.0;
#pragma unroll
    for (int i = 0; i < N * N; i++) M[i] = A[i];
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = b[i];
#pragma unroll
    for (int k = 0; k < N; k++) {
        int p = k;
        double best = fabs(M[N * k + k]);
#pragma unroll
        for (int i = k + 1; i < N; i++) {
            const double v = fabs(M[N * i + i]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        // symmetric exchange of rows and columns k <-> p, and of the right-hand side (only the trailing block and
        // the finished columns of L in rows k, p are live)
#pragma unroll
        for (int i = k + 1; i < N; i++) {
            const bool sw = p == i;
#pragma unroll
            for (int c = 0; c < N; c++) cswap(sw, M[N * k + c], M[N * i + c]);
#pragma unroll
            for (int r = k; r < N; r++) cswap(sw, M[N * r + k], M[N * r + i]);
            cswap(sw, y[k], y[i]);
        }
        const double d = M[N * k + k];
        det *= d;
        inv[k] = d != 0.0 ? 1.0 / d : 0.0;
        if (d != 0.0) {
            double l[N];
#pragma unroll
            for (int i = k + 1; i < N; i++) { l[i] = M[N * i + k]; M[N * i + k] = l[i] * inv[k]; }
#pragma unroll
            for (int i = k + 1; i < N; i++) {
#pragma unroll
                for (int j = k + 1; j <= i; j++) {
                    M[N * i + j] -= M[N * i + k] * l[j];
                    M[N * j + i] = M[N * i + j];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) y[i] -= M[N * i + j] * y[j];
    }
#pragma unroll
    for (int i = 0; i < N; i++) y[i] *= inv[i];
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
#pragma unroll
        for (int j = i + 1; j < N; j++) y[i] -= M[N * j + i] * y[j];
    }
    // undo the exchanges, last first
#pragma unroll
    for (int k = N - 1; k >= 0; k--) {
#pragma unroll
        for (int i = k + 1; i < N; i++) cswap(piv[k] == i, y[k], y[i]);
    }
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = y[i];
    return det;
}

// SolveLinearSystem(JTJ, -JTr): false when |det| < 1e-6 or non-finite (Utility/Eigen.cpp:41-52)
template <int N>
__device__ inline bool solve_normal_equations(const double *JTJ, const double *JTr, double *x) {
    double nb[N];
    for (int i = 0; i < N; i++) nb[i] = -JTr[i];
    const double det = ldlt_solve<N>(JTJ, nb, x);
    return !(fabs(det) < 1e-6 || isnan(det) || isinf(det));
}

// R = Rz(x2) Ry(x1) Rx(x0), t = x3..5 (Utility/Eigen.cpp:58-68)
__device__ inline void vec6_to_T(const double *x, double *T) {
    double sa, ca, sb, cb, sg, cg;
    sincos(x[0], &sa, &ca);
    sincos(x[1], &sb, &cb);
    sincos(x[2], &sg, &cg);
    mat4_identity(T);
    T[0] = cg * cb;  T[1] = cg * sb * sa - sg * ca;  T[2] = cg * sb * ca + sg * sa;   T[3] = x[3];
    T[4] = sg * cb;  T[5] = sg * sb * sa + cg * ca;  T[6] = sg * sb * ca - cg * sa;   T[7] = x[4];
    T[8] = -sb;      T[9] = cb * sa;                 T[10] = cb * ca;                 T[11] = x[5];
}

// rotation by theta about unit axis g (Rodrigues) + translation
__device__ inline void axis_angle_to_T(double theta, const double *g, const double *t, double *T) {
    double s, c;
    sincos(theta, &s, &c);
    double v = 1.0 - c;
    mat4_identity(T);
    T[0] = c + g[0] * g[0] * v;        T[1] = g[0] * g[1] * v - g[2] * s; T[2] = g[0] * g[2] * v + g[1] * s;  T[3] = t[0];
    T[4] = g[1] * g[0] * v + g[2] * s; T[5] = c + g[1] * g[1] * v;        T[6] = g[1] * g[2] * v - g[0] * s;  T[7] = t[1];
    T[8] = g[2] * g[0] * v - g[1] * s; T[9] = g[2] * g[1] * v + g[0] * s; T[10] = c + g[2] * g[2] * v;        T[11] = t[2];
}

__device__ inline double det3(const double *A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// Rotation R maximising tr(R * Sigma^T)... i.e. the Kabsch/Umeyama rotation U S V^T for Sigma = U D V^T
// with S = diag(1,1,sign(det U det V)) (Umeyama.h:131-143).  One-sided Jacobi on the columns of Sigma
// gives V and the scaled left vectors; the third left vector is rebuilt as u0 x u1 with V forced to
// det +1, which IS the sign correction (R = [u0 u1 u0xu1] V^T).  sv (nullable) receives D . S.
__device__ inline void kabsch_rotation(const double *Sigma, double *R, double *sv_dot_S) {
    double B[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) B[i] = Sigma[i];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double a = 0, b = 0, g = 0;
                for (int r = 0; r < 3; r++) {
                    a += B[3 * r + p] * B[3 * r + p];
                    b += B[3 * r + q] * B[3 * r + q];
                    g += B[3 * r + p] * B[3 * r + q];
                }
                if (fabs(g) <= 1e-300) continue;
                double rel = fabs(g) / sqrt(a * b + 1e-300);
                if (rel > off) off = rel;
                double zeta = (b - a) / (2.0 * g);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < 3; r++) {
                    double bp = B[3 * r + p], bq = B[3 * r + q];
                    B[3 * r + p] = c * bp - s * bq;
                    B[3 * r + q] = s * bp + c * bq;
                    double wp = W[3 * r + p], wq = W[3 * r + q];
                    W[3 * r + p] = c * wp - s * wq;
                    W[3 * r + q] = s * wp + c * wq;
                }
            }
        if (off < 1e-15) break;
    }
    double nrm[3];
    int ord[3] = {0, 1, 2};
    for (int c = 0; c < 3; c++) nrm[c] = sqrt(B[c] * B[c] + B[3 + c] * B[3 + c] + B[6 + c] * B[6 + c]);
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (nrm[ord[j]] > nrm[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double U[9], V[9];
    for (int k = 0; k < 3; k++)
        for (int r = 0; r < 3; r++) V[3 * r + k] = W[3 * r + ord[k]];
    // force det V = +1 by flipping the direction paired with the smallest singular value
    double flip = det3(V) < 0 ? -1.0 : 1.0;
    for (int r = 0; r < 3; r++) V[3 * r + 2] *= flip;
    const double tiny = nrm[ord[0]] * 1e-14;
    bool ok0 = nrm[ord[0]] > 0.0, ok1 = nrm[ord[1]] > tiny;
    if (!ok0) {  // Sigma == 0: any rotation is optimal; Eigen returns U = V = I
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (sv_dot_S) *sv_dot_S = 0.0;
        return;
    }
    for (int r = 0; r < 3; r++) U[3 * r + 0] = B[3 * r + ord[0]] / nrm[ord[0]];
    if (ok1) {
        for (int r = 0; r < 3; r++) U[3 * r + 1] = B[3 * r + ord[1]] / nrm[ord[1]];
    } else {  // rank 1: pick any unit vector orthogonal to u0
        double u0[3] = {U[0], U[3], U[6]};
        int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[m] = 1.0;
        double d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
        double u1[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
        double l = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int r = 0; r < 3; r++) U[3 * r + 1] = u1[r] / l;
    }
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            R[3 * r + c] = U[3 * r + 0] * V[3 * c + 0] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
    if (sv_dot_S) {
        // sign of the third singular value relative to (u0 x u1, v2): Sigma v2 = +-s2 u2
        double sv2 = 0.0;
        for (int r = 0; r < 3; r++) {
            double av = Sigma[3 * r + 0] * V[0 + 2] + Sigma[3 * r + 1] * V[3 + 2] + Sigma[3 * r + 2] * V[6 + 2];
            sv2 += av * U[3 * r + 2];
        }
        *sv_dot_S = nrm[ord[0]] + nrm[ord[1]] + sv2;
    }
}

#endif  // __CUDACC__

}  // namespace vb
