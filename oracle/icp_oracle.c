/* oracle/icp_oracle.c — plain-C restatement of the reference's correspondence search, estimators
 * and ICP loop.  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/oracle.h for the rules and the
 * parity status).  Every function cites the reference lines it follows; paths are relative to
 * /root/reference, O3D = thirdparty/Open3D.
 *
 * Compiled with -ffp-contract=off so the double arithmetic is the plain IEEE sequence written here.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================================================
 * Nearest-neighbour index.  The reference builds a FLANN single KD-tree per RegistrationICP call
 * (O3D/src/Core/Geometry/KDTreeFlann.cpp:191-208) and asks it for the 1 nearest point within a radius
 * (KDTreeFlann.cpp:165-189).  The KD-tree is exact, so any exact index returns the same neighbour; this
 * one is a sorted uniform grid with cell >= radius searched over the 27 surrounding cells.
 * ================================================================================================= */
struct vo_index {
    const double *xyz; /* borrowed: caller keeps the target alive */
    int64_t n;
    double cell, inv_cell;
    double lo[3];
    int64_t dim[3];
    int64_t ncell_used;
    int64_t *keys;   /* sorted unique cell keys */
    int64_t *start;  /* ncell_used+1 offsets into order */
    int32_t *order;  /* point indices sorted by (key, index) */
};

typedef struct { int64_t key; int32_t idx; } vo_kv;

static int kv_cmp(const void *a, const void *b) {
    const vo_kv *x = (const vo_kv *)a, *y = (const vo_kv *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static void cell_of(const vo_index *ix, const double *p, int64_t c[3]) {
    for (int a = 0; a < 3; a++) c[a] = (int64_t)floor((p[a] - ix->lo[a]) * ix->inv_cell);
}

vo_index *vo_index_create(const double *xyz, int64_t n, double cell) {
    if (n <= 0 || !(cell > 0.0)) return NULL;
    vo_index *ix = (vo_index *)calloc(1, sizeof(vo_index));
    ix->xyz = xyz; ix->n = n; ix->cell = cell; ix->inv_cell = 1.0 / cell;
    double hi[3];
    for (int a = 0; a < 3; a++) { ix->lo[a] = xyz[a]; hi[a] = xyz[a]; }
    for (int64_t i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            double v = xyz[3 * i + a];
            if (v < ix->lo[a]) ix->lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    for (int a = 0; a < 3; a++) ix->dim[a] = (int64_t)floor((hi[a] - ix->lo[a]) * ix->inv_cell) + 1;
    vo_kv *kv = (vo_kv *)malloc(sizeof(vo_kv) * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        int64_t c[3];
        cell_of(ix, xyz + 3 * i, c);
        kv[i].key = (c[2] * ix->dim[1] + c[1]) * ix->dim[0] + c[0];
        kv[i].idx = (int32_t)i;
    }
    qsort(kv, (size_t)n, sizeof(vo_kv), kv_cmp);
    ix->order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    ix->keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    ix->start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        ix->order[i] = kv[i].idx;
        if (i == 0 || kv[i].key != kv[i - 1].key) { ix->keys[u] = kv[i].key; ix->start[u] = i; u++; }
    }
    ix->start[u] = n;
    ix->ncell_used = u;
    free(kv);
    return ix;
}

void vo_index_destroy(vo_index *ix) {
    if (!ix) return;
    free(ix->order); free(ix->keys); free(ix->start); free(ix);
}

static int64_t find_key(const vo_index *ix, int64_t key) {
    int64_t lo = 0, hi = ix->ncell_used - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (ix->keys[mid] == key) return mid;
        if (ix->keys[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* flann::L2<double> for 3-D data (O3D/3rdparty/flann/algorithms/dist.h:150-177): the tail loop adds the
 * squared differences one at a time, i.e. ((dx*dx)+dy*dy)+dz*dz. */
static double l2_3(const double *a, const double *b) {
    double r = 0.0, d;
    d = a[0] - b[0]; r += d * d;
    d = a[1] - b[1]; r += d * d;
    d = a[2] - b[2]; r += d * d;
    return r;
}

static void knn1_one(const vo_index *ix, const double *q, double r2, int32_t *oi, double *od) {
    int64_t c[3];
    cell_of(ix, q, c);
    int32_t best = -1;
    double bd = r2; /* KNNRadiusResultSet starts with worst_dist = radius (result_set.h:529-545) */
    for (int64_t dz = -1; dz <= 1; dz++) {
        int64_t z = c[2] + dz;
        if (z < 0 || z >= ix->dim[2]) continue;
        for (int64_t dy = -1; dy <= 1; dy++) {
            int64_t y = c[1] + dy;
            if (y < 0 || y >= ix->dim[1]) continue;
            for (int64_t dx = -1; dx <= 1; dx++) {
                int64_t x = c[0] + dx;
                if (x < 0 || x >= ix->dim[0]) continue;
                int64_t u = find_key(ix, (z * ix->dim[1] + y) * ix->dim[0] + x);
                if (u < 0) continue;
                for (int64_t s = ix->start[u]; s < ix->start[u + 1]; s++) {
                    int32_t j = ix->order[s];
                    double d = l2_3(q, ix->xyz + 3 * (int64_t)j);
                    /* accept iff strictly closer (result_set.h:582 rejects dist >= worst_dist);
                     * equal distance -> lowest index (documented tie rule) */
                    if (d < bd || (d == bd && best >= 0 && j < best)) { bd = d; best = j; }
                }
            }
        }
    }
    *oi = best;
    *od = best >= 0 ? bd : 0.0;
}

int vo_knn1(const vo_index *ix, const double *q, int64_t nq, double radius, int32_t *out_idx,
            double *out_d2) {
    if (!ix || radius <= 0.0 || radius > ix->cell * (1.0 + 1e-12)) return -1;
    /* KDTreeFlann.cpp:185: the radius handed to FLANN is float(radius*radius) */
    double r2 = (double)(float)(radius * radius);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; i++) knn1_one(ix, q + 3 * i, r2, out_idx + i, out_d2 + i);
    return 0;
}

int vo_knn1_brute(const double *tgt, int64_t n, const double *q, int64_t nq, int32_t *out_idx,
                  double *out_d2) {
    for (int64_t i = 0; i < nq; i++) {
        int32_t best = -1;
        double bd = 0.0;
        for (int64_t j = 0; j < n; j++) {
            double d = l2_3(q + 3 * i, tgt + 3 * j);
            if (best < 0 || d < bd) { bd = d; best = (int32_t)j; }
        }
        out_idx[i] = best;
        out_d2[i] = bd;
    }
    return 0;
}

/* ===================================================================================================
 * Small dense linear algebra the reference gets from Eigen.
 * ================================================================================================= */
static void mat4_identity(double T[16]) {
    memset(T, 0, sizeof(double) * 16);
    T[0] = T[5] = T[10] = T[15] = 1.0;
}

static void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
            R[4 * r + c] = s;
        }
    memcpy(C, R, sizeof(R));
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    double R[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += A[3 * r + k] * B[3 * k + c];
            R[3 * r + c] = s;
        }
    memcpy(C, R, sizeof(R));
}

static double det3(const double A[9]) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}

/* JacobiSVD<Matrix3d>(A, FullU|FullV) stand-in: one-sided Jacobi (Hestenes) on the columns of A.
 * Returns A = U diag(s) V^T with s sorted descending, U and V orthogonal (either determinant sign,
 * like Eigen's).  Null directions of U are completed by cross products. */
static void svd3(const double A[9], double U[9], double s[3], double V[9]) {
    double B[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(B, A, sizeof(B));
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
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < 3; r++) {
                    double bp = B[3 * r + p], bq = B[3 * r + q];
                    B[3 * r + p] = c * bp - sn * bq;
                    B[3 * r + q] = sn * bp + c * bq;
                    double wp = W[3 * r + p], wq = W[3 * r + q];
                    W[3 * r + p] = c * wp - sn * wq;
                    W[3 * r + q] = sn * wp + c * wq;
                }
            }
        if (off < 1e-15) break;
    }
    double nrm[3];
    int ord[3] = {0, 1, 2};
    for (int c = 0; c < 3; c++)
        nrm[c] = sqrt(B[c] * B[c] + B[3 + c] * B[3 + c] + B[6 + c] * B[6 + c]);
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (nrm[ord[j]] > nrm[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double tiny = nrm[ord[0]] * 1e-14;
    int rank = 0;
    for (int k = 0; k < 3; k++) {
        int c = ord[k];
        s[k] = nrm[c];
        for (int r = 0; r < 3; r++) V[3 * r + k] = W[3 * r + c];
        if (nrm[c] > tiny && nrm[c] > 0.0) {
            for (int r = 0; r < 3; r++) U[3 * r + k] = B[3 * r + c] / nrm[c];
            rank = k + 1;
        }
    }
    /* complete U to an orthonormal basis where A is rank deficient */
    if (rank == 0) {
        double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(U, I, sizeof(I));
    } else if (rank == 1) {
        double u0[3] = {U[0], U[3], U[6]};
        int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2)
                                            : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[m] = 1.0;
        double d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
        double u1[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
        double l = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int r = 0; r < 3; r++) U[3 * r + 1] = u1[r] / l;
        rank = 2;
    }
    if (rank == 2) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

/* Eigen::umeyama (O3D/3rdparty/Eigen/Eigen/src/Geometry/Umeyama.h:93-162) on gathered correspondences.
 * src/dst are the 3xK matrices the reference builds at TransformationEstimation.cpp:52-57. */
int vo_estimate_p2p(const double *src, const double *tgt, const int32_t *corr, int64_t k,
                    int with_scaling, double T[16]) {
    mat4_identity(T);
    if (k <= 0) return 0; /* corres.empty() -> Identity (TransformationEstimation.cpp:51) */
    double one_over_n = 1.0 / (double)k;
    double sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (int64_t i = 0; i < k; i++) {
        const double *s = src + 3 * (int64_t)corr[2 * i], *d = tgt + 3 * (int64_t)corr[2 * i + 1];
        for (int a = 0; a < 3; a++) { sm[a] += s[a]; dm[a] += d[a]; }
    }
    for (int a = 0; a < 3; a++) { sm[a] *= one_over_n; dm[a] *= one_over_n; } /* Umeyama.h:118-119 */
    double sigma[9] = {0}, src_var = 0.0;
    for (int64_t i = 0; i < k; i++) {
        const double *s = src + 3 * (int64_t)corr[2 * i], *d = tgt + 3 * (int64_t)corr[2 * i + 1];
        double sd[3] = {s[0] - sm[0], s[1] - sm[1], s[2] - sm[2]};  /* :122-123 demeaning */
        double dd[3] = {d[0] - dm[0], d[1] - dm[1], d[2] - dm[2]};
        src_var += sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2];   /* :126 */
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) sigma[3 * r + c] += dd[r] * sd[c]; /* :129 */
    }
    src_var *= one_over_n;
    for (int e = 0; e < 9; e++) sigma[e] *= one_over_n;
    double U[9], sv[3], V[9];
    svd3(sigma, U, sv, V);                                           /* :131 */
    double S[3] = {1, 1, 1};
    if (det3(U) * det3(V) < 0) S[2] = -1;                            /* :139-140 */
    double US[9], Vt[9], R[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { US[3 * r + c] = U[3 * r + c] * S[c]; Vt[3 * r + c] = V[3 * c + r]; }
    mat3_mul(US, Vt, R);                                             /* :143 */
    double c = 1.0;
    if (with_scaling) c = 1.0 / src_var * (sv[0] * S[0] + sv[1] * S[1] + sv[2] * S[2]); /* :148 */
    for (int r = 0; r < 3; r++) {
        double rs = R[3 * r] * sm[0] + R[3 * r + 1] * sm[1] + R[3 * r + 2] * sm[2];
        T[4 * r + 3] = dm[r] - c * rs;                               /* :151-152 / :157-158 */
        for (int cc = 0; cc < 3; cc++) T[4 * r + cc] = c * R[3 * r + cc];
    }
    return 0;
}

double vo_rmse_p2p(const double *src, const double *tgt, const int32_t *corr, int64_t k) {
    if (k <= 0) return 0.0; /* src/constrained_ICP.cpp:17 */
    double err = 0.0;
    for (int64_t i = 0; i < k; i++) {
        const double *s = src + 3 * (int64_t)corr[2 * i], *d = tgt + 3 * (int64_t)corr[2 * i + 1];
        double x = s[0] - d[0], y = s[1] - d[1], z = s[2] - d[2];
        err += x * x + y * y + z * z; /* squaredNorm, :20 */
    }
    return sqrt(err / (double)k);
}

/* determinant by LU with partial pivoting (what MatrixXd::determinant() does for a 6x6) */
static double det_n(const double *A, int n) {
    double M[36];
    memcpy(M, A, sizeof(double) * (size_t)(n * n));
    double det = 1.0;
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (fabs(M[n * r + c]) > fabs(M[n * p + c])) p = r;
        if (M[n * p + c] == 0.0) return 0.0;
        if (p != c) {
            for (int k = 0; k < n; k++) { double t = M[n * c + k]; M[n * c + k] = M[n * p + k]; M[n * p + k] = t; }
            det = -det;
        }
        det *= M[n * c + c];
        for (int r = c + 1; r < n; r++) {
            double f = M[n * r + c] / M[n * c + c];
            for (int k = c; k < n; k++) M[n * r + k] -= f * M[n * c + k];
        }
    }
    return det;
}

/* A.ldlt().solve(b): robust Cholesky with diagonal pivoting (Utility/Eigen.cpp:45).  A symmetric n x n. */
static void ldlt_solve_n(const double *A, const double *b, double *x, int n) {
    double M[36], y[6];
    int perm[6];
    memcpy(M, A, sizeof(double) * (size_t)(n * n));
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++)
            if (fabs(M[n * i + i]) > fabs(M[n * p + p])) p = i;
        if (p != k) { /* symmetric row/column swap */
            for (int c = 0; c < n; c++) { double t = M[n * k + c]; M[n * k + c] = M[n * p + c]; M[n * p + c] = t; }
            for (int r = 0; r < n; r++) { double t = M[n * r + k]; M[n * r + k] = M[n * r + p]; M[n * r + p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        double d = M[n * k + k];
        if (d == 0.0) continue;
        for (int i = k + 1; i < n; i++) M[n * i + k] /= d; /* L column */
        for (int i = k + 1; i < n; i++)
            for (int j = k + 1; j <= i; j++) {
                M[n * i + j] -= M[n * i + k] * d * M[n * j + k];
                M[n * j + i] = M[n * i + j];
            }
    }
    for (int i = 0; i < n; i++) y[i] = b[perm[i]];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) y[i] -= M[n * i + j] * y[j];
    for (int i = 0; i < n; i++) y[i] = (M[n * i + i] != 0.0) ? y[i] / M[n * i + i] : 0.0;
    for (int i = n - 1; i >= 0; i--)
        for (int j = i + 1; j < n; j++) y[i] -= M[n * j + i] * y[j];
    for (int i = 0; i < n; i++) x[perm[i]] = y[i];
}

/* SolveLinearSystem(JTJ, -JTr) (Utility/Eigen.cpp:35-56): |det| < 1e-6 or non-finite -> no solution. */
static int solve_sys(const double *JTJ, const double *JTr, double *x, int n) {
    double det = det_n(JTJ, n);
    if (fabs(det) < 1e-6 || isnan(det) || isinf(det)) {
        for (int i = 0; i < n; i++) x[i] = 0.0;
        return 0;
    }
    double nb[6];
    for (int i = 0; i < n; i++) nb[i] = -JTr[i];
    ldlt_solve_n(JTJ, nb, x, n);
    return 1;
}

int vo_solve6(const double JTJ[36], const double JTr[6], double x[6]) { return solve_sys(JTJ, JTr, x, 6); }

/* TransformVector6dToMatrix4d (Utility/Eigen.cpp:58-68): R = Rz(x2) * Ry(x1) * Rx(x0), t = x3..5. */
void vo_vec6_to_T(const double x[6], double T[16]) {
    double ca = cos(x[0]), sa = sin(x[0]), cb = cos(x[1]), sb = sin(x[1]), cg = cos(x[2]), sg = sin(x[2]);
    double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca};
    double Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb};
    double Rz[9] = {cg, -sg, 0, sg, cg, 0, 0, 0, 1};
    double ZY[9], R[9];
    mat3_mul(Rz, Ry, ZY);
    mat3_mul(ZY, Rx, R);
    mat4_identity(T);
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[4 * r + c] = R[3 * r + c];
        T[4 * r + 3] = x[3 + r];
    }
}

/* TransformationEstimationPointToPlane::ComputeTransformation (TransformationEstimation.cpp:75-103) with
 * ComputeJTJandJTr (Utility/Eigen.cpp:137-182) accumulated sequentially (the reference's OpenMP merge
 * order is thread dependent; sums agree to rounding). */
int vo_estimate_p2plane(const double *src, const double *tgt, const double *nrm, const int32_t *corr,
                        int64_t k, double T[16]) {
    mat4_identity(T);
    if (k <= 0 || !nrm) return 0; /* :79-80 */
    double JTJ[36] = {0}, JTr[6] = {0};
    for (int64_t i = 0; i < k; i++) {
        const double *vs = src + 3 * (int64_t)corr[2 * i];
        const double *vt = tgt + 3 * (int64_t)corr[2 * i + 1];
        const double *nt = nrm + 3 * (int64_t)corr[2 * i + 1];
        double r = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] + (vs[2] - vt[2]) * nt[2]; /* :87 */
        double J[6] = {vs[1] * nt[2] - vs[2] * nt[1], vs[2] * nt[0] - vs[0] * nt[2],
                       vs[0] * nt[1] - vs[1] * nt[0], nt[0], nt[1], nt[2]};                      /* :88-89 */
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) JTJ[6 * a + b] += J[a] * J[b];
            JTr[a] += J[a] * r;
        }
    }
    double x[6];
    if (!solve_sys(JTJ, JTr, x, 6)) return 0; /* is_success false -> Identity (:102) */
    vo_vec6_to_T(x, T);
    return 0;
}

/* Gravity-constrained 4-DoF step (north-star extension).  x = [theta, tx, ty, tz];
 * J4 = [(vs x nt).g ; nt]; solve the 4x4 normal equations with the same det guard; update.R =
 * AngleAxis(theta, g) (Rodrigues), update.t = x1..3. */
int vo_estimate_p2plane_gravity(const double *src, const double *tgt, const double *nrm,
                                const int32_t *corr, int64_t k, const double g_in[3], double T[16]) {
    mat4_identity(T);
    if (k <= 0 || !nrm) return 0;
    double gl = sqrt(g_in[0] * g_in[0] + g_in[1] * g_in[1] + g_in[2] * g_in[2]);
    if (!(gl > 0.0)) return -1;
    double g[3] = {g_in[0] / gl, g_in[1] / gl, g_in[2] / gl};
    double A[16] = {0}, b[4] = {0};
    for (int64_t i = 0; i < k; i++) {
        const double *vs = src + 3 * (int64_t)corr[2 * i];
        const double *vt = tgt + 3 * (int64_t)corr[2 * i + 1];
        const double *nt = nrm + 3 * (int64_t)corr[2 * i + 1];
        double r = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] + (vs[2] - vt[2]) * nt[2];
        double cx = vs[1] * nt[2] - vs[2] * nt[1], cy = vs[2] * nt[0] - vs[0] * nt[2],
               cz = vs[0] * nt[1] - vs[1] * nt[0];
        double J[4] = {cx * g[0] + cy * g[1] + cz * g[2], nt[0], nt[1], nt[2]};
        for (int a = 0; a < 4; a++) {
            for (int c = 0; c < 4; c++) A[4 * a + c] += J[a] * J[c];
            b[a] += J[a] * r;
        }
    }
    double x[4];
    if (!solve_sys(A, b, x, 4)) return 0;
    double th = x[0], c = cos(th), s = sin(th), v = 1.0 - c;
    double R[9] = {c + g[0] * g[0] * v,        g[0] * g[1] * v - g[2] * s, g[0] * g[2] * v + g[1] * s,
                   g[1] * g[0] * v + g[2] * s, c + g[1] * g[1] * v,        g[1] * g[2] * v - g[0] * s,
                   g[2] * g[0] * v - g[1] * s, g[2] * g[1] * v + g[0] * s, c + g[2] * g[2] * v};
    for (int r = 0; r < 3; r++) {
        for (int cc = 0; cc < 3; cc++) T[4 * r + cc] = R[3 * r + cc];
        T[4 * r + 3] = x[1 + r];
    }
    return 0;
}

/* ===================================================================================================
 * ICP loop
 * ================================================================================================= */
/* PointCloud::Transform (O3D/src/Core/Geometry/PointCloud.cpp:75-87): p <- (T [p,1])_xyz, n <- (T [n,0])_xyz */
static void transform_cloud(double *xyz, double *nrm, int64_t m, const double T[16]) {
    for (int64_t i = 0; i < m; i++) {
        double *p = xyz + 3 * i, q[3];
        for (int r = 0; r < 3; r++)
            q[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3] * 1.0;
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    }
    if (nrm)
        for (int64_t i = 0; i < m; i++) {
            double *p = nrm + 3 * i, q[3];
            for (int r = 0; r < 3; r++)
                q[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2];
            p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
        }
}

static int is_identity(const double T[16]) {
    /* Eigen isIdentity(prec=1e-12): off-diagonals negligible vs 1, diagonals approx 1 */
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            double v = T[4 * r + c];
            if (r == c) { if (fabs(v - 1.0) > 1e-12 * fmin(fabs(v), 1.0)) return 0; }
            else if (fabs(v) > 1e-12) return 0;
        }
    return 1;
}

/* GetRegistrationResultAndCorrespondences (Registration.cpp:41-96) */
static int64_t corr_pass(const vo_index *ix, const double *pcd, int64_t m, double max_dist,
                         int32_t *idx, double *d2, int32_t *corr, double *fitness, double *rmse) {
    vo_knn1(ix, pcd, m, max_dist, idx, d2);
    int64_t k = 0;
    double e2 = 0.0;
    for (int64_t i = 0; i < m; i++)
        if (idx[i] >= 0) { corr[2 * k] = (int32_t)i; corr[2 * k + 1] = idx[i]; e2 += d2[i]; k++; }
    if (k == 0) { *fitness = 0.0; *rmse = 0.0; }                       /* :87-89 */
    else { *fitness = (double)k / (double)m; *rmse = sqrt(e2 / (double)k); } /* :91-93 */
    return k;
}

int vo_registration_icp(const vo_index *ix, const double *tgt, const double *tgt_nrm, int64_t n,
                        const double *src, const double *src_nrm, int64_t m, double max_dist,
                        const double init[16], int estimator, const double gravity[3],
                        double rel_fitness, double rel_rmse, int max_iter, double out_T[16],
                        double *out_fitness, double *out_rmse, int32_t *out_ncorr, int32_t *out_iters,
                        int32_t *out_corr, double *trace) {
    (void)n;
    memcpy(out_T, init, sizeof(double) * 16);
    *out_fitness = 0.0; *out_rmse = 0.0; *out_ncorr = 0; *out_iters = 0;
    if (max_dist <= 0.0) return -1;                                   /* Registration.cpp:148-151 */
    int needs_normals = (estimator == VO_P2PLANE || estimator == VO_P2PLANE_GRAVITY);
    if (needs_normals && (!src_nrm || !tgt_nrm)) return -1;           /* :152-157 */

    double T[16];
    memcpy(T, init, sizeof(T));
    double *pcd = (double *)malloc(sizeof(double) * 3 * (size_t)(m > 0 ? m : 1));
    double *pn = src_nrm ? (double *)malloc(sizeof(double) * 3 * (size_t)(m > 0 ? m : 1)) : NULL;
    memcpy(pcd, src, sizeof(double) * 3 * (size_t)m);                 /* :162 deep copy */
    if (pn) memcpy(pn, src_nrm, sizeof(double) * 3 * (size_t)m);
    if (!is_identity(init)) transform_cloud(pcd, pn, m, init);        /* :163-165 */
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    double *d2 = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    int32_t *corr = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(m > 0 ? m : 1));
    double fit, rmse;
    int64_t k = corr_pass(ix, pcd, m, max_dist, idx, d2, corr, &fit, &rmse); /* :167 */
    if (trace) { trace[0] = fit; trace[1] = rmse; trace[2] = (double)k; memcpy(trace + 3, T, sizeof(T)); }
    int it = 0;
    for (int i = 0; i < max_iter; i++) {                              /* :169 */
        double upd[16];
        switch (estimator) {                                          /* :172 virtual call */
            case VO_P2P: case VO_P2P_CICP: vo_estimate_p2p(pcd, tgt, corr, k, 0, upd); break;
            case VO_P2PLANE: vo_estimate_p2plane(pcd, tgt, tgt_nrm, corr, k, upd); break;
            case VO_P2PLANE_GRAVITY: vo_estimate_p2plane_gravity(pcd, tgt, tgt_nrm, corr, k, gravity, upd); break;
            default: mat4_identity(upd);
        }
        mat4_mul(upd, T, T);                                          /* :174 */
        transform_cloud(pcd, pn, m, upd);                             /* :175 */
        double pf = fit, pr = rmse;                                   /* :176 backup */
        k = corr_pass(ix, pcd, m, max_dist, idx, d2, corr, &fit, &rmse); /* :177-178 */
        it = i + 1;
        if (trace) {
            double *row = trace + 19 * it;
            row[0] = fit; row[1] = rmse; row[2] = (double)k; memcpy(row + 3, T, sizeof(T));
        }
        if (fabs(pf - fit) < rel_fitness && fabs(pr - rmse) < rel_rmse) break; /* :179-183 */
    }
    memcpy(out_T, T, sizeof(T));
    *out_fitness = fit; *out_rmse = rmse; *out_ncorr = (int32_t)k; *out_iters = it;
    if (out_corr) memcpy(out_corr, corr, sizeof(int32_t) * 2 * (size_t)k);
    free(pcd); free(pn); free(idx); free(d2); free(corr);
    return 0;
}

/* feh::RegisterModelToScene (src/annotation.cpp:29-64) */
int vo_register_model_to_scene(const double *scan, const double *scan_nrm, int64_t n,
                               const double *model, const double *model_nrm, int64_t m, int level,
                               double threshold, int point_to_plane, double out_T[16],
                               int32_t *out_ncorr, int32_t *out_best_level) {
    vo_index *ix = vo_index_create(scan, n, threshold);
    if (!ix) return -1;
    double interval = 2 * M_PI / level;                               /* :35 */
    int32_t best_k = 0, best_i = -1;
    double bestT[16];
    mat4_identity(bestT); /* best_result default-constructed: transformation_ = Identity */
    for (int i = 0; i < level; ++i) {                                 /* :37 */
        double a = interval * i, c = cos(a), s = sin(a);
        /* AngleAxis(a, UnitY).toRotationMatrix() (:41-43) */
        double init[16] = {c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1};
        double T[16], fit, rmse;
        int32_t k, iters;
        vo_registration_icp(ix, scan, scan_nrm, n, model, model_nrm, m, threshold, init,
                            point_to_plane ? VO_P2PLANE : VO_P2P_CICP, NULL, 1e-6, 1e-6, 30,
                            T, &fit, &rmse, &k, &iters, NULL, NULL);  /* :45-57, default criteria */
        if (k > best_k) { best_k = k; best_i = i; memcpy(bestT, T, sizeof(T)); } /* :59-61 strict > */
    }
    memcpy(out_T, bestT, sizeof(bestT));
    if (out_ncorr) *out_ncorr = best_k;
    if (out_best_level) *out_best_level = best_i;
    vo_index_destroy(ix);
    return 0;
}
