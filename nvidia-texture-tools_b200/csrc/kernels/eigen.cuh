// Principal axis of a small point set, operation-for-operation as the reference computes it for the BC6H / BC7 "rough"
// endpoint fit:
//   Fit::computePrincipalComponent_EigenSolver (3-D)    src/nvmath/Fitting.cpp:275-281   (centroid :61-73, covariance :115-142)
//   Fit::computePrincipalComponent_EigenSolver (4-D)    src/nvmath/Fitting.cpp:310-316   (centroid :91-103, covariance :173-204)
//   eigenSolveSymmetric3 / EigenSolver3_Tridiagonal / EigenSolver3_QLAlgorithm           :421-598
//   eigenSolveSymmetric4 / EigenSolver4_Tridiagonal / EigenSolver4_QLAlgorithm           :607-820
// Every +,-,*,/ and sqrtf is a single IEEE operation (library built with -fmad=false), so the result is bit-identical
// to the pinned reference build.
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

// QL iteration with implicit shifts on a symmetric tridiagonal matrix (diag, subd); mat accumulates the rotations.
// Returns false when an eigenvalue needs more than 32 iterations (the reference then returns a zero vector).
template <int N> NVB_DEV bool eigen_ql(float mat[N][N], float diag[N], float subd[N]) {
    const int maxiter = 32;
    for (int ell = 0; ell < N; ell++) {
        int iter;
        for (iter = 0; iter < maxiter; iter++) {
            int m;
            for (m = ell; m < N - 1; m++) {
                const float dd = fabsf(diag[m]) + fabsf(diag[m + 1]);
                if (fabsf(subd[m]) + dd == dd) break;
            }
            if (m == ell) break;
            float g = (diag[ell + 1] - diag[ell]) / (2 * subd[ell]);
            float r = sqrtf(g * g + 1);
            if (g < 0) g = diag[m] - diag[ell] + subd[ell] / (g - r);
            else g = diag[m] - diag[ell] + subd[ell] / (g + r);
            float s = 1, c = 1, p = 0;
            for (int i = m - 1; i >= ell; i--) {
                float f = s * subd[i], b = c * subd[i];
                if (fabsf(f) >= fabsf(g)) {
                    c = g / f;
                    r = sqrtf(c * c + 1);
                    subd[i + 1] = f * r;
                    s = 1 / r;
                    c *= s;
                } else {
                    s = f / g;
                    r = sqrtf(s * s + 1);
                    subd[i + 1] = g * r;
                    c = 1 / r;
                    s *= c;
                }
                g = diag[i + 1] - p;
                r = (diag[i] - g) * s + 2 * b * c;
                p = s * r;
                diag[i + 1] = g + p;
                g = c * r - b;
                for (int k = 0; k < N; k++) {
                    f = mat[k][i + 1];
                    mat[k][i + 1] = s * mat[k][i] + c * f;
                    mat[k][i] = c * mat[k][i] - s * f;
                }
            }
            diag[ell] -= p;
            subd[ell] = g;
            subd[m] = 0;
        }
        if (iter == maxiter) return false;
    }
    return true;
}

// First principal axis of n 3-D points (n <= 16).  pts[i][0..2].
NVB_DEV void principal_axis3(int n, const float (*pts)[3], float dir[3]) {
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    for (int i = 0; i < n; i++) {
        cx += pts[i][0];
        cy += pts[i][1];
        cz += pts[i][2];
    }
    const float is = 1.0f / (float)n;  // Vector3::operator/=(float) multiplies by the reciprocal (Vector.inl:135-141)
    cx *= is;
    cy *= is;
    cz *= is;
    float cov[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const float vx = pts[i][0] - cx, vy = pts[i][1] - cy, vz = pts[i][2] - cz;
        cov[0] += vx * vx;
        cov[1] += vx * vy;
        cov[2] += vx * vz;
        cov[3] += vy * vy;
        cov[4] += vy * vz;
        cov[5] += vz * vz;
    }
    dir[0] = dir[1] = dir[2] = 0.0f;
    if (cov[0] == 0 && cov[3] == 0 && cov[5] == 0) return;
    // Householder reduction of the 3x3 (closed form)
    float mat[3][3], diag[3], subd[3];
    {
        const float a = cov[0];
        float b = cov[1], c = cov[2];
        const float d = cov[3], e = cov[4], f = cov[5];
        diag[0] = a;
        subd[2] = 0.f;
        if (fabsf(c) >= 1e-08f) {
            const float ell = sqrtf(b * b + c * c);
            b /= ell;
            c /= ell;
            const float q = 2 * b * e + c * (f - d);
            diag[1] = d + c * q;
            diag[2] = f - c * q;
            subd[0] = ell;
            subd[1] = e - b * q;
            mat[0][0] = 1; mat[0][1] = 0; mat[0][2] = 0;
            mat[1][0] = 0; mat[1][1] = b; mat[1][2] = c;
            mat[2][0] = 0; mat[2][1] = c; mat[2][2] = -b;
        } else {
            diag[1] = d;
            diag[2] = f;
            subd[0] = b;
            subd[1] = e;
            mat[0][0] = 1; mat[0][1] = 0; mat[0][2] = 0;
            mat[1][0] = 0; mat[1][1] = 1; mat[1][2] = 0;
            mat[2][0] = 0; mat[2][1] = 0; mat[2][2] = 1;
        }
    }
    if (!eigen_ql<3>(mat, diag, subd)) return;
    // the eigenvector of the largest eigenvalue after the reference's three conditional swaps (:456-470)
    int order[3] = {0, 1, 2};
    float ev[3] = {diag[0], diag[1], diag[2]};
    if (ev[2] > ev[0] && ev[2] > ev[1]) {
        float t = ev[0]; ev[0] = ev[2]; ev[2] = t;
        int o = order[0]; order[0] = order[2]; order[2] = o;
    }
    if (ev[1] > ev[0]) {
        float t = ev[0]; ev[0] = ev[1]; ev[1] = t;
        int o = order[0]; order[0] = order[1]; order[1] = o;
    }
    const int j = order[0];
    dir[0] = mat[0][j];
    dir[1] = mat[1][j];
    dir[2] = mat[2][j];
}

// First principal axis of n 4-D points.  The reference's 4x4 Householder step never runs: its epsilon is
// 1e-6 * max(FLT_MAX, ...) (Fitting.cpp:686-690), so every column is "already tridiagonal" and the QL iteration sees
// the diagonal and first sub-diagonal of the covariance matrix with Q = identity.  Reproduced as is.
NVB_DEV void principal_axis4(int n, const float (*pts)[4], float dir[4]) {
    float c[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) c[k] += pts[i][k];
    for (int k = 0; k < 4; k++) c[k] /= (float)n;  // Vector4::operator/=(float) is a true division (Vector.inl:242-248)
    float cov[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const float vx = pts[i][0] - c[0], vy = pts[i][1] - c[1], vz = pts[i][2] - c[2], vw = pts[i][3] - c[3];
        cov[0] += vx * vx;
        cov[1] += vx * vy;
        cov[2] += vx * vz;
        cov[3] += vx * vw;
        cov[4] += vy * vy;
        cov[5] += vy * vz;
        cov[6] += vy * vw;
        cov[7] += vz * vz;
        cov[8] += vz * vw;
        cov[9] += vw * vw;
    }
    dir[0] = dir[1] = dir[2] = dir[3] = 0.0f;
    if (cov[0] == 0 && cov[4] == 0 && cov[7] == 0 && cov[9] == 0) return;
    float mat[4][4], diag[4], subd[4];
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++) mat[i][k] = (i == k) ? 1.0f : 0.0f;
    diag[0] = cov[0];
    diag[1] = cov[4];
    diag[2] = cov[7];
    diag[3] = cov[9];
    subd[0] = cov[1];
    subd[1] = cov[5];
    subd[2] = cov[8];
    subd[3] = 0.0f;
    if (!eigen_ql<4>(mat, diag, subd)) return;
    // selection sort by eigenvalue, descending, swapping when strictly greater (:645-655)
    int order[4] = {0, 1, 2, 3};
    float ev[4] = {diag[0], diag[1], diag[2], diag[3]};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (ev[j] > ev[i]) {
                float t = ev[i]; ev[i] = ev[j]; ev[j] = t;
                int o = order[i]; order[i] = order[j]; order[j] = o;
            }
    const int j = order[0];
    for (int k = 0; k < 4; k++) dir[k] = mat[k][j];
}

}  // namespace nvb
