// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
// MINCO s=3 non-uniform (minimum-jerk quintic spline through waypoints): the step on both sides of the hot loop
// (SURVEY §8f #1). Restates minco.hpp:397-655 (MINCO_S3NU) and the no-pivot banded LU of minco.hpp:43-198.
// Row layout of the 6N x 6N system follows minco.hpp:433-513 so the unpivoted elimination sees the same pivots.
#pragma once
#include <vector>
#include <cmath>
#include <algorithm>

namespace orc {

struct Band {  // bandwidth 6 below / 6 above, row-major band storage: a(i,j) at d[i*13 + (j-i+6)]
    int n = 0; std::vector<double> d;
    void create(int n_) { n = n_; d.assign((size_t)n * 13, 0.0); }
    double &at(int i, int j) { return d[(size_t)i * 13 + (j - i + 6)]; }
    double at(int i, int j) const { return d[(size_t)i * 13 + (j - i + 6)]; }
    void lu() {  // minco.hpp:99-131
        for (int k = 0; k <= n - 2; k++) {
            const int iM = std::min(k + 6, n - 1), jM = std::min(k + 6, n - 1);
            const double piv = at(k, k);
            for (int i = k + 1; i <= iM; i++) if (at(i, k) != 0.0) at(i, k) /= piv;
            for (int j = k + 1; j <= jM; j++) {
                const double c = at(k, j);
                if (c != 0.0) for (int i = k + 1; i <= iM; i++) if (at(i, k) != 0.0) at(i, j) -= at(i, k) * c;
            }
        }
    }
    // rhs: n x 3 column-major (Eigen MatrixX3d)
    void solve(double *b) const {  // minco.hpp:136-164
        for (int j = 0; j <= n - 1; j++) {
            const int iM = std::min(j + 6, n - 1);
            for (int i = j + 1; i <= iM; i++) if (at(i, j) != 0.0) for (int c = 0; c < 3; c++) b[c * n + i] -= at(i, j) * b[c * n + j];
        }
        for (int j = n - 1; j >= 0; j--) {
            for (int c = 0; c < 3; c++) b[c * n + j] /= at(j, j);
            const int iM = std::max(0, j - 6);
            for (int i = iM; i <= j - 1; i++) if (at(i, j) != 0.0) for (int c = 0; c < 3; c++) b[c * n + i] -= at(i, j) * b[c * n + j];
        }
    }
    void solve_adj(double *b) const {  // minco.hpp:169-197
        for (int j = 0; j <= n - 1; j++) {
            for (int c = 0; c < 3; c++) b[c * n + j] /= at(j, j);
            const int iM = std::min(j + 6, n - 1);
            for (int i = j + 1; i <= iM; i++) if (at(j, i) != 0.0) for (int c = 0; c < 3; c++) b[c * n + i] -= at(j, i) * b[c * n + j];
        }
        for (int j = n - 1; j >= 0; j--) {
            const int iM = std::max(0, j - 6);
            for (int i = iM; i <= j - 1; i++) if (at(j, i) != 0.0) for (int c = 0; c < 3; c++) b[c * n + i] -= at(j, i) * b[c * n + j];
        }
    }
};

struct MincoS3 {
    int N = 0;
    double head[9], tail[9];  // 3x3 column-major: col 0 pos, col 1 vel, col 2 acc
    Band A;
    std::vector<double> b;     // 6N x 3 column-major
    std::vector<double> T1, T2, T3, T4, T5;

    void set_conditions(const double *h, const double *t, int n) {  // minco.hpp:417-431
        N = n; for (int i = 0; i < 9; i++) { head[i] = h[i]; tail[i] = t[i]; }
        A.create(6 * N); b.assign((size_t)18 * N, 0.0);
    }
    double &B(int r, int c) { return b[(size_t)c * 6 * N + r]; }
    double B(int r, int c) const { return b[(size_t)c * 6 * N + r]; }

    // inPs: 3 x (N-1) column-major
    void set_parameters(const double *inPs, const double *ts) {  // minco.hpp:433-513
        T1.assign(ts, ts + N); T2.resize(N); T3.resize(N); T4.resize(N); T5.resize(N);
        for (int i = 0; i < N; i++) { T2[i] = T1[i] * T1[i]; T3[i] = T2[i] * T1[i]; T4[i] = T2[i] * T2[i]; T5[i] = T4[i] * T1[i]; }
        A.create(6 * N); std::fill(b.begin(), b.end(), 0.0);
        A.at(0, 0) = 1.0; A.at(1, 1) = 1.0; A.at(2, 2) = 2.0;
        for (int c = 0; c < 3; c++) { B(0, c) = head[0 * 3 + c]; B(1, c) = head[1 * 3 + c]; B(2, c) = head[2 * 3 + c]; }
        for (int i = 0; i < N - 1; i++) {
            const int r = 6 * i;
            // jerk and snap continuity
            A.at(r + 3, r + 3) = 6.0; A.at(r + 3, r + 4) = 24.0 * T1[i]; A.at(r + 3, r + 5) = 60.0 * T2[i]; A.at(r + 3, r + 9) = -6.0;
            A.at(r + 4, r + 4) = 24.0; A.at(r + 4, r + 5) = 120.0 * T1[i]; A.at(r + 4, r + 10) = -24.0;
            // end position = waypoint
            A.at(r + 5, r) = 1.0; A.at(r + 5, r + 1) = T1[i]; A.at(r + 5, r + 2) = T2[i]; A.at(r + 5, r + 3) = T3[i]; A.at(r + 5, r + 4) = T4[i]; A.at(r + 5, r + 5) = T5[i];
            // position / velocity / acceleration continuity
            A.at(r + 6, r) = 1.0; A.at(r + 6, r + 1) = T1[i]; A.at(r + 6, r + 2) = T2[i]; A.at(r + 6, r + 3) = T3[i]; A.at(r + 6, r + 4) = T4[i]; A.at(r + 6, r + 5) = T5[i]; A.at(r + 6, r + 6) = -1.0;
            A.at(r + 7, r + 1) = 1.0; A.at(r + 7, r + 2) = 2 * T1[i]; A.at(r + 7, r + 3) = 3 * T2[i]; A.at(r + 7, r + 4) = 4 * T3[i]; A.at(r + 7, r + 5) = 5 * T4[i]; A.at(r + 7, r + 7) = -1.0;
            A.at(r + 8, r + 2) = 2.0; A.at(r + 8, r + 3) = 6 * T1[i]; A.at(r + 8, r + 4) = 12 * T2[i]; A.at(r + 8, r + 5) = 20 * T3[i]; A.at(r + 8, r + 8) = -2.0;
            for (int c = 0; c < 3; c++) B(r + 5, c) = inPs[3 * i + c];
        }
        const int e = 6 * N, l = N - 1;
        A.at(e - 3, e - 6) = 1.0; A.at(e - 3, e - 5) = T1[l]; A.at(e - 3, e - 4) = T2[l]; A.at(e - 3, e - 3) = T3[l]; A.at(e - 3, e - 2) = T4[l]; A.at(e - 3, e - 1) = T5[l];
        A.at(e - 2, e - 5) = 1.0; A.at(e - 2, e - 4) = 2 * T1[l]; A.at(e - 2, e - 3) = 3 * T2[l]; A.at(e - 2, e - 2) = 4 * T3[l]; A.at(e - 2, e - 1) = 5 * T4[l];
        A.at(e - 1, e - 4) = 2; A.at(e - 1, e - 3) = 6 * T1[l]; A.at(e - 1, e - 2) = 12 * T2[l]; A.at(e - 1, e - 1) = 20 * T3[l];
        for (int c = 0; c < 3; c++) { B(e - 3, c) = tail[0 * 3 + c]; B(e - 2, c) = tail[1 * 3 + c]; B(e - 1, c) = tail[2 * 3 + c]; }
        A.lu();
        A.solve(b.data());
    }
    double rdot(int r1, int r2) const { return B(r1, 0) * B(r2, 0) + B(r1, 1) * B(r2, 1) + B(r1, 2) * B(r2, 2); }
    double energy() const {  // minco.hpp:530-544
        double e = 0;
        for (int i = 0; i < N; i++) {
            const int r = 6 * i;
            e += 36.0 * rdot(r + 3, r + 3) * T1[i] + 144.0 * rdot(r + 4, r + 3) * T2[i] + 192.0 * rdot(r + 4, r + 4) * T3[i] +
                 240.0 * rdot(r + 5, r + 3) * T3[i] + 720.0 * rdot(r + 5, r + 4) * T4[i] + 720.0 * rdot(r + 5, r + 5) * T5[i];
        }
        return e;
    }
    void energy_grad_coeffs(double *g) const {  // minco.hpp:549-565; g: 6N x 3 col-major
        for (int c = 0; c < 3; c++)
            for (int i = 0; i < N; i++) {
                const int r = 6 * i; double *o = g + (size_t)c * 6 * N + r;
                o[5] = 240.0 * B(r + 3, c) * T3[i] + 720.0 * B(r + 4, c) * T4[i] + 1440.0 * B(r + 5, c) * T5[i];
                o[4] = 144.0 * B(r + 3, c) * T2[i] + 384.0 * B(r + 4, c) * T3[i] + 720.0 * B(r + 5, c) * T4[i];
                o[3] = 72.0 * B(r + 3, c) * T1[i] + 144.0 * B(r + 4, c) * T2[i] + 240.0 * B(r + 5, c) * T3[i];
                o[0] = o[1] = o[2] = 0.0;
            }
    }
    void energy_grad_times(double *g) const {  // minco.hpp:567-581
        for (int i = 0; i < N; i++) {
            const int r = 6 * i;
            g[i] = 36.0 * rdot(r + 3, r + 3) + 288.0 * rdot(r + 4, r + 3) * T1[i] + 576.0 * rdot(r + 4, r + 4) * T2[i] +
                   720.0 * rdot(r + 5, r + 3) * T2[i] + 2880.0 * rdot(r + 5, r + 4) * T3[i] + 3600.0 * rdot(r + 5, r + 5) * T4[i];
        }
    }
    // propogateGrad (minco.hpp:584-654): gradP 3 x (N-1) col-major, gradT N
    void propagate_grad(const double *gradC, const double *gradT, double *gradP, double *gradTout) const {
        std::vector<double> adj(gradC, gradC + (size_t)18 * N);
        A.solve_adj(adj.data());
        auto ADJ = [&](int r, int c) { return adj[(size_t)c * 6 * N + r]; };
        for (int i = 0; i < N - 1; i++) for (int c = 0; c < 3; c++) gradP[3 * i + c] = ADJ(6 * i + 5, c);
        for (int i = 0; i < N; i++) {
            const int r = 6 * i; double s = 0;
            for (int c = 0; c < 3; c++) {
                const double nvel = -(B(r + 1, c) + 2.0 * T1[i] * B(r + 2, c) + 3.0 * T2[i] * B(r + 3, c) + 4.0 * T3[i] * B(r + 4, c) + 5.0 * T4[i] * B(r + 5, c));
                const double nacc = -(2.0 * B(r + 2, c) + 6.0 * T1[i] * B(r + 3, c) + 12.0 * T2[i] * B(r + 4, c) + 20.0 * T3[i] * B(r + 5, c));
                const double njer = -(6.0 * B(r + 3, c) + 24.0 * T1[i] * B(r + 4, c) + 60.0 * T2[i] * B(r + 5, c));
                if (i < N - 1) {
                    const double nsna = -(24.0 * B(r + 4, c) + 120.0 * T1[i] * B(r + 5, c));
                    const double ncra = -120.0 * B(r + 5, c);
                    // Eigen's cwiseProduct(...).sum() over a 6x3 block runs column by column
                    s += nsna * ADJ(r + 3, c) + ncra * ADJ(r + 4, c) + nvel * ADJ(r + 5, c) + nvel * ADJ(r + 6, c) + nacc * ADJ(r + 7, c) + njer * ADJ(r + 8, c);
                } else {
                    s += nvel * ADJ(6 * N - 3, c) + nacc * ADJ(6 * N - 2, c) + njer * ADJ(6 * N - 1, c);
                }
            }
            gradTout[i] = s + gradT[i];
        }
    }
};

}  // namespace orc
