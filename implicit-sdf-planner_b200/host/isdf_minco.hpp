// Host-side MINCO (s = 3, non-uniform time): the step on both sides of the hot loop — it produces the coefficient block the
// kernels consume and pulls (gradC, gradT) back onto (waypoints, durations). Counterpart of minco::MINCO_S3NU
// (utils/include/utils/minco.hpp:397-655) and its banded solver (:43-198); plain C++, no Eigen.
//
// Layout conventions follow the reference so the block can be handed to isdf_eval_* unchanged:
//   coeffs / gradC : 6N x 3 column-major, row 6i+k = t^k coefficient of piece i
//   head / tail    : 3 x 3 column-major, columns = position, velocity, acceleration
//   inner points   : 3 x (N-1) column-major
#pragma once
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstddef>

namespace isdf_host {

// Banded matrix with half-bandwidth 6, stored by diagonals: diag[d][j] holds A(j + d - 6, j)... kept as one flat array.
class Band6 {
public:
    void reset(int n) { n_ = n; a_.assign((size_t)13 * n, 0.0); }
    double &operator()(int i, int j) { return a_[(size_t)(i - j + 6) * n_ + j]; }
    double operator()(int i, int j) const { return a_[(size_t)(i - j + 6) * n_ + j]; }
    int size() const { return n_; }

    // in-place LU without pivoting (the junction-row ordering below keeps the pivots well conditioned)
    void factor() {
        for (int k = 0; k + 1 < n_; k++) {
            const int last = std::min(k + 6, n_ - 1);
            const double piv = (*this)(k, k);
            for (int i = k + 1; i <= last; i++) {
                double &m = (*this)(i, k);
                if (m == 0.0) continue;
                m /= piv;
                for (int j = k + 1; j <= last; j++) {
                    const double u = (*this)(k, j);
                    if (u != 0.0) (*this)(i, j) -= m * u;
                }
            }
        }
    }
    // rhs: n x ncol column-major
    void solve(double *b, int ncol) const {
        for (int c = 0; c < ncol; c++) {
            double *x = b + (size_t)c * n_;
            for (int j = 0; j < n_; j++) {
                const int last = std::min(j + 6, n_ - 1);
                for (int i = j + 1; i <= last; i++) { const double m = (*this)(i, j); if (m != 0.0) x[i] -= m * x[j]; }
            }
            for (int j = n_ - 1; j >= 0; j--) {
                x[j] /= (*this)(j, j);
                for (int i = std::max(0, j - 6); i < j; i++) { const double u = (*this)(i, j); if (u != 0.0) x[i] -= u * x[j]; }
            }
        }
    }
    void solve_transposed(double *b, int ncol) const {
        for (int c = 0; c < ncol; c++) {
            double *x = b + (size_t)c * n_;
            for (int j = 0; j < n_; j++) {
                x[j] /= (*this)(j, j);
                const int last = std::min(j + 6, n_ - 1);
                for (int i = j + 1; i <= last; i++) { const double u = (*this)(j, i); if (u != 0.0) x[i] -= u * x[j]; }
            }
            for (int j = n_ - 1; j >= 0; j--)
                for (int i = std::max(0, j - 6); i < j; i++) { const double m = (*this)(j, i); if (m != 0.0) x[i] -= m * x[j]; }
        }
    }

private:
    int n_ = 0;
    std::vector<double> a_;
};

class MincoS3 {
public:
    void setConditions(const double *head3x3, const double *tail3x3, int pieces) {   // minco.hpp:417-431
        N_ = pieces;
        for (int k = 0; k < 9; k++) { head_[k] = head3x3[k]; tail_[k] = tail3x3[k]; }
        c_.assign((size_t)18 * N_, 0.0);
        T_.assign(N_, 0.0);
    }
    int pieces() const { return N_; }
    const std::vector<double> &coeffs() const { return c_; }     // getCoeffs (minco.hpp:545-548)
    const std::vector<double> &durations() const { return T_; }

    // derivative-order `der` row of the monomial basis at time t, scaled like the reference's rows (no factorial normalisation)
    static void basis(double t, int der, double out[6]) {
        for (int k = 0; k < 6; k++) {
            if (k < der) { out[k] = 0.0; continue; }
            double c = 1.0;
            for (int q = 0; q < der; q++) c *= (k - q);
            double p = 1.0;
            for (int q = 0; q < k - der; q++) p *= t;
            out[k] = c * p;
        }
    }

    void setParameters(const double *innerPts, const double *durations) {           // minco.hpp:433-513
        const int n = 6 * N_;
        for (int i = 0; i < N_; i++) T_[i] = durations[i];
        A_.reset(n);
        std::fill(c_.begin(), c_.end(), 0.0);
        auto put = [&](int row, int col0, const double r[6], double sgn) { for (int k = 0; k < 6; k++) if (r[k] != 0.0) A_(row, col0 + k) = sgn * r[k]; };
        double r[6];
        // start: position, velocity, acceleration
        for (int d = 0; d < 3; d++) { basis(0.0, d, r); put(d, 0, r, 1.0); for (int ax = 0; ax < 3; ax++) C(d, ax) = head_[d * 3 + ax]; }
        for (int i = 0; i + 1 < N_; i++) {
            const int c0 = 6 * i, c1 = 6 * (i + 1), r0 = 6 * i + 3;
            // jerk and snap continuity, waypoint, then position / velocity / acceleration continuity
            basis(T_[i], 3, r); put(r0, c0, r, 1.0); basis(0.0, 3, r); put(r0, c1, r, -1.0);
            basis(T_[i], 4, r); put(r0 + 1, c0, r, 1.0); basis(0.0, 4, r); put(r0 + 1, c1, r, -1.0);
            basis(T_[i], 0, r); put(r0 + 2, c0, r, 1.0);
            for (int ax = 0; ax < 3; ax++) C(r0 + 2, ax) = innerPts[3 * i + ax];
            for (int d = 0; d < 3; d++) { basis(T_[i], d, r); put(r0 + 3 + d, c0, r, 1.0); basis(0.0, d, r); put(r0 + 3 + d, c1, r, -1.0); }
        }
        for (int d = 0; d < 3; d++) {
            basis(T_[N_ - 1], d, r); put(n - 3 + d, n - 6, r, 1.0);
            for (int ax = 0; ax < 3; ax++) C(n - 3 + d, ax) = tail_[d * 3 + ax];
        }
        A_.factor();
        A_.solve(c_.data(), 3);
    }

    // integral of squared jerk and its partial derivatives (minco.hpp:530-582)
    double getEnergy() const {
        double e = 0.0;
        for (int i = 0; i < N_; i++) {
            const double t1 = T_[i], t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
            for (int ax = 0; ax < 3; ax++) {
                const double c3 = C(6 * i + 3, ax), c4 = C(6 * i + 4, ax), c5 = C(6 * i + 5, ax);
                e += 36.0 * c3 * c3 * t1 + 144.0 * c4 * c3 * t2 + 192.0 * c4 * c4 * t3 + 240.0 * c5 * c3 * t3 + 720.0 * c5 * c4 * t4 + 720.0 * c5 * c5 * t5;
            }
        }
        return e;
    }
    void getEnergyPartialGradByCoeffs(double *g) const {
        for (int ax = 0; ax < 3; ax++)
            for (int i = 0; i < N_; i++) {
                const double t1 = T_[i], t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
                const double c3 = C(6 * i + 3, ax), c4 = C(6 * i + 4, ax), c5 = C(6 * i + 5, ax);
                double *o = g + (size_t)ax * 6 * N_ + 6 * i;
                o[0] = o[1] = o[2] = 0.0;
                o[3] = 72.0 * c3 * t1 + 144.0 * c4 * t2 + 240.0 * c5 * t3;
                o[4] = 144.0 * c3 * t2 + 384.0 * c4 * t3 + 720.0 * c5 * t4;
                o[5] = 240.0 * c3 * t3 + 720.0 * c4 * t4 + 1440.0 * c5 * t5;
            }
    }
    void getEnergyPartialGradByTimes(double *g) const {
        for (int i = 0; i < N_; i++) {
            const double t1 = T_[i], t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2;
            double s = 0.0;
            for (int ax = 0; ax < 3; ax++) {
                const double c3 = C(6 * i + 3, ax), c4 = C(6 * i + 4, ax), c5 = C(6 * i + 5, ax);
                s += 36.0 * c3 * c3 + 288.0 * c4 * c3 * t1 + 576.0 * c4 * c4 * t2 + 720.0 * c5 * c3 * t2 + 2880.0 * c5 * c4 * t3 + 3600.0 * c5 * c5 * t4;
            }
            g[i] = s;
        }
    }

    // propogateGrad (minco.hpp:584-654): (dJ/dc, dJ/dT partial) -> (dJ/d inner points 3 x (N-1), dJ/dT total)
    void propagateGrad(const double *gradC, const double *gradT, double *gradPts, double *gradTimes) const {
        const int n = 6 * N_;
        std::vector<double> lam(gradC, gradC + (size_t)18 * N_);
        A_.solve_transposed(lam.data(), 3);
        auto L = [&](int row, int ax) { return lam[(size_t)ax * n + row]; };
        for (int i = 0; i + 1 < N_; i++) for (int ax = 0; ax < 3; ax++) gradPts[3 * i + ax] = L(6 * i + 5, ax);
        // dA/dT_i acts on the rows evaluated at t = T_i: each such row's derivative is the next-order basis row
        double r[6];
        for (int i = 0; i < N_; i++) {
            double s = 0.0;
            for (int ax = 0; ax < 3; ax++) {
                double dval[6];   // derivatives 1..5 of piece i at its end time
                for (int d = 1; d <= 5; d++) {
                    basis(T_[i], d, r);
                    double v = 0.0;
                    for (int k = 0; k < 6; k++) v += r[k] * C(6 * i + k, ax);
                    dval[d] = v;
                }
                if (i + 1 < N_) {
                    const int r0 = 6 * i + 3;
                    s -= dval[4] * L(r0, ax) + dval[5] * L(r0 + 1, ax) + dval[1] * L(r0 + 2, ax) + dval[1] * L(r0 + 3, ax) +
                         dval[2] * L(r0 + 4, ax) + dval[3] * L(r0 + 5, ax);
                } else {
                    s -= dval[1] * L(n - 3, ax) + dval[2] * L(n - 2, ax) + dval[3] * L(n - 1, ax);
                }
            }
            gradTimes[i] = s + gradT[i];
        }
    }

private:
    double &C(int row, int ax) { return c_[(size_t)ax * 6 * N_ + row]; }
    double C(int row, int ax) const { return c_[(size_t)ax * 6 * N_ + row]; }
    int N_ = 0;
    double head_[9], tail_[9];
    Band6 A_;
    std::vector<double> c_, T_;
};

}  // namespace isdf_host
