// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's utils/trajectory.hpp, found FIRST on the include path when oracle/ref_minco_wrap.cpp compiles the reference's
// utils/minco.hpp unmodified: minco.hpp needs nothing of it but a Trajectory<D> that getTrajectory() can fill (clear / reserve /
// emplace_back(duration, 3 x (D+1) coefficient matrix, highest power first)). The real header drags in root_finder.hpp (companion-matrix
// eigenvalues, Eigen::Map) — far outside what a checker's Eigen stand-in should imitate, and none of it is on the path being pinned.
#pragma once
#include <Eigen/Eigen>
#include <vector>
template <int D>
class Trajectory {
public:
    std::vector<double> durations;
    std::vector<Eigen::Matrix<double, 3, D + 1>> coeffs;
    void clear() { durations.clear(); coeffs.clear(); }
    void reserve(int n) { durations.reserve((size_t)n); coeffs.reserve((size_t)n); }
    template <class E> void emplace_back(double duration, const Eigen::Base<E> &cmat) { durations.push_back(duration); coeffs.emplace_back(cmat); }
    int getPieceNum() const { return (int)durations.size(); }
};
