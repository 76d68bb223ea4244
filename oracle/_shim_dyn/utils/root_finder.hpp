// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's utils/root_finder.hpp, found FIRST on the include path when the reference's utils/trajectory.hpp is compiled
// unmodified for oracle/_ref/libref_minco.so. trajectory.hpp's Piece / Trajectory are class TEMPLATES: only the members the wrapper calls
// (getPos_Vel_Acc_Jerk, locatePieceIdx, getTotalDuration, emplace_back, ...) are instantiated; the members that use the polynomial root finder
// (getMaxVelRate, checkMaxAccRate, ...) are merely parsed, for which these declarations — never defined, never called — are enough. The real
// header needs companion-matrix eigenvalues and Eigen::Map, far outside what a checker's Eigen stand-in should imitate.
#pragma once
#include <Eigen/Eigen>
#include <set>
namespace RootFinder {
template <class... A> Eigen::VectorXd polySqr(A &&...);
template <class... A> Eigen::VectorXd polyConv(A &&...);
template <class... A> double polyVal(A &&...);
template <class... A> int countRoots(A &&...);
template <class... A> std::set<double> solvePolynomial(A &&...);
}  // namespace RootFinder
