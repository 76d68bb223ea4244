// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU (FP64, serial) restatement of the reference's collision cost/gradient hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// build, link, import or execute anything in this directory. The product (libisdf_b200.so) never does.
//
// Parity status: the reference ships no golden vectors, known-answer tests or fixtures for this path and its build
// cannot run here (needs Eigen/ROS/PCL/gfortran). Pinned against REFERENCE-COMPILED code (oracle/Makefile `ref`, outputs in oracle/_ref/):
//   - igl/FastWindingNumberForSoups.h, unmodified                              -> winding numbers, oracle mode WN_REF (tests/test_oracle_mesh.py)
//   - utils/flatness.hpp, unmodified, against the element-access Eigen stand-in -> flatness forward / adjoint, bit-identical / 7e-16
//   - utils/minco.hpp, unmodified, against the eager Eigen stand-in              -> orc::Minco: coefficients, dE/dC, adjoint solve bit-identical,
//                                                                                  sums (energy, gradByTimes) to the ulp (tests/test_reference_pins.py)
//   - utils/trajectory.hpp, unmodified, same stand-in (root_finder.hpp: declarations only) -> orc::Traj: piece search, local time, pos / vel / acc / jerk
//                                                                                  bit-identical, junctions and out-of-range times included
//   - map_manager/src/Gridmap3D.cpp, unmodified, same stand-in + ROS stand-ins    -> orc::Grid: getGridIndex (clamping quirk), cube centres, in-map test and
//                                                                                  the AABB gather built on them bit-identical
//   - utils/lbfgs.hpp, unmodified, against the eager Eigen stand-in              -> the product's host L-BFGS driver: every evaluated point, value,
//                                                                                  return code and evaluation count identical
// (the eager stand-in adds reductions left to right where Eigen adds packet-wise: what those two pins fix is the reference's logic and operation
// order, see oracle/_shim_dyn/Eigen/Eigen). Otherwise by (i) finite-difference checks of cost vs gradient (tests/test_oracle_fd.py) and
// (ii) brute-force point-triangle distance as known answer for the BVH query. The penalty loops themselves (hpp:432-649, 766-866), the swept-volume
// search (swm) and the Shape.hpp shape classes need the real Eigen: "parity unpinned" by reference-owned vectors for those.
//
// Reference shorthand used in citations (all under /root/reference/src):
//   hpp:   planner_algorithm/include/planner_algorithm/back_end_optimizer.hpp
//   swm:   swept_volume/include/swept_volume/sw_manager.hpp
//   Shape.hpp / Shape.cpp: utils/include/utils/Shape.hpp, utils/src/Shape.cpp
//   flat:  utils/include/utils/flatness.hpp      traj: utils/include/utils/trajectory.hpp
//   pcs:   map_manager/include/map_manager/PCSmap_manager.h     grid: map_manager/src/Gridmap3D.cpp
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(double s, const V3 &a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(const V3 &a, double s) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
inline V3 &operator+=(V3 &a, const V3 &b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3 &a, const V3 &b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
// Eigen's .normalized(): divides by the norm when it is > 0
inline V3 normalized(const V3 &a) {
    double n2 = dot(a, a);
    if (n2 > 0.0) { double n = std::sqrt(n2); return V3(a.x / n, a.y / n, a.z / n); }
    return a;
}

// row-major 3x3
struct M3 {
    double m[9];
    double operator()(int r, int c) const { return m[3 * r + c]; }
    double &operator()(int r, int c) { return m[3 * r + c]; }
    static M3 identity() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0; return r; }
};
inline V3 mul(const M3 &A, const V3 &v) {
    return V3(A(0, 0) * v.x + A(0, 1) * v.y + A(0, 2) * v.z,
              A(1, 0) * v.x + A(1, 1) * v.y + A(1, 2) * v.z,
              A(2, 0) * v.x + A(2, 1) * v.y + A(2, 2) * v.z);
}
inline V3 mulT(const M3 &A, const V3 &v) {  // A^T v  (== row-vector v^T A)
    return V3(A(0, 0) * v.x + A(1, 0) * v.y + A(2, 0) * v.z,
              A(0, 1) * v.x + A(1, 1) * v.y + A(2, 1) * v.z,
              A(0, 2) * v.x + A(1, 2) * v.y + A(2, 2) * v.z);
}

// Eigen::Quaterniond(w,x,y,z).toRotationMatrix() — no normalisation (hpp:492, hpp:614, swm:340).
inline M3 quat_to_R(const double q[4]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 R;
    R(0, 0) = 1.0 - (tyy + tzz); R(0, 1) = txy - twz;         R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;         R(1, 1) = 1.0 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;         R(2, 1) = tyz + twx;         R(2, 2) = 1.0 - (txx + tyy);
    return R;
}

// smoothedL1 (hpp:332-356). Outputs untouched when x < 0 (quirk Q1).
inline bool smoothedL1(double x, double mu, double &f, double &df) {
    if (x < 0.0) return false;
    if (x > mu) { f = x - 0.5 * mu; df = 1.0; return true; }
    const double xdmu = x / mu;
    const double sqrxdmu = xdmu * xdmu;
    const double mumxd2 = mu - 0.5 * x;
    f = mumxd2 * sqrxdmu * xdmu;
    df = sqrxdmu * ((-0.5) * xdmu + 3.0 * mumxd2 / mu);
    return true;
}

// g_rel^T * (dR^T/dq_k) * d for k = w,x,y,z  (hpp:127-173 used at hpp:816-819 / hpp:859-862)
inline void quat_jac_contract(const double q[4], const V3 &g, const V3 &d, double out[4]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    // D_w
    {
        V3 r0(0, 2 * z, -2 * y), r1(-2 * z, 0, 2 * x), r2(2 * y, -2 * x, 0);
        out[0] = g.x * dot(r0, d) + g.y * dot(r1, d) + g.z * dot(r2, d);
    }
    {
        V3 r0(0, 2 * y, 2 * z), r1(2 * y, -4 * x, 2 * w), r2(2 * z, -2 * w, -4 * x);
        out[1] = g.x * dot(r0, d) + g.y * dot(r1, d) + g.z * dot(r2, d);
    }
    {
        V3 r0(-4 * y, 2 * x, -2 * w), r1(2 * x, 0, 2 * z), r2(2 * w, 2 * z, -4 * y);
        out[2] = g.x * dot(r0, d) + g.y * dot(r1, d) + g.z * dot(r2, d);
    }
    {
        V3 r0(-4 * z, 2 * w, 2 * x), r1(-2 * w, -4 * z, 2 * y), r2(2 * x, 2 * y, 0);
        out[3] = g.x * dot(r0, d) + g.y * dot(r1, d) + g.z * dot(r2, d);
    }
}

// ---------------------------------------------------------------------------------------------
// Differential-flatness map (flat:36-148) and its adjoint (flat:230-406).
// The forward is restated; the adjoint is re-derived by hand in vector form (same mathematical
// adjoint — SURVEY §9 verified the reference's is the true one); tests pin it against central differences.
struct Flat {
    double mass, grav, dh, dv, cp, veps, dh_over_m;
    void reset(double m, double g, double dh_, double dv_, double cp_, double veps_) {  // flat:36-51
        mass = m; grav = g; dh = dh_; dv = dv_; cp = cp_; veps = veps_; dh_over_m = dh / mass;
    }
    // intermediates shared by forward / adjoint
    struct Mid {
        double s, kap, lam, n2, n, td, od, ot;
        V3 zu, z, xi, dz;
    };
    void mid(const V3 &v, const V3 &a, const V3 &j, Mid &m) const {
        m.s = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z + veps);
        m.kap = 1.0 + cp * m.s;
        V3 w(m.kap * v.x, m.kap * v.y, m.kap * v.z);
        m.zu = V3(a.x + dh_over_m * w.x, a.y + dh_over_m * w.y, a.z + dh_over_m * w.z + grav);
        const double s0 = m.zu.x * m.zu.x, s1 = m.zu.y * m.zu.y, s2 = m.zu.z * m.zu.z;
        m.n2 = s0 + s1 + s2;
        m.n = std::sqrt(m.n2);
        m.z = V3(m.zu.x / m.n, m.zu.y / m.n, m.zu.z / m.n);
        m.td = std::sqrt(2.0 * (1.0 + m.z.z));
        // body-rate part (flat:121-146)
        const double ng_den = m.n2 * m.n;
        const double ng00 = (s1 + s2) / ng_den, ng11 = (s0 + s2) / ng_den, ng22 = (s0 + s1) / ng_den;
        const double ng01 = -(m.zu.x * m.zu.y) / ng_den, ng02 = -(m.zu.x * m.zu.z) / ng_den,
                     ng12 = -(m.zu.y * m.zu.z) / ng_den;
        const double vda = v.x * a.x + v.y * a.y + v.z * a.z;
        m.lam = cp * vda / m.s;
        V3 dw(m.kap * a.x + m.lam * v.x, m.kap * a.y + m.lam * v.y, m.kap * a.z + m.lam * v.z);
        m.xi = V3(j.x + dh_over_m * dw.x, j.y + dh_over_m * dw.y, j.z + dh_over_m * dw.z);
        m.dz = V3(ng00 * m.xi.x + ng01 * m.xi.y + ng02 * m.xi.z,
                  ng01 * m.xi.x + ng11 * m.xi.y + ng12 * m.xi.z,
                  ng02 * m.xi.x + ng12 * m.xi.y + ng22 * m.xi.z);
        m.od = m.z.z + 1.0;
        m.ot = m.dz.z / m.od;
    }
    // optimizated_forward, quaternion only (flat:53-86)
    void forward(const V3 &v, const V3 &a, const V3 &j, double q[4]) const {
        Mid m; mid(v, a, j, m);
        q[0] = 0.5 * m.td; q[1] = -m.z.y / m.td; q[2] = m.z.x / m.td; q[3] = 0.0;
    }
    // optimizated_forward, quaternion + body rate (flat:88-148)
    void forward(const V3 &v, const V3 &a, const V3 &j, double q[4], V3 &omg) const {
        Mid m; mid(v, a, j, m);
        q[0] = 0.5 * m.td; q[1] = -m.z.y / m.td; q[2] = m.z.x / m.td; q[3] = 0.0;
        omg.x = -m.dz.y + m.z.y * m.ot;
        omg.y = m.dz.x - m.z.x * m.ot;
        omg.z = (m.z.y * m.dz.x - m.z.x * m.dz.y) / m.od;
    }
    // adjoint of (quat, omg) wrt (pos, vel, acc, jer); same contract as backwardthreadsafe (flat:230-406)
    void backward(const V3 &v, const V3 &a, const V3 &j, const V3 &pos_grad, const V3 &vel_grad,
                  const double qb[4], const V3 &ob, V3 &gp, V3 &gv, V3 &ga, V3 &gj) const {
        Mid m; mid(v, a, j, m);
        const V3 &z = m.z, &dz = m.dz, &zu = m.zu, &xi = m.xi;
        // omega -> (z, dz)
        V3 dzb(ob.y + ob.z * z.y / m.od, -ob.x - ob.z * z.x / m.od, 0.0);
        const double otb = ob.x * z.y - ob.y * z.x;
        V3 zb(-ob.y * m.ot - ob.z * dz.y / m.od, ob.x * m.ot + ob.z * dz.x / m.od, 0.0);
        double odb = -ob.z * (z.y * dz.x - z.x * dz.y) / (m.od * m.od) - otb * dz.z / (m.od * m.od);
        dzb.z = otb / m.od;
        // quaternion -> z
        const double tdb = 0.5 * qb[0] + qb[1] * z.y / (m.td * m.td) - qb[2] * z.x / (m.td * m.td);
        zb.y += -qb[1] / m.td;
        zb.x += qb[2] / m.td;
        zb.z = odb + tdb / m.td;
        // dz = P xi, P = I/n - zu zu^T / n^3
        const double n3 = m.n2 * m.n, n5 = n3 * m.n2;
        const double zu_xi = dot(zu, xi), zu_dzb = dot(zu, dzb), dzb_xi = dot(dzb, xi);
        V3 xib(dzb.x / m.n - zu.x * zu_dzb / n3, dzb.y / m.n - zu.y * zu_dzb / n3, dzb.z / m.n - zu.z * zu_dzb / n3);
        V3 zub;
        for (int k = 0; k < 3; k++)
            zub[k] = -dzb_xi * zu[k] / n3 - (zu_xi * dzb[k] + zu_dzb * xi[k]) / n3 + 3.0 * zu_dzb * zu_xi * zu[k] / n5;
        // z = zu / n
        const double zb_z = dot(zb, z);
        for (int k = 0; k < 3; k++) zub[k] += (zb[k] - zb_z * z[k]) / m.n;
        // zu = a + h w + g e3 ; xi = j + h dw
        V3 wb = dh_over_m * zub;
        V3 dwb = dh_over_m * xib;
        gj = xib;
        ga = zub + m.kap * dwb;
        double kapb = dot(a, dwb) + dot(v, wb);
        const double lamb = dot(v, dwb);
        V3 vt = m.lam * dwb + m.kap * wb;
        // lam = cp (v.a)/s
        vt += (lamb * cp / m.s) * a;
        ga += (lamb * cp / m.s) * v;
        double sb = -lamb * m.lam / m.s + cp * kapb;
        vt += (sb / m.s) * v;
        gv = vt + vel_grad;
        gp = pos_grad;
    }
};

// beta basis at local time s (hpp:472-483): value and first four derivatives of [1,s,..,s^5]
inline void beta_basis(double s1, double b0[6], double b1[6], double b2[6], double b3[6], double b4[6]) {
    const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
    b0[0] = 1.0; b0[1] = s1; b0[2] = s2; b0[3] = s3; b0[4] = s4; b0[5] = s5;
    b1[0] = 0.0; b1[1] = 1.0; b1[2] = 2.0 * s1; b1[3] = 3.0 * s2; b1[4] = 4.0 * s3; b1[5] = 5.0 * s4;
    b2[0] = 0.0; b2[1] = 0.0; b2[2] = 2.0; b2[3] = 6.0 * s1; b2[4] = 12.0 * s2; b2[5] = 20.0 * s3;
    b3[0] = 0.0; b3[1] = 0.0; b3[2] = 0.0; b3[3] = 6.0; b3[4] = 24.0 * s1; b3[5] = 60.0 * s2;
    b4[0] = 0.0; b4[1] = 0.0; b4[2] = 0.0; b4[3] = 0.0; b4[4] = 24.0; b4[5] = 120.0 * s1;
}

}  // namespace orc
