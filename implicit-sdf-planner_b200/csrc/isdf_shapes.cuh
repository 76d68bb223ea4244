// Device-side robot-shape SDFs: the BasicShape virtual surface (Shape.hpp:469-472) flattened into a warp-uniform
// switch (no virtual calls on the device), the finite-difference gradient contract of DEFINE_USEFUL_FUNCTION
// (Shape.hpp:32-88) and the mesh shape (Shape.cpp:105-151; isdf_mesh.cuh).
//
// params[] layout per kind (isdf_set_shape_named fills the reference's hard-coded members):
//   BALL {radius}                       POINT {}                         TORUS {ring radius, tube radius}
//   CAPPED_TORUS {sc.x, sc.y, ra, rb}   CAPPED_CONE {ra, rb}             ROUNDED_CONE {r1, r2, h}
//   WIREFRAME_BOX {sx, sy, sz, thick}   BEND_LINEAR {capsule half-length, radius}
//   TWIST_BOX / BEND_BOX {sx, sy, sz, k}   TABLE {a1(3), b1(3), a2(3), b2(3)}   TREFOIL {}
//   SMOOTH_DIFFERENCE / SMOOTH_INTERSECTION {sx, sy, sz, sphere radius}   CSG {}   BOX {hx, hy, hz} (half extents)
#pragma once
#include "isdf_math.cuh"
#include "isdf_mesh.cuh"
#include "../../include/isdf.h"

namespace isdf {

struct DevShape {
    int kind;
    double par[12];
    double rot[9];     // Generalshape::Rotate, row-major; applied as (p - trans) * Rotate  (row vector)
    double trans[3];
    double csg_ry[9];  // rotate(c, acos(0), Y x X) and rotate(c, acos(0), Z x X) of Shape.hpp:2021-2059, built on the host
    double csg_rz[9];
    DevMesh mesh;
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmax(fmin(v, hi), lo); }

// length(max(q,0)) + min(max(q.x,q.y,q.z),0)
__device__ __forceinline__ double sd_box_q(double qx, double qy, double qz) {
    const double ax = fmax(qx, 0.0), ay = fmax(qy, 0.0), az = fmax(qz, 0.0);
    return sqrt(ax * ax + ay * ay + az * az) + fmin(fmax(qx, fmax(qy, qz)), 0.0);
}

__device__ __forceinline__ d3 shape_pre(const DevShape &S, d3 p) {
    const double x = p.x - S.trans[0], y = p.y - S.trans[1], z = p.z - S.trans[2];
    return mk3(x * S.rot[0] + y * S.rot[3] + z * S.rot[6],
               x * S.rot[1] + y * S.rot[4] + z * S.rot[7],
               x * S.rot[2] + y * S.rot[5] + z * S.rot[8]);
}

// getonlySDF for the analytic kinds (mesh handled by the caller). p_in is the body-frame query point.
static __device__ double shape_sdf_analytic(const DevShape &S, d3 p_in) {
    if (S.kind == ISDF_SHAPE_BALL) return len3(p_in) - S.par[0];   // Shape.hpp:616-619
    if (S.kind == ISDF_SHAPE_POINT) return len3(p_in);             // Shape.hpp:647-650
    const d3 p = shape_pre(S, p_in);
    switch (S.kind) {
    case ISDF_SHAPE_TORUS: {  // Shape.hpp:843-848
        const double qx = sqrt(p.x * p.x + p.z * p.z) - S.par[0];
        return sqrt(qx * qx + p.y * p.y) - S.par[1];
    }
    case ISDF_SHAPE_CAPPED_TORUS: {  // Shape.hpp:913-919
        const double px = fabs(p.x);
        const double k = (S.par[1] * px > S.par[0] * p.y) ? (px * S.par[0] + p.y * S.par[1]) : sqrt(px * px + p.y * p.y);
        return sqrt((px * px + p.y * p.y + p.z * p.z) + S.par[2] * S.par[2] - 2.0 * S.par[2] * k) - S.par[3];
    }
    case ISDF_SHAPE_CAPPED_CONE: {  // Shape.hpp:955-974 with a = (0,0,-1), b = (0,0,1)
        const double ra = S.par[0], rb = S.par[1];
        const double rba = rb - ra;
        const double baba = 4.0;                                    // |b - a|^2
        const double pz1 = p.z + 1.0;                               // (p - a).z
        const double papa = p.x * p.x + p.y * p.y + pz1 * pz1;
        const double paba = (p.x * 0.0 + p.y * 0.0 + pz1 * 2.0) / baba;
        const double x = sqrt(papa - paba * paba * baba);
        const double cax = fmax(0.0, x - (paba < 0.5 ? ra : rb));
        const double cay = fabs(paba - 0.5) - 0.5;
        const double k = rba * rba + baba;
        const double f = clampd((rba * (x - ra) + paba * baba) / k, 0.0, 1.0);
        const double cbx = x - ra - f * rba;
        const double cby = paba - f;
        const double s = (cbx < 0 && cay < 0) ? -1.0 : 1.0;
        const double d = sqrt(fmin(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba));
        return s * sqrt(fabs(d)) / fabs(baba);
    }
    case ISDF_SHAPE_ROUNDED_CONE: {  // Shape.hpp:1018-1029
        const double r1 = S.par[0], r2 = S.par[1], h = S.par[2];
        const double qx = sqrt(p.x * p.x + p.y * p.y), qy = p.z;
        const double b = (r1 - r2) / h;
        const double a = sqrt(1.0 - b * b);
        const double k = -b * qx + a * qy;
        if (k < 0) return sqrt(qx * qx + qy * qy) - r1;
        if (k > a * h) return sqrt(qx * qx + (qy - h) * (qy - h)) - r2;
        return (a * qx + b * qy) - r1;
    }
    case ISDF_SHAPE_WIREFRAME_BOX: {  // Shape.hpp:1066-1085
        const double th = S.par[3];
        const double px = fabs(p.x) - S.par[0] / 2 - th / 2, py = fabs(p.y) - S.par[1] / 2 - th / 2, pz = fabs(p.z) - S.par[2] / 2 - th / 2;
        const double qx = fabs(px + th / 2) - th / 2, qy = fabs(py + th / 2) - th / 2, qz = fabs(pz + th / 2) - th / 2;
        return fmin(fmin(sd_box_q(px, qy, qz), sd_box_q(qx, py, qz)), sd_box_q(qx, qy, pz));
    }
    case ISDF_SHAPE_BEND_LINEAR: {  // Shape.hpp:1141-1151 / :1206-1216
        const double hz = S.par[0], rad = S.par[1];
        // t = clip((p - p0).ab / ab.ab), p0 = (0,0,-1), ab = (0,0,2)
        const double t = clampd((p.x * 0.0 + p.y * 0.0 + (p.z + 1.0) * 2.0) / 4.0, 0.0, 1.0);
        const double u = 2 * t - 1;
        const double e = (t < 0.5) ? (2 * t * t) : (-0.5 * (u * (u - 2) - 1));
        const d3 pp = mk3(e * -1.0 + p.x, e * 0.0 + p.y, e * 0.0 + p.z);
        // capsule from (0,0,-hz) to (0,0,hz)
        const d3 pa = mk3(pp.x, pp.y, pp.z + hz);
        const double baz = hz + hz;
        const double hh = fmax(fmin((pa.x * 0.0 + pa.y * 0.0 + pa.z * baz) / (baz * baz), 1.0), 0.0);
        const d3 r = mk3(pa.x - hh * 0.0, pa.y - hh * 0.0, pa.z - hh * baz);
        return len3(r) - rad;
    }
    case ISDF_SHAPE_TWIST_BOX:
    case ISDF_SHAPE_BEND_BOX: {  // Shape.hpp:1253-1267 / :1307-1321
        const double ang = (S.kind == ISDF_SHAPE_TWIST_BOX) ? S.par[3] * p.z : S.par[3] * p.x;
        const double c = cos(ang), s = sin(ang);
        const double rx = c * p.x - s * p.y, ry = s * p.x + c * p.y;
        return sd_box_q(fabs(rx) - S.par[0] / 2.0, fabs(ry) - S.par[1] / 2.0, fabs(p.z) - S.par[2] / 2.0);
    }
    case ISDF_SHAPE_TABLE: {  // Shape.hpp:1362-1381
        const double x = fabs(p.x), y = fabs(p.y), z = p.z;
        const double *a1 = S.par, *b1 = S.par + 3, *a2 = S.par + 6, *b2 = S.par + 9;
        const double f1 = sd_box_q(fabs(x - (a1[0] + b1[0]) * 0.5) - (b1[0] - a1[0]) * 0.5,
                                   fabs(y - (a1[1] + b1[1]) * 0.5) - (b1[1] - a1[1]) * 0.5,
                                   fabs(z - (a1[2] + b1[2]) * 0.5) - (b1[2] - a1[2]) * 0.5);
        const double f2 = sd_box_q(fabs(x - (a2[0] + b2[0]) * 0.5) - (b2[0] - a2[0]) * 0.5,
                                   fabs(y - (a2[1] + b2[1]) * 0.5) - (b2[1] - a2[1]) * 0.5,
                                   fabs(z - (a2[2] + b2[2]) * 0.5) - (b2[2] - a2[2]) * 0.5);
        return fmin(f1, f2);
    }
    case ISDF_SHAPE_TREFOIL: {  // Shape.hpp:1464-1502
        const double PI = 3.14159265358979323846;
        const double a = atan2(p.y, p.x);
        double qx = sqrt(p.x * p.x + p.y * p.y) - 3.5, qy = -p.z;
        {
            double s, c; sincos(1.5 * a, &s, &c);
            const double nx = qx * c + qy * s, ny = qy * c + (-qx) * s; qx = nx; qy = ny;
        }
        {
            const double a2 = -PI * (floor(atan2(qy, qx) / PI + 0.5));
            double s, c; sincos(a2, &s, &c);
            const double nx = qx * c + qy * s, ny = qy * c + (-qx) * s; qx = nx; qy = ny;
        }
        qx -= 1.0;
        const double dx = fabs(qx) - 0.2, dy = fabs(qy) - 0.2;
        const double mx = fmax(dx, 0.0), my = fmax(dy, 0.0);
        const double d = (fmin(fmax(dx, dy), 0.0) + sqrt(mx * mx + my * my)) - 0.05;
        return 0.4 * fmin(d, 100.0);
    }
    case ISDF_SHAPE_SMOOTH_DIFFERENCE: {  // Shape.hpp:1537-1549
        const double b = sd_box_q(fabs(p.x) - S.par[0] / 2.0, fabs(p.y) - S.par[1] / 2.0, fabs(p.z) - S.par[2] / 2.0);
        const double sp = len3(p) - S.par[3];
        const double h = clampd(0.5 - 0.5 * (b + sp) / 0.25, 0.0, 1.0);
        return (b - (b + sp) * h) + 0.25 * h * (1.0 - h);
    }
    case ISDF_SHAPE_SMOOTH_INTERSECTION: {  // Shape.hpp:1592-1604 / :1647-1659
        const double b = sd_box_q(fabs(p.x) - S.par[0] / 2.0, fabs(p.y) - S.par[1] / 2.0, fabs(p.z) - S.par[2] / 2.0);
        const double sp = len3(p) - S.par[3];
        const double h = clampd(0.5 - 0.5 * (sp - b) / 0.25, 0.0, 1.0);
        return (sp + (b - sp) * h) + 0.25 * h * (1.0 - h);
    }
    case ISDF_SHAPE_CSG: {  // Shape.hpp:2286-2298
        const double f = fmax(len3(p) - 3.0, sd_box_q(fabs(p.x) - 4.5 / 2.0, fabs(p.y) - 4.5 / 2.0, fabs(p.z) - 4.5 / 2.0));
        const double c1 = sqrt(p.x * p.x + p.y * p.y) - 1.5;
        const double *A = S.csg_ry, *B = S.csg_rz;
        const double ax = A[0] * p.x + A[1] * p.y + A[2] * p.z, ay = A[3] * p.x + A[4] * p.y + A[5] * p.z;
        const double bx = B[0] * p.x + B[1] * p.y + B[2] * p.z, by = B[3] * p.x + B[4] * p.y + B[5] * p.z;
        const double c2 = sqrt(ax * ax + ay * ay) - 1.5, c3 = sqrt(bx * bx + by * by) - 1.5;
        return fmax(f, -fmin(fmin(c1, c2), c3));
    }
    case ISDF_SHAPE_BOX:  // Shape.hpp:2343-2351
        return sd_box_q(fabs(p.x) - S.par[0], fabs(p.y) - S.par[1], fabs(p.z) - S.par[2]);
    default: return 0.0;
    }
}

__device__ __forceinline__ double shape_sdf(const DevShape &S, d3 p) {
    if (S.kind == ISDF_SHAPE_MESH) { d3 g; return mesh_sdf_grad(S.mesh, p, 1e300, g); }
    return shape_sdf_analytic(S, p);
}

// gradient after the SDF at p is known: DEFINE_USEFUL_FUNCTION's normalised central difference (Shape.hpp:33-57),
// Ball/Point closed form (Shape.hpp:624-627), Box one-sided un-normalised FD (Shape.hpp:2363-2377).
__device__ __forceinline__ d3 shape_grad_analytic(const DevShape &S, d3 p, double sdf_at_p) {
    if (S.kind == ISDF_SHAPE_BALL || S.kind == ISDF_SHAPE_POINT) return unit3(p);
    if (S.kind == ISDF_SHAPE_BOX) {
        const double dx = 0.01;
        const double gx = shape_sdf_analytic(S, mk3(p.x + dx, p.y, p.z)) - sdf_at_p;
        const double gy = shape_sdf_analytic(S, mk3(p.x, p.y + dx, p.z)) - sdf_at_p;
        const double gz = shape_sdf_analytic(S, mk3(p.x, p.y, p.z + dx)) - sdf_at_p;
        return mk3(gx / dx, gy / dx, gz / dx);
    }
    const double dx = 0.000005;
    // the macro steps "temp -= dx; ...; temp += 2*dx": the upper abscissa is (p - dx) + 2 dx, not p + dx.
    // One rolled loop over the three axes keeps a single inlined copy of the SDF switch per call site.
    double gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        d3 lo = p, hi = p;
        if (a == 0) { lo.x = p.x - dx; hi.x = lo.x + 2 * dx; }
        else if (a == 1) { lo.y = p.y - dx; hi.y = lo.y + 2 * dx; }
        else { lo.z = p.z - dx; hi.z = lo.z + 2 * dx; }
        const double fl = shape_sdf_analytic(S, lo);
        const double diff = shape_sdf_analytic(S, hi) - fl;
        if (a == 0) gx = diff; else if (a == 1) gy = diff; else gz = diff;
    }
    return unit3(mk3(gx / (2 * dx), gy / (2 * dx), gz / (2 * dx)));
}

// getSDFwithGrad1
__device__ __forceinline__ double shape_sdf_grad(const DevShape &S, d3 p, d3 &g) {
    if (S.kind == ISDF_SHAPE_MESH) return mesh_sdf_grad(S.mesh, p, 1e300, g);
    const double s = shape_sdf_analytic(S, p);
    g = shape_grad_analytic(S, p, s);
    return s;
}
// getonlyGrad1
__device__ __forceinline__ d3 shape_grad(const DevShape &S, d3 p) {
    if (S.kind == ISDF_SHAPE_MESH) { d3 g; mesh_sdf_grad(S.mesh, p, 1e300, g); return g; }
    double s = 0.0;
    if (S.kind == ISDF_SHAPE_BOX) s = shape_sdf_analytic(S, p);
    return shape_grad_analytic(S, p, s);
}

}  // namespace isdf
