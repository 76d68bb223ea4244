// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
// Robot-shape implicit SDFs: the BasicShape virtual surface (Shape.hpp:469-472), the FD-gradient macro
// (Shape.hpp:32-88), the analytic shapes (Shape.hpp:603-665, 824-2391) and the mesh shape (Shape.cpp:105-151).
#pragma once
#include "oracle_math.hpp"
#include "oracle_mesh.hpp"
#include <stdexcept>
#include <string>

namespace orc {

// Numeric ids are shared with include/isdf.h (isdf_shape_kind); tests assert they match.
enum ShapeKind {
    SK_BALL = 0, SK_POINT = 1, SK_TORUS = 2, SK_CAPPED_TORUS = 3, SK_CAPPED_CONE = 4, SK_ROUNDED_CONE = 5,
    SK_WIREFRAME_BOX = 6, SK_BEND_LINEAR = 7, SK_TWIST_BOX = 8, SK_BEND_BOX = 9, SK_TABLE = 10,
    SK_TREFOIL = 11, SK_SMOOTH_DIFFERENCE = 12, SK_SMOOTH_INTERSECTION = 13, SK_CSG = 14, SK_BOX = 15,
    SK_MESH = 16
};

inline double clipd(double v, double lo, double hi) { return std::max(std::min(v, hi), lo); }
inline double box_sdf(const V3 &q) {  // the recurring "length(max(q,0)) + min(max(q.x,q.y,q.z),0)" idiom
    const double qx = std::max(q.x, 0.0), qy = std::max(q.y, 0.0), qz = std::max(q.z, 0.0);
    return std::sqrt(qx * qx + qy * qy + qz * qz) + std::min(std::max(q.x, std::max(q.y, q.z)), 0.0);
}

// Abstract surface identical in meaning to shape::BasicShape (Shape.hpp:469-472); body-frame row-vector point in.
struct Shape {
    int kind = SK_BALL;
    double par[12] = {0};  // per-kind constants (defaults = the reference's hard-coded members)
    M3 Rotate = M3::identity();  // Generalshape::Rotate / trans via getTransform (Shape.hpp:776-780)
    V3 trans;
    const Mesh *mesh = nullptr;

    // (pos_rel - trans) * Rotate with pos_rel a row vector
    V3 pre(const V3 &p) const { return mulT(Rotate, p - trans); }

    // ---- getonlySDF(pos_rel) -----------------------------------------------------------------
    double sdf(const V3 &p_in) const {
        switch (kind) {
        case SK_BALL: return norm(p_in) - par[0];  // Shape.hpp:616-619 (no pre-transform)
        case SK_POINT: return norm(p_in);          // Shape.hpp:647-650
        case SK_MESH: {                             // Shape.cpp:105-114
            V3 c; int tri;
            const double d2 = mesh->closest(p_in, c, tri);
            return mesh->sign(p_in) * std::sqrt(d2);
        }
        default: break;
        }
        const V3 p = pre(p_in);
        switch (kind) {
        case SK_TORUS: {  // Shape.hpp:843-848 (Torus 2.5/0.3), :879-884 (Torus_big 3.5/0.3)
            const double qx = std::sqrt(p.x * p.x + p.z * p.z) - par[0], qy = p.y;
            return std::sqrt(qx * qx + qy * qy) - par[1];
        }
        case SK_CAPPED_TORUS: {  // Shape.hpp:913-919; sc = (sin 40, cos 40) in radians (quirk Q9)
            const double scx = par[0], scy = par[1], ra = par[2], rb = par[3];
            const double px = std::abs(p.x);
            const double k = (scy * px > scx * p.y) ? (px * scx + p.y * scy) : std::sqrt(px * px + p.y * p.y);
            return std::sqrt((px * px + p.y * p.y + p.z * p.z) + ra * ra - 2.0 * ra * k) - rb;
        }
        case SK_CAPPED_CONE: {  // Shape.hpp:955-974; not a true distance (sqrt of sqrt, /|baba|), quirk Q9
            const double ra = par[0], rb = par[1];
            const V3 a(0, 0, -1), b(0, 0, 1);
            const double rba = rb - ra;
            const double baba = dot(b - a, b - a);
            const double papa = dot(p - a, p - a);
            const double paba = dot(p - a, b - a) / baba;
            const double x = std::sqrt(papa - paba * paba * baba);
            const double cax = std::max(0.0, x - (paba < 0.5 ? ra : rb));
            const double cay = std::abs(paba - 0.5) - 0.5;
            const double k = rba * rba + baba;
            const double f = clipd((rba * (x - ra) + paba * baba) / k, 0.0, 1.0);
            const double cbx = x - ra - f * rba;
            const double cby = paba - f;
            const double s = (cbx < 0 && cay < 0) ? -1 : 1;
            const double d = std::sqrt(std::min(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba));
            return s * std::sqrt(std::abs(d)) / std::abs(baba);
        }
        case SK_ROUNDED_CONE: {  // Shape.hpp:1018-1029
            const double r1 = par[0], r2 = par[1], h = par[2];
            const double qx = std::sqrt(p.x * p.x + p.y * p.y), qy = p.z;
            const double b = (r1 - r2) / h;
            const double a = std::sqrt(1.0 - b * b);
            const double k = -b * qx + a * qy;
            const double c1 = std::sqrt(qx * qx + qy * qy) - r1;
            const double c2 = std::sqrt(qx * qx + (qy - h) * (qy - h)) - r2;
            const double c3 = (a * qx + b * qy) - r1;
            return (k < 0) ? c1 : ((k > a * h) ? c2 : c3);
        }
        case SK_WIREFRAME_BOX: {  // Shape.hpp:1066-1085
            const double th = par[3];
            V3 ps(std::abs(p.x) - par[0] / 2 - th / 2, std::abs(p.y) - par[1] / 2 - th / 2,
                  std::abs(p.z) - par[2] / 2 - th / 2);
            V3 q(std::abs(ps.x + th / 2) - th / 2, std::abs(ps.y + th / 2) - th / 2, std::abs(ps.z + th / 2) - th / 2);
            auto g = [](double a, double b, double c) {
                const double ax = std::max(a, 0.0), bx = std::max(b, 0.0), cx = std::max(c, 0.0);
                return std::sqrt(ax * ax + bx * bx + cx * cx) + std::min(std::max(a, std::max(b, c)), 0.0);
            };
            return std::min(std::min(g(ps.x, q.y, q.z), g(q.x, ps.y, q.z)), g(q.x, q.y, ps.z));
        }
        case SK_BEND_LINEAR: {  // Shape.hpp:1141-1151 (capsule z=±2, r .25), :1206-1216 (big: ±3.2, r .45)
            const double hz = par[0], rad = par[1];
            const V3 p0(0, 0, -1), p1(0, 0, 1), v(-1, 0, 0);
            const V3 ab = p1 - p0;
            const double t = clipd(dot(p - p0, ab) / dot(ab, ab), 0, 1);
            const double u = 2 * t - 1;
            const double ea = 2 * t * t;
            const double eb = -0.5 * (u * (u - 2) - 1);
            const double e = (t < 0.5) ? ea : eb;  // in_out_quad (Shape.hpp:1113-1119)
            const V3 pp = e * v + p;
            const V3 ca(0, 0, -hz), cb(0, 0, hz);  // capsule (Shape.hpp:1129-1139)
            const V3 pa = pp - ca, ba = cb - ca;
            const double h = std::max(std::min(dot(pa, ba) / dot(ba, ba), 1.0), 0.0);
            return norm(pa - h * ba) - rad;
        }
        case SK_TWIST_BOX:  // Shape.hpp:1253-1267 (angle k*z)
        case SK_BEND_BOX: { // Shape.hpp:1307-1321 (angle k*x)
            const double k = par[3];
            const double ang = (kind == SK_TWIST_BOX) ? k * p.z : k * p.x;
            const double c = std::cos(ang), s = std::sin(ang);
            const V3 pr(c * p.x - s * p.y, s * p.x + c * p.y, p.z);
            return box_sdf(V3(std::abs(pr.x) - par[0] / 2.0, std::abs(pr.y) - par[1] / 2.0, std::abs(pr.z) - par[2] / 2.0));
        }
        case SK_TABLE: {  // Shape.hpp:1362-1381: union of two boxes, mirrored in x and y
            const V3 a1(par[0], par[1], par[2]), b1(par[3], par[4], par[5]);
            const V3 a2(par[6], par[7], par[8]), b2(par[9], par[10], par[11]);
            const V3 pp(std::abs(p.x), std::abs(p.y), p.z);
            auto one = [&](const V3 &a, const V3 &b) {
                const V3 c = (a + b) * 0.5, hf = (b - a) * 0.5;  // (a+b)/2, (b-a)/2.0
                return box_sdf(V3(std::abs(pp.x - c.x) - hf.x, std::abs(pp.y - c.y) - hf.y, std::abs(pp.z - c.z) - hf.z));
            };
            return std::min(one(a1, b1), one(a2, b2));
        }
        case SK_TREFOIL: {  // Shape.hpp:1464-1502
            const double r = 3.5, py = -p.z, dstFar = 100;
            double qx = p.x, qy = p.y;
            const double a = std::atan2(qy, qx);
            const double nq = std::sqrt(qx * qx + qy * qy);
            qx = nq - r; qy = py;
            auto rot2 = [](double &x, double &y, double ang) {
                const double c = std::cos(ang), s = std::sin(ang);
                const double nx = x * c + y * s, ny = y * c + (-x) * s;
                x = nx; y = ny;
            };
            rot2(qx, qy, 1.5 * a);
            rot2(qx, qy, -M_PI * (std::floor(std::atan2(qy, qx) / M_PI + 0.5)));
            qx -= 1.0;
            const double dx = std::abs(qx) - 0.2, dy = std::abs(qy) - 0.2;
            const double mx = std::max(dx, 0.0), my = std::max(dy, 0.0);
            const double d = (std::min(std::max(dx, dy), 0.0) + std::sqrt(mx * mx + my * my)) - 0.05;
            double dMin = dstFar;
            if (d < dMin) dMin = d;
            return 0.4 * dMin;
        }
        case SK_SMOOTH_DIFFERENCE: {  // Shape.hpp:1537-1549
            const double b = box_sdf(V3(std::abs(p.x) - par[0] / 2.0, std::abs(p.y) - par[1] / 2.0, std::abs(p.z) - par[2] / 2.0));
            const double sp = norm(p) - par[3];
            const double h = clipd(0.5 - 0.5 * (b + sp) / 0.25, 0.0, 1.0);
            const double m = b - (b + sp) * h;
            return m + 0.25 * h * (1.0 - h);
        }
        case SK_SMOOTH_INTERSECTION: {  // Shape.hpp:1592-1604 (3,3,.5 / r 1), :1647-1659 (big 9,9,1.5 / r 3)
            const double b = box_sdf(V3(std::abs(p.x) - par[0] / 2.0, std::abs(p.y) - par[1] / 2.0, std::abs(p.z) - par[2] / 2.0));
            const double sp = norm(p) - par[3];
            const double h = clipd(0.5 - 0.5 * (sp - b) / 0.25, 0.0, 1.0);
            const double m = sp + (b - sp) * h;
            return m + 0.25 * h * (1.0 - h);
        }
        case SK_CSG: {  // Shape.hpp:2286-2298: (sphere(3) ∩ box(4.5)) − (cyl ∪ rot(cyl) ∪ rot(cyl)), all k = 0
            const double d_sph = norm(p) - 3.0;
            const double d_box = box_sdf(V3(std::abs(p.x) - 4.5 / 2.0, std::abs(p.y) - 4.5 / 2.0, std::abs(p.z) - 4.5 / 2.0));
            const double f = std::max(d_sph, d_box);  // intersectionOp k=0 (Shape.hpp:2178-2196)
            // rotate_to(c, X, X) returns c itself (Shape.hpp:2043-2051); rotate_to(c,X,Y) -> rotate(c, acos(0), Y×X);
            // rotate_to(c,X,Z) -> rotate(c, acos(0), Z×X)  (Shape.hpp:2021-2059)
            const double ang = std::acos(0.0), sn = std::sin(ang), cs = std::cos(ang), mm = 1 - cs;
            auto rotmat = [&](double x, double y, double z, const V3 &v) {
                M3 R;
                R(0, 0) = mm * x * x + cs;     R(0, 1) = mm * x * y + z * sn; R(0, 2) = mm * z * x - y * sn;
                R(1, 0) = mm * x * y - z * sn; R(1, 1) = mm * y * y + cs;     R(1, 2) = mm * y * z + x * sn;
                R(2, 0) = mm * z * x + y * sn; R(2, 1) = mm * y * z - x * sn; R(2, 2) = mm * z * z + cs;
                return mul(R, v);
            };
            auto cyl = [](const V3 &v) { return std::sqrt(v.x * v.x + v.y * v.y) - 1.5; };  // Shape.hpp:1812-1821
            const double c1 = cyl(p);
            const V3 vy = normalized(cross(V3(0, 1, 0), V3(1, 0, 0)));  // b.cross(a), a = X, b = Y
            const V3 vz = normalized(cross(V3(0, 0, 1), V3(1, 0, 0)));
            const double c2 = cyl(rotmat(vy.x, vy.y, vy.z, p));
            const double c3 = cyl(rotmat(vz.x, vz.y, vz.z, p));
            const double c4 = std::min(std::min(c1, c2), c3);  // unionOp k=0 (Shape.hpp:2087-2108)
            return std::max(f, -c4);                           // differenceOp k=0 (Shape.hpp:2134-2153)
        }
        case SK_BOX:  // Shape.hpp:2343-2351 (half-extents box_x,y,z)
            return box_sdf(V3(std::abs(p.x) - par[0], std::abs(p.y) - par[1], std::abs(p.z) - par[2]));
        default: throw std::runtime_error("unknown shape kind");
        }
    }

    // ---- getonlyGrad1(pos_rel) ---------------------------------------------------------------
    V3 grad(const V3 &p) const {
        switch (kind) {
        case SK_BALL: case SK_POINT: return normalized(p);  // Shape.hpp:624-627, :656-659
        case SK_MESH: {  // Shape.cpp:125-137
            V3 c; int tri;
            mesh->closest(p, c, tri);
            const double s = mesh->sign(p);
            return normalized(s * (-(c - p)));
        }
        case SK_BOX: {  // Shape.hpp:2363-2377: one-sided FD, dx = 0.01, NOT normalised (quirk Q9)
            const double dx = 0.01, s0 = sdf(p);
            V3 t = p; t.x += dx; const double gx = sdf(t) - s0;
            t = p; t.y += dx; const double gy = sdf(t) - s0;
            t = p; t.z += dx; const double gz = sdf(t) - s0;
            return V3(gx / dx, gy / dx, gz / dx);
        }
        default: break;
        }
        // DEFINE_USEFUL_FUNCTION (Shape.hpp:33-57): central difference, dx = 5e-6, normalised
        const double dx = 0.000005;
        V3 t = p;
        t.x -= dx; double old = sdf(t); t.x += 2 * dx; const double gx = sdf(t) - old;
        t = p;
        t.y -= dx; old = sdf(t); t.y += 2 * dx; const double gy = sdf(t) - old;
        t = p;
        t.z -= dx; old = sdf(t); t.z += 2 * dx; const double gz = sdf(t) - old;
        return normalized(V3(gx / (2 * dx), gy / (2 * dx), gz / (2 * dx)));
    }

    // ---- getSDFwithGrad1(pos_rel, grad) ------------------------------------------------------
    double sdf_grad(const V3 &p, V3 &g) const {
        switch (kind) {
        case SK_BALL: g = normalized(p); return norm(p) - par[0];  // Shape.hpp:628-632
        case SK_POINT: g = normalized(p); return norm(p);          // Shape.hpp:660-664
        case SK_MESH: {                                              // Shape.cpp:139-151
            V3 c; int tri;
            const double d2 = mesh->closest(p, c, tri);
            const double s = mesh->sign(p);
            g = normalized(s * (-(c - p)));
            return s * std::sqrt(d2);
        }
        default: break;
        }
        g = grad(p);    // helperfunc (Shape.hpp:58-81) / Box::getSDFwithGrad1 (Shape.hpp:2378-2388)
        return sdf(p);
    }

    // "fair CPU" variant: the value is needed only where sdf < reach (the hinge's support), so the culls the GPU kernels
    // use are applied here as well — all EXACT, i.e. they never change the cost or the gradient:
    //   mesh     : outside the AABB inflated by reach -> inactive; BVH search bounded by reach; a point with no triangle within reach is
    //              inactive unless it is inside (winding number), in which case the full search runs;
    //   analytic : the six finite-difference evaluations of getSDFwithGrad1 only when the hinge is active.
    // Returns false when the pair is inactive (sdf >= reach).
    bool sdf_grad_reach(const V3 &p, double reach, double &s_out, V3 &g) const {
        if (kind == SK_MESH) {
            const Mesh::Node &root = mesh->nodes[0];
            for (int a = 0; a < 3; a++) if (p[a] < root.lo[a] - reach || p[a] > root.hi[a] + reach) return false;
            V3 c; int tri;
            double d2 = mesh->closest_within(p, reach * reach, c, tri);
            double s;
            if (tri < 0) {
                if (mesh->sign(p) > 0) return false;
                d2 = mesh->closest(p, c, tri); s = -1.0;
            } else s = mesh->sign(p);
            g = normalized(s * (-(c - p)));
            s_out = s * std::sqrt(d2);
            return s_out < reach;
        }
        if (kind == SK_BALL || kind == SK_POINT) { s_out = sdf_grad(p, g); return s_out < reach; }
        s_out = sdf(p);
        if (!(s_out < reach)) return false;
        g = grad(p);
        return true;
    }
};

// Factory mirroring SweptVolumeManager::shapeConstructors + initShape (swm:74-123, 255-275):
// a known analytic name gives that class with its hard-coded constants; anything else is the mesh shape.
inline bool make_named_shape(const std::string &name, Shape &s) {
    auto set = [&](int k, std::initializer_list<double> v) { s.kind = k; int i = 0; for (double d : v) s.par[i++] = d; };
    if (name == "Ball") set(SK_BALL, {1.0});                                       // Shape.hpp:609
    else if (name == "Point") set(SK_POINT, {});
    else if (name == "Torus") set(SK_TORUS, {2.5, 0.3});                          // Shape.hpp:827-828
    else if (name == "Torus_big") set(SK_TORUS, {3.5, 0.3});                      // Shape.hpp:863-864
    else if (name == "Cappedtorus") set(SK_CAPPED_TORUS, {std::sin(40), std::cos(40), 3.5, 0.3});  // Shape.hpp:898-900
    else if (name == "CappedCone") set(SK_CAPPED_CONE, {2.0, 0.8});               // Shape.hpp:936-937
    else if (name == "RoundedCone") set(SK_ROUNDED_CONE, {1.5, 0.6, 4.5});        // Shape.hpp:1003-1005
    else if (name == "WireframeBox") set(SK_WIREFRAME_BOX, {1.8, 2.5, 3.5, 0.1}); // Shape.hpp:1052-1053
    else if (name == "BendLinear") set(SK_BEND_LINEAR, {2.0, 0.25});              // Shape.hpp:1149
    else if (name == "BendLinear_big") set(SK_BEND_LINEAR, {3.2, 0.45});          // Shape.hpp:1214
    else if (name == "TwistBox") set(SK_TWIST_BOX, {2.0, 2.0, 2.0, 3.14159265358979323846 / 6});  // Shape.hpp:1239-1240
    else if (name == "BendBox") set(SK_BEND_BOX, {2.0, 2.0, 2.0, 0.5});           // Shape.hpp:1293-1294
    else if (name == "Table") set(SK_TABLE, {0, 0, 0, 3.5, 1.75, 0.7, 2.8, 1.05, 0.0, 3.5, 1.75, 2.8});  // Shape.hpp:1346-1349
    else if (name == "Trefoil") set(SK_TREFOIL, {});
    else if (name == "SmoothDifference") set(SK_SMOOTH_DIFFERENCE, {3.0, 3.0, 0.5, 1.0});    // Shape.hpp:1520-1521
    else if (name == "SmoothIntersection") set(SK_SMOOTH_INTERSECTION, {3.0, 3.0, 0.5, 1.0}); // Shape.hpp:1575-1576
    else if (name == "SmoothIntersection_big") set(SK_SMOOTH_INTERSECTION, {9.0, 9.0, 1.5, 3.0}); // Shape.hpp:1630-1631
    else if (name == "CSG") set(SK_CSG, {});
    else return false;
    return true;
}

}  // namespace orc
