// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
// Triangle-mesh signed distance as the reference obtains it from libigl 2.4.0 (third party, shipped only as
// lib/libigl.a + declaration headers; call sites Shape.cpp:85-86, 110-112, 131-133, 144-146):
//   * igl::AABB<MatrixXd,3>::squared_distance  — exact closest point on the closest triangle. Restated here as
//     Ericson's closest-point-on-triangle inside a median-split AABB tree (tree only prunes; result is the exact min).
//   * igl::fast_winding_number(fwn_bvh, 2.0, p)  — Barill et al. 2018 order-2 Barnes–Hut approximation of the
//     generalised winding number, FP32. The reference uses s = 1 - 2 w *without* thresholding (Shape.cpp:111),
//     so its SDF value carries the FWN far-field error. Oracle policy: the sign comes from the
//     *exact* winding number (sum of van Oosterom–Strackee solid angles / 4π) thresholded at 0.5 => s = ±1.
//     Modes: WN_EXACT (brute force), WN_BH (own FP64 dipole Barnes–Hut tree, thresholded — used for large CPU
//     baseline runs; validated against WN_EXACT), WN_RAW (s = 1 - 2 w_exact, unthresholded; for comparisons
//     with the compiled reference FWN header in oracle/_ref), WN_REF (s = 1 - 2 w with w from the reference-compiled FWN
//     header itself: the reference-faithful SDF value, used to MEASURE what the ±1 policy deviates by).
#pragma once
#include "oracle_math.hpp"
#include <numeric>

namespace orc {

// Ericson, Real-Time Collision Detection §5.1.5 — closest point on triangle (a,b,c) to p.
inline V3 closest_on_triangle(const V3 &p, const V3 &a, const V3 &b, const V3 &c) {
    const V3 ab = b - a, ac = c - a, ap = p - a;
    const double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) return a;
    const V3 bp = p - b;
    const double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) return b;
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) { const double v = d1 / (d1 - d3); return a + v * ab; }
    const V3 cp = p - c;
    const double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) return c;
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { const double w = d2 / (d2 - d6); return a + w * ac; }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return b + w * (c - b);
    }
    const double denom = 1.0 / (va + vb + vc);
    const double v = vb * denom, w = vc * denom;
    return a + v * ab + w * ac;
}

// signed solid angle of triangle (a,b,c) seen from q (van Oosterom & Strackee 1983)
inline double solid_angle(const V3 &q, const V3 &a, const V3 &b, const V3 &c) {
    const V3 A = a - q, B = b - q, C = c - q;
    const double la = norm(A), lb = norm(B), lc = norm(C);
    const double num = dot(A, cross(B, C));
    const double den = la * lb * lc + dot(A, B) * lc + dot(B, C) * la + dot(C, A) * lb;
    return 2.0 * std::atan2(num, den);
}

enum WindingMode { WN_EXACT = 0, WN_BH = 1, WN_RAW = 2, WN_REF = 3 };

struct Mesh {
    std::vector<V3> V;
    std::vector<int> F;  // 3 per triangle
    int wn_mode = WN_EXACT;
    // WN_REF ("reference-faithful" sign): w comes from the REFERENCE-COMPILED igl/FastWindingNumberForSoups.h (oracle/_ref/libref_fwn.so,
    // attached at run time by oracle_capi.cpp — the oracle itself never contains reference code) and s = 1 - 2 w is NOT thresholded,
    // exactly as Shape.cpp:110-111 / :131-132 / :144-145 do: FP32, order 2, accuracy scale 2.0.
    void *ref_fwn = nullptr;
    void (*ref_fwn_query)(void *, const double *, int, double, double *) = nullptr;

    // ---- distance BVH -------------------------------------------------------------------------
    struct Node { V3 lo, hi; int left, right, tri; };  // leaf: tri >= 0
    std::vector<Node> nodes;
    // ---- winding BH tree (dipole) --------------------------------------------------------------
    struct WNode { V3 centre, an; double r2; int left, right, first, count; };
    std::vector<WNode> wnodes;
    std::vector<int> worder;

    int ntri() const { return (int)F.size() / 3; }
    const V3 &vtx(int t, int k) const { return V[F[3 * t + k]]; }

    // Shape.cpp:38-50 pre-transform by poly_params: V <- Rz(yaw) Ry(pitch) Rx(roll) V + trans (degrees in)
    void pretransform(const double poly_params[6]) {
        const double PI = 3.14159265358979323846;
        const double r = poly_params[3] * PI / 180.0, p = poly_params[4] * PI / 180.0, y = poly_params[5] * PI / 180.0;
        M3 Rx = M3::identity(), Ry = M3::identity(), Rz = M3::identity();
        Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
        Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
        Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y);
        for (auto &v : V) v = mul(Rz, mul(Ry, mul(Rx, v))) + V3(poly_params[0], poly_params[1], poly_params[2]);
    }

    void build() {  // tree.init(V,F); igl::fast_winding_number(V,F,2,fwn_bvh)  (Shape.cpp:85-86)
        nodes.clear();
        std::vector<int> idx(ntri());
        std::iota(idx.begin(), idx.end(), 0);
        std::vector<V3> cen(ntri());
        for (int t = 0; t < ntri(); t++) cen[t] = (vtx(t, 0) + vtx(t, 1) + vtx(t, 2)) * (1.0 / 3.0);
        if (ntri() > 0) build_rec(idx, 0, ntri(), cen);
        build_wn();
    }
    int build_rec(std::vector<int> &idx, int b, int e, const std::vector<V3> &cen) {
        Node nd; nd.lo = V3(1e300, 1e300, 1e300); nd.hi = V3(-1e300, -1e300, -1e300); nd.left = nd.right = nd.tri = -1;
        V3 clo = nd.lo, chi = nd.hi;
        for (int i = b; i < e; i++) {
            for (int k = 0; k < 3; k++) {
                const V3 &v = vtx(idx[i], k);
                for (int a = 0; a < 3; a++) { nd.lo[a] = std::min(nd.lo[a], v[a]); nd.hi[a] = std::max(nd.hi[a], v[a]); }
            }
            for (int a = 0; a < 3; a++) { clo[a] = std::min(clo[a], cen[idx[i]][a]); chi[a] = std::max(chi[a], cen[idx[i]][a]); }
        }
        const int me = (int)nodes.size();
        nodes.push_back(nd);
        if (e - b == 1) { nodes[me].tri = idx[b]; return me; }
        int ax = 0;
        if (chi.y - clo.y > chi[ax] - clo[ax]) ax = 1;
        if (chi.z - clo.z > chi[ax] - clo[ax]) ax = 2;
        const int mid = (b + e) / 2;
        std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e,
                         [&](int x, int y) { return cen[x][ax] < cen[y][ax]; });
        const int l = build_rec(idx, b, mid, cen);
        const int r = build_rec(idx, mid, e, cen);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    static double box_d2(const V3 &p, const V3 &lo, const V3 &hi) {
        double d = 0;
        for (int a = 0; a < 3; a++) { const double e = std::max(std::max(lo[a] - p[a], p[a] - hi[a]), 0.0); d += e * e; }
        return d;
    }
    // returns squared distance; c = closest point, tri = triangle id
    double closest(const V3 &p, V3 &c, int &tri) const {
        double best = 1e300; tri = -1;
        if (nodes.empty()) return best;
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node &nd = nodes[stack[--sp]];
            if (box_d2(p, nd.lo, nd.hi) >= best) continue;
            if (nd.tri >= 0) {
                const V3 q = closest_on_triangle(p, vtx(nd.tri, 0), vtx(nd.tri, 1), vtx(nd.tri, 2));
                const V3 d = p - q; const double d2 = dot(d, d);
                if (d2 < best) { best = d2; c = q; tri = nd.tri; }
                continue;
            }
            const double dl = box_d2(p, nodes[nd.left].lo, nodes[nd.left].hi);
            const double dr = box_d2(p, nodes[nd.right].lo, nodes[nd.right].hi);
            if (dl < dr) { stack[sp++] = nd.right; stack[sp++] = nd.left; }
            else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
        }
        return best;
    }
    // same search started from bound2 instead of infinity: tri = -1 when no triangle lies within sqrt(bound2) (the "fair CPU" arm)
    double closest_within(const V3 &p, double bound2, V3 &c, int &tri) const {
        double best = bound2; tri = -1;
        if (nodes.empty()) return best;
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node &nd = nodes[stack[--sp]];
            if (box_d2(p, nd.lo, nd.hi) >= best) continue;
            if (nd.tri >= 0) {
                const V3 q = closest_on_triangle(p, vtx(nd.tri, 0), vtx(nd.tri, 1), vtx(nd.tri, 2));
                const V3 d = p - q; const double d2 = dot(d, d);
                if (d2 < best) { best = d2; c = q; tri = nd.tri; }
                continue;
            }
            const double dl = box_d2(p, nodes[nd.left].lo, nodes[nd.left].hi);
            const double dr = box_d2(p, nodes[nd.right].lo, nodes[nd.right].hi);
            if (dl < dr) { stack[sp++] = nd.right; stack[sp++] = nd.left; }
            else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
        }
        return best;
    }
    double closest_brute(const V3 &p, V3 &c, int &tri) const {
        double best = 1e300; tri = -1;
        for (int t = 0; t < ntri(); t++) {
            const V3 q = closest_on_triangle(p, vtx(t, 0), vtx(t, 1), vtx(t, 2));
            const V3 d = p - q; const double d2 = dot(d, d);
            if (d2 < best) { best = d2; c = q; tri = t; }
        }
        return best;
    }

    // ---- winding number -----------------------------------------------------------------------
    double winding_exact(const V3 &q) const {
        double om = 0;
        for (int t = 0; t < ntri(); t++) om += solid_angle(q, vtx(t, 0), vtx(t, 1), vtx(t, 2));
        return om / (4.0 * 3.14159265358979323846);
    }
    void build_wn() {
        wnodes.clear(); worder.resize(ntri());
        std::iota(worder.begin(), worder.end(), 0);
        if (ntri() > 0) build_wn_rec(0, ntri());
    }
    int build_wn_rec(int b, int e) {
        WNode nd; nd.left = nd.right = -1; nd.first = b; nd.count = e - b;
        double asum = 0; V3 c, an;
        V3 lo(1e300, 1e300, 1e300), hi(-1e300, -1e300, -1e300);
        for (int i = b; i < e; i++) {
            const int t = worder[i];
            const V3 n2 = cross(vtx(t, 1) - vtx(t, 0), vtx(t, 2) - vtx(t, 0));  // 2 * area * normal
            const double ar = 0.5 * norm(n2);
            const V3 ce = (vtx(t, 0) + vtx(t, 1) + vtx(t, 2)) * (1.0 / 3.0);
            c += ar * ce; asum += ar; an += 0.5 * n2;
            for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], ce[a]); hi[a] = std::max(hi[a], ce[a]); }
        }
        nd.centre = asum > 0 ? c * (1.0 / asum) : (lo + hi) * 0.5;
        nd.an = an;
        double r2 = 0;
        for (int i = b; i < e; i++)
            for (int k = 0; k < 3; k++) { const V3 d = vtx(worder[i], k) - nd.centre; r2 = std::max(r2, dot(d, d)); }
        nd.r2 = r2;
        const int me = (int)wnodes.size();
        wnodes.push_back(nd);
        if (e - b <= 8) return me;
        int ax = 0;
        if (hi.y - lo.y > hi[ax] - lo[ax]) ax = 1;
        if (hi.z - lo.z > hi[ax] - lo[ax]) ax = 2;
        const int mid = (b + e) / 2;
        std::nth_element(worder.begin() + b, worder.begin() + mid, worder.begin() + e, [&](int x, int y) {
            return (vtx(x, 0)[ax] + vtx(x, 1)[ax] + vtx(x, 2)[ax]) < (vtx(y, 0)[ax] + vtx(y, 1)[ax] + vtx(y, 2)[ax]);
        });
        const int l = build_wn_rec(b, mid), r = build_wn_rec(mid, e);
        wnodes[me].left = l; wnodes[me].right = r;
        return me;
    }
    // Barnes–Hut with opening criterion |q-c|^2 > beta^2 r^2 (beta = 2, as fast_winding_number(…, 2.0, …))
    double winding_bh(const V3 &q) const {
        if (wnodes.empty()) return 0;
        const double PI4 = 4.0 * 3.14159265358979323846;
        double om = 0;
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const WNode &nd = wnodes[stack[--sp]];
            const V3 d = nd.centre - q;
            const double d2 = dot(d, d);
            if (d2 > 4.0 * nd.r2 && d2 > 0) {
                om += dot(nd.an, d) / (d2 * std::sqrt(d2));  // dipole term
            } else if (nd.left < 0) {
                for (int i = nd.first; i < nd.first + nd.count; i++) {
                    const int t = worder[i];
                    om += solid_angle(q, vtx(t, 0), vtx(t, 1), vtx(t, 2));
                }
            } else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
        }
        return om / PI4;
    }
    // s = 1 - 2 w  (Shape.cpp:111) under the oracle's sign policy
    double sign(const V3 &q) const {
        if (wn_mode == WN_RAW) return 1.0 - 2.0 * winding_exact(q);
        if (wn_mode == WN_REF) {
            const double p[3] = {q.x, q.y, q.z};
            double w = 0.0;
            ref_fwn_query(ref_fwn, p, 1, 2.0, &w);
            return 1.0 - 2.0 * w;
        }
        const double w = (wn_mode == WN_BH) ? winding_bh(q) : winding_exact(q);
        return (w > 0.5) ? -1.0 : 1.0;
    }
};

}  // namespace orc
