// Device-side triangle-mesh signed distance: the Generalshape mesh path (Shape.cpp:105-151), which the reference
// evaluates with libigl's AABB tree (exact closest point) and fast winding number (sign).
//
// B200 layout: a binary BVH whose nodes carry BOTH children's boxes (one 112-byte fetch decides the descent),
// leaves of <= 4 triangles stored in leaf order as (a, b-a, c-a), 7 pseudonormals per triangle (face, 3 edges,
// 3 vertices) for the sign, and a coarse body-frame inside/outside bitmap that resolves the sign of points that
// have no triangle within the query bound. The whole structure for a 4000-triangle robot is < 2 MB: L2-resident.
//
// Sign: for a closed, consistently oriented mesh the generalised winding number is exactly 0 outside / 1 inside,
// so s = 1 - 2w = ±1 (Shape.cpp:111). The device gets the same ±1 from the angle-weighted pseudonormal at the
// closest feature (Baerentzen & Aanaes 2005) — no O(#triangles) solid-angle sum per query.
#pragma once
#include "isdf_math.cuh"

namespace isdf {

struct BvhNode {       // 112 bytes
    double lbox[6];    // left child lo(3), hi(3)
    double rbox[6];    // right child
    int left, right;   // >= 0: node index; < 0: leaf, ~code with code = first_tri*4 + (count-1)
    int pad0, pad1;
};

struct DevMesh {
    const BvhNode *nodes;   // nodes[0] is the root (a single-leaf mesh still gets one node with right = left)
    const double *tris;     // 9 doubles per triangle in leaf order: a, ab, ac
    const double *pnormals; // 21 doubles per triangle: face, edge ab, edge bc, edge ca, vertex a, b, c
    const uint32_t *inside; // sign bitmap over [glo, glo + gdim*gcell), bit = 1 -> cell centre inside
    int ntris;
    int gdim[3];
    double glo[3];
    double gcell;           // cell edge; cell diagonal < sign_radius
    double sign_radius;     // a point with no triangle within this distance shares its cell's sign (0 = bitmap unusable)
    double blo[3], bhi[3];  // mesh AABB
};

ISDF_HD double box_dist2(const double *b, d3 p) {
    const double ex = fmax(fmax(b[0] - p.x, p.x - b[3]), 0.0);
    const double ey = fmax(fmax(b[1] - p.y, p.y - b[4]), 0.0);
    const double ez = fmax(fmax(b[2] - p.z, p.z - b[5]), 0.0);
    return ex * ex + ey * ey + ez * ez;
}

// Closest point on triangle (a, a+ab, a+ac) by Voronoi-region classification; feat: 0 face, 1 ab, 2 bc, 3 ca, 4 a, 5 b, 6 c
ISDF_HD d3 tri_closest(d3 p, d3 a, d3 ab, d3 ac, int &feat) {
    const d3 ap = p - a;
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) { feat = 4; return a; }
    const d3 bp = ap - ab;
    const double d3_ = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3_ >= 0.0 && d4 <= d3_) { feat = 5; return a + ab; }
    const double vc = d1 * d4 - d3_ * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3_ <= 0.0) { feat = 1; return a + (d1 / (d1 - d3_)) * ab; }
    const d3 cp = ap - ac;
    const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { feat = 6; return a + ac; }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { feat = 3; return a + (d2 / (d2 - d6)) * ac; }
    const double va = d3_ * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3_) >= 0.0 && (d5 - d6) >= 0.0) {
        feat = 2;
        const double w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        return (a + ab) + w * (ac - ab);
    }
    feat = 0;
    const double den = 1.0 / (va + vb + vc);
    return a + (vb * den) * ab + (vc * den) * ac;
}

// Exact nearest triangle within sqrt(bound2). Returns squared distance (bound2 if none found; tri = -1).
__host__ __device__ inline double mesh_closest(const DevMesh &M, d3 p, double bound2, d3 &cbest, int &tri, int &feat) {
    double best = bound2;
    tri = -1; feat = 0;
    int stack[48];
    double stack_d2[48];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const BvhNode *nd = M.nodes + cur;
            const double dl = box_dist2(nd->lbox, p), dr = box_dist2(nd->rbox, p);
            const int l = nd->left, r = nd->right;
            const bool hl = dl < best, hr = (dr < best) && (r != l);
            if (hl && hr) {
                if (dl <= dr) { stack[sp] = r; stack_d2[sp++] = dr; cur = l; } else { stack[sp] = l; stack_d2[sp++] = dl; cur = r; }
                continue;
            }
            if (hl) { cur = l; continue; }
            if (hr) { cur = r; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 2, cnt = (code & 3) + 1;
            for (int t = first; t < first + cnt; t++) {
                const double *T = M.tris + 9 * (size_t)t;
                int f;
                const d3 q = tri_closest(p, mk3(T[0], T[1], T[2]), mk3(T[3], T[4], T[5]), mk3(T[6], T[7], T[8]), f);
                const d3 e = p - q;
                const double dd = dot3(e, e);
                if (dd < best) { best = dd; cbest = q; tri = t; feat = f; }
            }
        }
        // pop the next subtree that can still beat the current best
        do {
            if (sp == 0) return best;
            cur = stack[--sp];
        } while (!(stack_d2[sp] < best));
    }
}

ISDF_HD bool mesh_cell_inside(const DevMesh &M, d3 p) {
    const int ix = (int)floor((p.x - M.glo[0]) / M.gcell), iy = (int)floor((p.y - M.glo[1]) / M.gcell), iz = (int)floor((p.z - M.glo[2]) / M.gcell);
    if (ix < 0 || iy < 0 || iz < 0 || ix >= M.gdim[0] || iy >= M.gdim[1] || iz >= M.gdim[2]) return false;
    const size_t bit = ((size_t)ix * M.gdim[1] + iy) * M.gdim[2] + iz;
    return (M.inside[bit >> 5] >> (bit & 31)) & 1u;
}

// getSDFwithGrad1 for the mesh shape (Shape.cpp:139-151): sdf = s * dist, grad = normalise(s * (p - c)).
// `reach` prunes the search: the caller only needs the value when sdf < reach (pass 1e300 for "always").
// If nothing lies within reach the point is either farther out than reach (returns +reach, grad untouched => the
// hinge is inactive) or deep inside (full search).
__host__ __device__ inline double mesh_sdf_grad(const DevMesh &M, d3 p, double reach, d3 &g) {
    d3 c = mk3(0, 0, 0);
    int tri, feat;
    double d2;
    const bool bounded = (reach < 1e150) && (M.sign_radius > 0.0) && (reach >= M.sign_radius);
    if (bounded) {
        d2 = mesh_closest(M, p, reach * reach, c, tri, feat);
        if (tri < 0) {
            if (!mesh_cell_inside(M, p)) return reach;   // outside and at least `reach` away
            d2 = mesh_closest(M, p, 1e300, c, tri, feat); // deep inside: rare
        }
    } else {
        d2 = mesh_closest(M, p, 1e300, c, tri, feat);
    }
    const d3 e = p - c;
    const double *pn = M.pnormals + 21 * (size_t)tri + 3 * feat;
    double side = e.x * pn[0] + e.y * pn[1] + e.z * pn[2];
    if (side == 0.0) {  // on the surface or numerically tangent: fall back to the face normal
        const double *fn = M.pnormals + 21 * (size_t)tri;
        side = e.x * fn[0] + e.y * fn[1] + e.z * fn[2];
    }
    const double s = (side < 0.0) ? -1.0 : 1.0;
    g = unit3(s * e);
    return s * sqrt(d2);
}

}  // namespace isdf
