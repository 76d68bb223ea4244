// CPU-run invariant checks of the product's mesh acceleration structures (compiled with nvcc, executed on the host — no GPU needed):
// the pruning bounds used by the device searches must be LOWER bounds and the trees must cover every triangle exactly once.
//   usage: host_mesh_check <V.bin> <F.bin> nV nF   (float64 / int32 raw arrays)   prints "HOST MESH OK ..." or the first violation
#include <cstdint>
#include "../implicit-sdf-planner_b200/csrc/isdf_host_mesh.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace isdf;

static double obb_dist2(const double *o, d3 p) {   // the expression of wide_closest (isdf_mesh.cuh)
    const d3 r = mk3(p.x - o[0], p.y - o[1], p.z - o[2]);
    const double e0 = fmax(fabs(r.x * o[3] + r.y * o[4] + r.z * o[5]) - o[12], 0.0);
    const double e1 = fmax(fabs(r.x * o[6] + r.y * o[7] + r.z * o[8]) - o[13], 0.0);
    const double e2 = fmax(fabs(r.x * o[9] + r.y * o[10] + r.z * o[11]) - o[14], 0.0);
    return e0 * e0 + e1 * e1 + e2 * e2;
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int nV = atoi(argv[3]), nF = atoi(argv[4]);
    std::vector<double> V((size_t)3 * nV); std::vector<int32_t> F((size_t)3 * nF);
    FILE *f = fopen(argv[1], "rb"); if (!f || fread(V.data(), 8, V.size(), f) != V.size()) return 2; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(F.data(), 4, F.size(), f) != F.size()) return 2; fclose(f);
    const double poly[6] = {0.1, -0.2, 0.05, 120.0, 15.0, -30.0};
    HostMesh hm; std::string err;
    if (!build_host_mesh(V.data(), nV, F.data(), nF, poly, 0.866, hm, err)) { printf("build failed: %s\n", err.c_str()); return 1; }
    const DevMesh M = hm.view();
    // 1. the 32-ary tree's leaves partition the triangles
    std::vector<int> seen(hm.ntris, 0);
    int nleaf = 0;
    for (const WideNode &w : hm.wnodes)
        for (int k = 0; k < 32; k++) {
            const int ch = w.child[k];
            if (ch == WIDE_EMPTY || ch >= 0) continue;
            const int code = ~ch, first = code >> 3, cnt = (code & 7) + 1;
            nleaf++;
            for (int t = first; t < first + cnt; t++) { if (t < 0 || t >= hm.ntris) { printf("leaf out of range\n"); return 1; } seen[t]++; }
            // 2. the leaf's axis-aligned and oriented boxes contain its vertices
            const double *o = &hm.leaf_obb[(size_t)15 * first];
            for (int t = first; t < first + cnt; t++) {
                const double *T = &hm.tris[(size_t)TRI_STRIDE * t];
                const d3 vs[3] = {mk3(T[0], T[1], T[2]), mk3(T[0] + T[3], T[1] + T[4], T[2] + T[5]), mk3(T[0] + T[6], T[1] + T[7], T[2] + T[8])};
                for (const d3 &v : vs) {
                    if (obb_dist2(o, v) != 0.0) { printf("vertex outside its leaf's oriented box (tri %d)\n", t); return 1; }
                    // (triangles are stored as a, b-a, c-a: a + (b-a) can differ from the b the box was built from by one ulp)
                    const double tol = 1e-12;
                    if (v.x < w.lo[0][k] - tol || v.x > w.hi[0][k] + tol || v.y < w.lo[1][k] - tol || v.y > w.hi[1][k] + tol || v.z < w.lo[2][k] - tol || v.z > w.hi[2][k] + tol) {
                        printf("vertex outside its leaf's box (tri %d)\n", t); return 1; }
                }
            }
        }
    for (int t = 0; t < hm.ntris; t++) if (seen[t] != 1) { printf("triangle %d covered %d times\n", t, seen[t]); return 1; }
    // 3. lower-bound property at random points near and far: box bounds <= exact distance to every triangle of the leaf;
    //    and the host-compiled search (same code as the device's lane search) equals brute force
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    double worst = 0.0;
    for (int q = 0; q < 400; q++) {
        const double scale = (q % 2) ? 3.0 : 0.8;
        const d3 p = mk3(0.5 * (M.blo[0] + M.bhi[0]) + scale * U(rng), 0.5 * (M.blo[1] + M.bhi[1]) + scale * U(rng), 0.5 * (M.blo[2] + M.bhi[2]) + scale * U(rng));
        double brute = 1e300;
        for (int t = 0; t < hm.ntris; t++) {
            const double *T = &hm.tris[(size_t)TRI_STRIDE * t];
            int ft; const d3 c = tri_closest(p, mk3(T[0], T[1], T[2]), mk3(T[3], T[4], T[5]), mk3(T[6], T[7], T[8]), ft);
            const d3 e = p - c; brute = fmin(brute, dot3(e, e));
        }
        d3 cb; int tri, feat;
        const double got = mesh_closest(M, p, 1e300, cb, tri, feat);
        if (got != brute) { printf("host search %.17g != brute force %.17g\n", got, brute); return 1; }
        for (const WideNode &w : hm.wnodes)
            for (int k = 0; k < 32; k++) {
                const int ch = w.child[k];
                if (ch == WIDE_EMPTY || ch >= 0) continue;
                const int code = ~ch, first = code >> 3, cnt = (code & 7) + 1;
                double exact = 1e300;
                for (int t = first; t < first + cnt; t++) {
                    const double *T = &hm.tris[(size_t)TRI_STRIDE * t];
                    int ft; const d3 c = tri_closest(p, mk3(T[0], T[1], T[2]), mk3(T[3], T[4], T[5]), mk3(T[6], T[7], T[8]), ft);
                    const d3 e = p - c; exact = fmin(exact, dot3(e, e));
                }
                const double ob = obb_dist2(&hm.leaf_obb[(size_t)15 * first], p);
                const double bx[6] = {w.lo[0][k], w.lo[1][k], w.lo[2][k], w.hi[0][k], w.hi[1][k], w.hi[2][k]};
                const double ab = box_dist2(bx, p);
                if (ob > exact * (1 + 1e-12) + 1e-24 || ab > exact * (1 + 1e-12) + 1e-24) { printf("bound above the exact distance: obb %.17g aabb %.17g exact %.17g\n", ob, ab, exact); return 1; }
                if (exact > 0) worst = fmax(worst, (sqrt(exact) - sqrt(fmax(ob, ab))));
            }
    }
    // 4. winding-number tree (ISDF_MESH_SIGN_WINDING): the hierarchical evaluation against the brute-force sum of exact solid angles, on the
    //    closed mesh (w in {0, 1}) and on the same mesh with every 7th face removed (an open surface: fractional w, the soup case)
    double wn_worst = 0.0;
    for (int open = 0; open < 2; open++) {
        std::vector<int32_t> F2;
        for (int t = 0; t < nF; t++) if (!open || t % 7 != 3) { F2.push_back(F[3 * t]); F2.push_back(F[3 * t + 1]); F2.push_back(F[3 * t + 2]); }
        HostMesh hw; std::string e2;
        if (!build_host_mesh(V.data(), nV, F2.data(), (int)F2.size() / 3, poly, 0.866, hw, e2, 2)) { printf("winding build failed: %s\n", e2.c_str()); return 1; }
        if (hw.sign_mode != MESH_SIGN_WINDING || hw.closed != (open == 0)) { printf("sign mode / closedness wrong (open %d closed %d)\n", open, (int)hw.closed); return 1; }
        HostMesh ha; std::string e3;
        if (!build_host_mesh(V.data(), nV, F2.data(), (int)F2.size() / 3, poly, 0.866, ha, e3, 0)) { printf("auto build failed: %s\n", e3.c_str()); return 1; }
        if ((ha.sign_mode == MESH_SIGN_WINDING) != (open == 1)) { printf("automatic sign mode wrong for open = %d\n", open); return 1; }
        HostMesh hx; std::string e4;
        if (build_host_mesh(V.data(), nV, F2.data(), (int)F2.size() / 3, poly, 0.866, hx, e4, 1) != (open == 0)) { printf("exact-sign mode must refuse exactly the open mesh\n"); return 1; }
        const DevMesh W = hw.view();
        std::mt19937_64 r2(7 + open);
        std::uniform_real_distribution<double> u01(0.0, 1.0);
        auto u = [&](std::mt19937_64 &g) { return u01(g); };
        for (int it = 0; it < 400; it++) {
            const d3 q = mk3(W.blo[0] - 0.7 + (W.bhi[0] - W.blo[0] + 1.4) * u(r2), W.blo[1] - 0.7 + (W.bhi[1] - W.blo[1] + 1.4) * u(r2), W.blo[2] - 0.7 + (W.bhi[2] - W.blo[2] + 1.4) * u(r2));
            double om = 0.0;
            for (int t = 0; t < hw.ntris; t++) om += tri_solid_angle(q, &hw.tris[(size_t)TRI_STRIDE * t]);
            const double w_exact = om / (4.0 * 3.14159265358979323846), w_tree = mesh_winding(W, q);
            wn_worst = fmax(wn_worst, fabs(w_tree - w_exact));
            if (!open && fabs(w_exact - round(w_exact)) > 1e-9) { printf("closed mesh: exact winding number %.12f is not an integer\n", w_exact); return 1; }
        }
    }
    if (wn_worst > 2.5e-4) { printf("winding tree deviates from the exact winding number by %.3e\n", wn_worst); return 1; }
    printf("HOST MESH OK tris %d leaves %d wide nodes %zu max bound slack %.4f winding tree max error %.2e\n", hm.ntris, nleaf, hm.wnodes.size(), worst, wn_worst);
    return 0;
}
