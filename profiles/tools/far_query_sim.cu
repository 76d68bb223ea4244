// Host emulation of wide_closest (isdf_mesh.cuh) counting the work of one warp-cooperative query: 32-ary nodes opened, passes of 8 leaves,
// triangle tests — for far-field queries (the swept-volume path: 0.9 .. 1.9 m from the surface) with (a) no seed, (b) a seed = the closest
// triangle of a NEIGHBOURING query (2 cm away), (c) nodes opened nearest-first at every level + seed. Compiled with nvcc for the host.
//   usage: far_query_sim <V.bin> <F.bin> nV nF
#include <cstdint>
#include "../../implicit-sdf-planner_b200/csrc/isdf_host_mesh.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
using namespace isdf;

static double obb_dist2(const double *o, d3 p) {
    const d3 r = mk3(p.x - o[0], p.y - o[1], p.z - o[2]);
    const double e0 = fmax(fabs(r.x * o[3] + r.y * o[4] + r.z * o[5]) - o[12], 0.0);
    const double e1 = fmax(fabs(r.x * o[6] + r.y * o[7] + r.z * o[8]) - o[13], 0.0);
    const double e2 = fmax(fabs(r.x * o[9] + r.y * o[10] + r.z * o[11]) - o[14], 0.0);
    return e0 * e0 + e1 * e1 + e2 * e2;
}
struct Count { long nodes = 0, passes = 0, tris = 0, leaves = 0; };

static double tri_d2(const HostMesh &hm, d3 p, int t) {
    const double *T = &hm.tris[(size_t)TRI_STRIDE * t];
    int ft; const d3 c = tri_closest(p, mk3(T[0], T[1], T[2]), mk3(T[3], T[4], T[5]), mk3(T[6], T[7], T[8]), ft);
    const d3 e = p - c; return dot3(e, e);
}

// sort_leaves: leaves of a node tested nearest-first (device: in slot order)
static double sim(const HostMesh &hm, d3 p, int seed, bool sort_leaves, Count &C, int &tri_out) {
    double best = 1e300; tri_out = -1;
    if (seed >= 0) { best = tri_d2(hm, p, seed); tri_out = seed; }
    std::vector<std::pair<int, double>> stack;
    int cur = 0; double cur_d2 = 0;
    for (;;) {
        if (cur_d2 < best) {
            C.nodes++;
            const WideNode &nd = hm.wnodes[cur];
            double d2[32]; bool alive[32];
            std::vector<std::pair<double, int>> leaves;
            for (int k = 0; k < 32; k++) {
                const int ch = nd.child[k];
                const double ex = fmax(fmax(nd.lo[0][k] - p.x, p.x - nd.hi[0][k]), 0.0), ey = fmax(fmax(nd.lo[1][k] - p.y, p.y - nd.hi[1][k]), 0.0), ez = fmax(fmax(nd.lo[2][k] - p.z, p.z - nd.hi[2][k]), 0.0);
                d2[k] = ex * ex + ey * ey + ez * ez;
                alive[k] = ch != WIDE_EMPTY && d2[k] < best;
                if (alive[k] && ch < 0) { d2[k] = fmax(d2[k], obb_dist2(&hm.leaf_obb[(size_t)15 * ((~ch) >> 3)], p)); alive[k] = d2[k] < best; }
                if (alive[k] && ch < 0) leaves.push_back({d2[k], ch});
            }
            if (sort_leaves) std::sort(leaves.begin(), leaves.end());
            for (size_t base = 0; base < leaves.size(); base += 8) {
                bool any = false;
                for (size_t i = base; i < std::min(leaves.size(), base + 8); i++) {
                    if (!(leaves[i].first < best)) continue;
                    any = true; C.leaves++;
                }
                if (!any) continue;
                C.passes++;
                double nb = best; int nt = tri_out;
                for (size_t i = base; i < std::min(leaves.size(), base + 8); i++) {
                    if (!(leaves[i].first < best)) continue;
                    const int code = ~leaves[i].second, first = code >> 3, cnt = (code & 7) + 1;
                    for (int t = first; t < first + cnt; t++) { C.tris++; const double d = tri_d2(hm, p, t); if (d < nb) { nb = d; nt = t; } }
                }
                best = nb; tri_out = nt;
            }
            std::vector<std::pair<double, int>> inner;
            for (int k = 0; k < 32; k++) { const int ch = nd.child[k]; if (alive[k] && ch >= 0 && ch != WIDE_EMPTY && d2[k] < best) inner.push_back({d2[k], ch}); }
            std::sort(inner.begin(), inner.end());
            for (int i = (int)inner.size() - 1; i >= 0; i--) stack.push_back({inner[i].second, inner[i].first});   // farthest first
        }
        if (stack.empty()) return best;
        cur = stack.back().first; cur_d2 = stack.back().second; stack.pop_back();
    }
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int nV = atoi(argv[3]), nF = atoi(argv[4]);
    std::vector<double> V((size_t)3 * nV); std::vector<int32_t> F((size_t)3 * nF);
    FILE *f = fopen(argv[1], "rb"); if (!f || fread(V.data(), 8, V.size(), f) != V.size()) return 2; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(F.data(), 4, F.size(), f) != F.size()) return 2; fclose(f);
    const double poly[6] = {0, 0, 0, 120.0, 0, 0};
    HostMesh hm; std::string err;
    if (!build_host_mesh(V.data(), nV, F.data(), nF, poly, 0.866, hm, err)) { printf("build failed: %s\n", err.c_str()); return 1; }
    const DevMesh M = hm.view();
    printf("tris %d wide nodes %zu  bbox [%.2f %.2f %.2f]..[%.2f %.2f %.2f] gpad %.3f gcell %.3f\n", hm.ntris, hm.wnodes.size(), M.blo[0], M.blo[1], M.blo[2], M.bhi[0], M.bhi[1], M.bhi[2], M.gpad, M.gcell);
    std::mt19937_64 rng(5);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    for (double dist : {0.3, 0.9, 1.4, 1.9}) {
        Count a, b, c, d_; int n = 0;
        while (n < 300) {
            d3 dir = mk3(U(rng), U(rng), U(rng));
            const double l = sqrt(dot3(dir, dir)); if (l > 1 || l < 0.1) continue;
            const d3 ctr = mk3(0.5 * (M.blo[0] + M.bhi[0]), 0.5 * (M.blo[1] + M.bhi[1]), 0.5 * (M.blo[2] + M.bhi[2]));
            // walk outwards until the distance to the mesh is `dist`
            d3 p = ctr; double s = 0;
            int t0;
            Count tmp;
            for (int it = 0; it < 60; it++) { p = ctr + (s / l) * dir; const double dd = sqrt(sim(hm, p, -1, false, tmp, t0)); if (dd >= dist && s > 0.2) break; s += fmax(0.02, 0.7 * (dist - dd)); }
            const d3 pn = p + mk3(0.012, -0.009, 0.011);   // the neighbouring query (~2 cm away)
            int tn; sim(hm, pn, -1, false, tmp, tn);
            int t1, t2, t3, t4;
            const double r0 = sim(hm, p, -1, false, a, t1), r1 = sim(hm, p, tn, false, b, t2), r2 = sim(hm, p, tn, true, c, t3), r3 = sim(hm, p, -1, true, d_, t4);
            if (r0 != r1 || r0 != r2 || r0 != r3) { printf("MISMATCH\n"); return 1; }
            n++;
        }
        auto pr = [&](const char *nm, const Count &C) { printf("  %-28s nodes %.2f  leaf passes %.2f  alive leaves %.1f  triangle tests %.1f\n", nm, C.nodes / (double)n, C.passes / (double)n, C.leaves / (double)n, C.tris / (double)n); };
        printf("distance %.1f m:\n", dist);
        pr("no seed", a); pr("no seed, leaves sorted", d_); pr("seed = neighbour's triangle", b); pr("seed + leaves sorted", c);
    }
    return 0;
}
