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
#include <functional>
#include <array>
using namespace isdf;

static double obb_dist2(const double *o, d3 p) {
    const d3 r = mk3(p.x - o[0], p.y - o[1], p.z - o[2]);
    const double e0 = fmax(fabs(r.x * o[3] + r.y * o[4] + r.z * o[5]) - o[12], 0.0);
    const double e1 = fmax(fabs(r.x * o[6] + r.y * o[7] + r.z * o[8]) - o[13], 0.0);
    const double e2 = fmax(fabs(r.x * o[9] + r.y * o[10] + r.z * o[11]) - o[14], 0.0);
    return e0 * e0 + e1 * e1 + e2 * e2;
}
struct Count { long nodes = 0, passes = 0, tris = 0, leaves = 0; };

// oriented box of a set of triangles, same recipe as the leaves' (isdf_host_mesh.cuh): axis 0 = area-weighted mean normal
static void make_obb(const HostMesh &hm, const std::vector<int> &tris, double *o) {
    double n[3] = {0, 0, 0};
    std::vector<std::array<double, 3>> pts;
    for (int t : tris) {
        const double *T = &hm.tris[(size_t)TRI_STRIDE * t];
        pts.push_back({T[0], T[1], T[2]}); pts.push_back({T[0] + T[3], T[1] + T[4], T[2] + T[5]}); pts.push_back({T[0] + T[6], T[1] + T[7], T[2] + T[8]});
        n[0] += T[4] * T[8] - T[5] * T[7]; n[1] += T[5] * T[6] - T[3] * T[8]; n[2] += T[3] * T[7] - T[4] * T[6];
    }
    double ax[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (nl > 1e-12) {
        for (int a = 0; a < 3; a++) n[a] /= nl;
        // in-plane axis: direction of largest spread of the points projected into the plane (power iteration on the 2x2-in-3D covariance)
        double c[3] = {0, 0, 0};
        for (auto &q : pts) for (int a = 0; a < 3; a++) c[a] += q[a] / pts.size();
        double u[3] = {n[1], -n[0], 0};
        if (fabs(u[0]) + fabs(u[1]) < 1e-9) { u[0] = 1; u[1] = 0; u[2] = 0; }
        for (int it = 0; it < 30; it++) {
            double v[3] = {0, 0, 0};
            for (auto &q : pts) {
                double d[3] = {q[0] - c[0], q[1] - c[1], q[2] - c[2]};
                const double dn = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
                for (int a = 0; a < 3; a++) d[a] -= dn * n[a];
                const double s = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
                for (int a = 0; a < 3; a++) v[a] += s * d[a];
            }
            const double vl = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (vl < 1e-300) break;
            for (int a = 0; a < 3; a++) u[a] = v[a] / vl;
        }
        const double w[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
        for (int a = 0; a < 3; a++) { ax[0][a] = n[a]; ax[1][a] = u[a]; ax[2][a] = w[a]; }
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (auto &q : pts) for (int k = 0; k < 3; k++) { const double pr = q[0] * ax[k][0] + q[1] * ax[k][1] + q[2] * ax[k][2]; lo[k] = std::min(lo[k], pr); hi[k] = std::max(hi[k], pr); }
    for (int a = 0; a < 3; a++) { o[a] = 0; for (int k = 0; k < 3; k++) o[a] += 0.5 * (lo[k] + hi[k]) * ax[k][a]; }
    for (int k = 0; k < 3; k++) { for (int a = 0; a < 3; a++) o[3 + 3 * k + a] = ax[k][a]; o[12 + k] = 0.5 * (hi[k] - lo[k]) + 1e-9; }
}
static std::vector<double> g_node_obb;   // 15 per wide node (index = wide node id), built in main
static bool g_use_node_obb = false;

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
                if (g_use_node_obb && alive[k] && ch >= 0 && ch != WIDE_EMPTY) { d2[k] = fmax(d2[k], obb_dist2(&g_node_obb[(size_t)15 * ch], p)); alive[k] = d2[k] < best; }
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
    // oriented boxes of the internal nodes
    g_node_obb.assign((size_t)15 * hm.wnodes.size(), 0.0);
    {
        std::function<void(int, std::vector<int> &)> collect = [&](int nd, std::vector<int> &out) {
            for (int k = 0; k < 32; k++) {
                const int ch = hm.wnodes[nd].child[k];
                if (ch == WIDE_EMPTY) continue;
                if (ch < 0) { const int code = ~ch, first = code >> 3, cnt = (code & 7) + 1; for (int t = first; t < first + cnt; t++) out.push_back(t); }
                else collect(ch, out);
            }
        };
        for (size_t nd = 0; nd < hm.wnodes.size(); nd++) { std::vector<int> tr; collect((int)nd, tr); make_obb(hm, tr, &g_node_obb[15 * nd]); }
    }
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
            const double r0 = sim(hm, p, -1, false, a, t1), r1 = sim(hm, p, tn, false, b, t2), r3 = sim(hm, p, -1, true, d_, t4);
            g_use_node_obb = true; const double r2 = sim(hm, p, tn, false, c, t3); g_use_node_obb = false;
            if (r0 != r1 || r0 != r2 || r0 != r3) { printf("MISMATCH\n"); return 1; }
            n++;
        }
        auto pr = [&](const char *nm, const Count &C) { printf("  %-28s nodes %.2f  leaf passes %.2f  alive leaves %.1f  triangle tests %.1f\n", nm, C.nodes / (double)n, C.passes / (double)n, C.leaves / (double)n, C.tris / (double)n); };
        printf("distance %.1f m:\n", dist);
        pr("no seed", a); pr("no seed, leaves sorted", d_); pr("seed = neighbour's triangle", b); pr("seed + internal-node OBBs", c);
    }
    return 0;
}
