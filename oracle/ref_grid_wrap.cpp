// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the reference's own occupancy grid, compiled UNMODIFIED from where it lies:
//   /root/reference/src/map_manager/src/Gridmap3D.cpp + include/map_manager/GridMap3D.h
//     createGridMap :25-39, isInMap :41-101, getGridIndex :135-175 (incl. its "if (iy < 0) ix = 0" branches, quirk Q6),
//     getGridCubeCenter :177-194, isIndexOccupied(int,int,int) :239-284 (out of range => occupied)
// against oracle/_shim_dyn/Eigen/Eigen (Vector3d / Vector3i element access and one vector addition) and the ROS stand-ins beside it. The built
// oracle/_ref/libref_grid.so is kind "reference": it pins orc::Grid (oracle_planner.hpp). The AABB gather itself is an inline member of
// PCSmapManager (PCSmap_manager.h:130-170: projInMap, then a triple loop over getGridIndex(corner1)..getGridIndex(corner2) calling
// isIndexOccupied and getGridCubeCenter) whose header needs PCL; ref_points_in_aabb below RESTATES those 25 lines of glue around the
// reference-compiled GridMap3D calls — the index arithmetic, the clamping quirk, the occupancy test and the centre coordinates are the reference's.
#include "map_manager/GridMap3D.h"
#include <cstdint>
#include <cmath>

static GridMap3D *make(const uint8_t *occ, int X, int Y, int Z, const double *bmin, double res) {
    GridMap3D *g = new GridMap3D();
    g->grid_resolution = res;
    g->debug_output = false;
    g->createGridMap(Vector3d(bmin[0], bmin[1], bmin[2]), Vector3d(bmin[0] + X * res, bmin[1] + Y * res, bmin[2] + Z * res));
    if (g->X_size != X || g->Y_size != Y || g->Z_size != Z) { g->releaseMemory(); delete g; return nullptr; }   // ceil((max - min) / res) rounded up
    if (occ)
        for (int i = 0; i < X; i++) for (int j = 0; j < Y; j++) for (int k = 0; k < Z; k++)
            g->grid_map[(size_t)i * Y * Z + (size_t)j * Z + k] = occ[(size_t)i * Y * Z + (size_t)j * Z + k] ? 1.0 : 0.0;
    return g;
}

extern "C" {
// idx n x 3, centre n x 3 (getGridCubeCenter of that index), inmap n
int ref_grid_index(int X, int Y, int Z, const double *bmin, double res, int n, const double *pts, int *idx, double *centre, int *inmap) {
    GridMap3D *g = make(nullptr, X, Y, Z, bmin, res);
    if (!g) return -1;
    for (int q = 0; q < n; q++) {
        const Vector3d p(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]);
        const Vector3i id = g->getGridIndex(p);
        const Vector3d c = g->getGridCubeCenter(id);
        for (int k = 0; k < 3; k++) { idx[3 * q + k] = (int)id(k); centre[3 * q + k] = c(k); }
        inmap[q] = g->isInMap(p) ? 1 : 0;
    }
    g->releaseMemory(); delete g;
    return 0;
}
// PCSmapManager::getPointsInAABB (PCSmap_manager.h:148-170) around the reference-compiled grid; returns the count, writes up to cap centres
int ref_points_in_aabb(const uint8_t *occ, int X, int Y, int Z, const double *bmin, double res, const double *centre, double half, double *out, int cap) {
    GridMap3D *g = make(occ, X, Y, Z, bmin, res);
    if (!g) return -1;
    Vector3d c1(centre[0] - half, centre[1] - half, centre[2] - half), c2(centre[0] + half, centre[1] + half, centre[2] + half);
    for (int k = 0; k < 3; k++) {                                         // projInMap (PCSmap_manager.h:130-137)
        if (c1(k) < g->boundary_xyzmin(k)) c1(k) = g->boundary_xyzmin(k);
        if (c1(k) > g->boundary_xyzmax(k)) c1(k) = g->boundary_xyzmax(k);
        if (c2(k) < g->boundary_xyzmin(k)) c2(k) = g->boundary_xyzmin(k);
        if (c2(k) > g->boundary_xyzmax(k)) c2(k) = g->boundary_xyzmax(k);
    }
    const Vector3i i1 = g->getGridIndex(c1), i2 = g->getGridIndex(c2);
    int n = 0;
    for (int i = (int)i1(0); i <= (int)i2(0); i++)
        for (int j = (int)i1(1); j <= (int)i2(1); j++)
            for (int k = (int)i1(2); k <= (int)i2(2); k++)
                if (g->isIndexOccupied(i, j, k)) {
                    const Vector3d c = g->getGridCubeCenter(i, j, k);
                    if (n < cap) { out[3 * n] = c(0); out[3 * n + 1] = c(1); out[3 * n + 2] = c(2); }
                    n++;
                }
    g->releaseMemory(); delete g;
    return n;
}
}
