// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header): CPU restatement of the front end's attitude-kernel
// collision check — SURVEY §8f row 4. Parity unpinned by reference-owned vectors (the reference ships none for this path).
//
// Follows, in the reference tree /root/reference/src:
//   utils/include/utils/Shape.hpp:405-461        robot occupancy kernels, one per (roll, pitch) attitude: voxel (a,b,c) is set
//                                                when getonlySDF(pos, Rx(roll)*Ry(pitch)) <= max(front_end_safeh, res/2);
//   utils/include/utils/Shape.hpp:232-256        byte kernels: bit z of a row = or_mask[z % 8] (MSB first) in byte z / 8;
//   map_manager/.../PCSmap_manager.h:46-78       map kernel: occupancy inflated by (kernel_size-1)/2 free voxels per side, same packing;
//   swept_volume/.../sw_manager.hpp:775-849      kernelConv<true>: AND of the kernel's rows with the map bytes at the voxel;
//   swept_volume/.../sw_manager.hpp:852-913      visit_kernels_by_distance: level pose first, then BFS over the attitude grid;
//   swept_volume/.../sw_manager.hpp:914-941      checkKernelValue: father (roll, pitch) -> first collision-free child attitude.
// getonlySDF(pos, R_obj): analytic shapes evaluate their body at ((pos - trans) * Rotate) * R_obj (e.g. Shape.hpp:850-856),
// Ball / Point ignore R_obj (Shape.hpp:621-624, 650-653), the mesh shape evaluates at pos * R_obj (Shape.cpp:115-124).
#pragma once
#include "oracle_shapes.hpp"
#include "oracle_planner.hpp"
#include <queue>
#include <vector>
#include <cstdint>
#include <cmath>

namespace orc {

struct KernelCfg {
    double max_roll = 45, max_pitch = 45, ang_res = 9;   // degrees (config_CappedCone.yaml:62-64)
    double front_end_safeh = 0.0;
    double res = 1.0;                                      // kernelresu = occupancy_resolution (Shape.hpp:303)
    int ks = 13;                                           // kernel_size, odd
};

// Eigen::AngleAxisd(rotx, UnitX) * Eigen::AngleAxisd(roty, UnitY) assigned to a Matrix3d: both factors become quaternions
// (w = cos(a/2), vec = sin(a/2) * axis), the quaternion product is converted with toRotationMatrix().
inline M3 attitude_rotation(double rotx, double roty) {
    const double aw = std::cos(0.5 * rotx), ax = std::sin(0.5 * rotx) * 1.0, ay = std::sin(0.5 * rotx) * 0.0, az = std::sin(0.5 * rotx) * 0.0;
    const double bw = std::cos(0.5 * roty), bx = std::sin(0.5 * roty) * 0.0, by = std::sin(0.5 * roty) * 1.0, bz = std::sin(0.5 * roty) * 0.0;
    const double w = aw * bw - ax * bx - ay * by - az * bz;
    const double x = aw * bx + ax * bw + ay * bz - az * by;
    const double y = aw * by + ay * bw + az * bx - ax * bz;
    const double z = aw * bz + az * bw + ax * by - ay * bx;
    const double q[4] = {w, x, y, z};
    return quat_to_R(q);
}

struct FrontEnd {
    KernelCfg cfg;
    int xk = 0, yk = 0;
    std::vector<uint8_t> kernels;        // [xk*yk][ks^3] booleans, address a*ks*ks + b*ks + c
    std::vector<std::vector<uint8_t>> byte_kernels;   // per attitude: ks*ks*ceil(ks/8) bytes
    // inflated map kernel
    std::vector<uint8_t> map_kernel;
    int mx = 0, my = 0, mz = 0, mbytes = 0, X = 0, Y = 0, Z = 0;

    static double sdf_rot(const Shape &s, const V3 &pos, const M3 &R) {
        if (s.kind == SK_BALL || s.kind == SK_POINT) return s.sdf(pos);
        if (s.kind == SK_MESH) return s.sdf(mulT(R, pos));
        Shape body = s;                                   // body with the pre-transform already applied by hand
        body.Rotate = M3::identity(); body.trans = V3(0, 0, 0);
        return body.sdf(mulT(R, s.pre(pos)));
    }

    void build_kernels(const Shape &s, const KernelCfg &c) {
        cfg = c;
        xk = (int)std::floor(2 * c.max_roll / c.ang_res) + 1;     // sw_manager.hpp:135-136
        yk = (int)std::floor(2 * c.max_pitch / c.ang_res) + 1;
        const int ks = c.ks, n3 = ks * ks * ks;
        kernels.assign((size_t)xk * yk * n3, 0);
        byte_kernels.assign((size_t)xk * yk, std::vector<uint8_t>());
        const int size_side = (int)(0.5 * (ks - 1));
        const double safemargin = std::max(c.front_end_safeh, c.res / 2);
        int i = 0;
        for (double roll = -c.max_roll; roll <= c.max_roll && i < xk; roll += c.ang_res, i++) {
            int j = 0;
            for (double pitch = -c.max_pitch; pitch <= c.max_pitch && j < yk; pitch += c.ang_res, j++) {
                const double rotx = roll * M_PI / 180.0, roty = pitch * M_PI / 180.0;
                const M3 R = attitude_rotation(rotx, roty);
                uint8_t *k = &kernels[(size_t)(i * yk + j) * n3];
                for (int a = 0; a < ks; a++)
                    for (int b = 0; b < ks; b++)
                        for (int cc = 0; cc < ks; cc++) {
                            const double x = c.res * a - size_side * c.res, y = c.res * b - size_side * c.res, z = c.res * cc - size_side * c.res;
                            if (sdf_rot(s, V3(x, y, z), R) <= safemargin) k[a * ks * ks + b * ks + cc] = 1;
                        }
                // generateByteKernel (Shape.hpp:232-256)
                const int bl = (ks + 7) / 8;
                std::vector<uint8_t> &bk = byte_kernels[(size_t)i * yk + j];
                bk.assign((size_t)ks * ks * bl, 0);
                for (int a = 0; a < ks; a++)
                    for (int b = 0; b < ks; b++)
                        for (int cc = 0; cc < ks; cc++)
                            if (k[a * ks * ks + b * ks + cc]) bk[(size_t)a * ks * bl + b * bl + cc / 8] |= (uint8_t)(0x80u >> (cc % 8));
            }
        }
    }

    // generateMapKernel (PCSmap_manager.h:46-78); in-range occupancy only (the inflation margin stays free)
    void build_map_kernel(const uint8_t *occ, int X_, int Y_, int Z_) {
        X = X_; Y = Y_; Z = Z_;
        const int side = (cfg.ks - 1) / 2;
        mx = X + 2 * side; my = Y + 2 * side; mz = Z + 2 * side;
        mbytes = (mz + 7) / 8;
        map_kernel.assign((size_t)mx * my * mbytes + 8, 0);      // + slack: kernelConv reads one byte past its last row
        for (int x = 0; x < X; x++)
            for (int y = 0; y < Y; y++)
                for (int z = 0; z < Z; z++)
                    if (occ[((size_t)x * Y + y) * Z + z]) {
                        const int fz = z + side;
                        map_kernel[((size_t)(x + side) * my + (y + side)) * mbytes + fz / 8] |= (uint8_t)(0x80u >> (fz % 8));
                    }
    }

    // kernelConv<true> (sw_manager.hpp:821-846): true = collision-free
    bool conv(int ki, int kj, const int ind[3]) const {
        const int ks = cfg.ks, bl = (ks + 7) / 8;
        const std::vector<uint8_t> &bk = byte_kernels[(size_t)ki * yk + kj];
        for (int i = 0; i < ks; i++)
            for (int j = 0; j < ks; j++) {
                const size_t start = ((size_t)(ind[0] + i) * my + (ind[1] + j)) * mbytes + ind[2] / 8;
                const int off = ind[2] % 8;
                for (int k = 0; k < bl; k++) {
                    const uint8_t block = (uint8_t)((map_kernel[start + k] << off) | (map_kernel[start + k + 1] >> (8 - off)));
                    if (bk[(size_t)i * ks * bl + j * bl + k] & block) return false;
                }
            }
        return true;
    }

    // visit_kernels_by_distance (sw_manager.hpp:852-913)
    bool visit(int &ri, int &rj, int sx, int sy, const int ind[3], int maxdepth = 800) const {
        const int zi = (xk - 1) / 2, zj = (yk - 1) / 2;
        if (conv(zi, zj, ind)) { ri = zi; rj = zj; return true; }
        std::vector<uint8_t> visited((size_t)xk * yk, 0);
        std::queue<std::pair<int, int>> q;
        q.push({sx, sy});
        visited[(size_t)sx * yk + sy] = 1;
        const int dirs[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
        int deep = 0;
        while (!q.empty()) {
            deep++;
            const int x = q.front().first, y = q.front().second;
            q.pop();
            if (x != zi || y != zj) {
                if (conv(x, y, ind)) { ri = x; rj = y; return true; }
            }
            for (auto &d : dirs) {
                const int nx = x + d[0], ny = y + d[1];
                if (nx < 0 || nx >= xk || ny < 0 || ny >= yk) continue;
                if (visited[(size_t)nx * yk + ny]) continue;
                visited[(size_t)nx * yk + ny] = 1;
                q.push({nx, ny});
            }
            if (deep > maxdepth) return false;
        }
        return false;
    }

    // checkKernelValue (sw_manager.hpp:914-941)
    bool check(double father_roll, double father_pitch, double &child_roll, double &child_pitch, const int ind[3]) const {
        const int fi = (int)((father_roll + cfg.max_roll) / cfg.ang_res), fj = (int)((father_pitch + cfg.max_pitch) / cfg.ang_res);
        int ri = fi, rj = fj;
        if (visit(ri, rj, fi, fj, ind)) {
            child_roll = father_roll + (ri - fi) * cfg.ang_res;
            child_pitch = father_pitch + (rj - fj) * cfg.ang_res;
            return true;
        }
        return false;
    }
};

}  // namespace orc
