// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the reference's own flatness map, compiled UNMODIFIED from where it lies:
//   /root/reference/src/utils/include/utils/flatness.hpp
//     FlatnessMap::reset                         :36-51
//     FlatnessMap::optimizated_forward (quat)    :53-86
//     FlatnessMap::optimizated_forward (quat,omg):88-148
//     FlatnessMap::backwardthreadsafe            :230-406
// against oracle/_shim/Eigen/Eigen (element access only). The built oracle/_ref/libref_flat.so is kind "reference": it pins
// orc::Flat (oracle_math.hpp) and, through the discrete epilogue, the device flatness forward/adjoint (isdf_math.cuh).
#include <utils/flatness.hpp>

extern "C" {
// par = {mass, grav, dh, dv, cp, veps}
void ref_flat_forward_quat(const double *par, int n, const double *vel, const double *acc, const double *jer, double *quat) {
    flatness::FlatnessMap F; F.reset(par[0], par[1], par[2], par[3], par[4], par[5]);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d v, a, j; Eigen::Vector4d q;
        for (int k = 0; k < 3; k++) { v(k) = vel[3 * i + k]; a(k) = acc[3 * i + k]; j(k) = jer[3 * i + k]; }
        F.optimizated_forward(v, a, j, q);
        for (int k = 0; k < 4; k++) quat[4 * i + k] = q(k);
    }
}
void ref_flat_forward_quat_omg(const double *par, int n, const double *vel, const double *acc, const double *jer, double *quat, double *omg) {
    flatness::FlatnessMap F; F.reset(par[0], par[1], par[2], par[3], par[4], par[5]);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d v, a, j, o; Eigen::Vector4d q;
        for (int k = 0; k < 3; k++) { v(k) = vel[3 * i + k]; a(k) = acc[3 * i + k]; j(k) = jer[3 * i + k]; }
        F.optimizated_forward(v, a, j, q, o);
        for (int k = 0; k < 4; k++) quat[4 * i + k] = q(k);
        for (int k = 0; k < 3; k++) omg[3 * i + k] = o(k);
    }
}
// out: n x 12 = pos_total, vel_total, acc_total, jer_total
void ref_flat_backward(const double *par, int n, const double *vel, const double *acc, const double *jer, const double *pos_grad,
                       const double *vel_grad, const double *quat_grad, const double *omg_grad, double *out) {
    flatness::FlatnessMap F; F.reset(par[0], par[1], par[2], par[3], par[4], par[5]);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d v, a, j, pg, vg, og, tp, tv, ta, tj; Eigen::Vector4d qg;
        for (int k = 0; k < 3; k++) {
            v(k) = vel[3 * i + k]; a(k) = acc[3 * i + k]; j(k) = jer[3 * i + k];
            pg(k) = pos_grad[3 * i + k]; vg(k) = vel_grad[3 * i + k]; og(k) = omg_grad[3 * i + k];
        }
        for (int k = 0; k < 4; k++) qg(k) = quat_grad[4 * i + k];
        F.backwardthreadsafe(v, a, j, pg, vg, qg, og, tp, tv, ta, tj);
        for (int k = 0; k < 3; k++) { out[12 * i + k] = tp(k); out[12 * i + 3 + k] = tv(k); out[12 * i + 6 + k] = ta(k); out[12 * i + 9 + k] = tj(k); }
    }
}
}
