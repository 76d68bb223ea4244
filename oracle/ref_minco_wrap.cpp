// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the reference's own MINCO (s = 3, non-uniform time), compiled UNMODIFIED from where it lies:
//   /root/reference/src/utils/include/utils/minco.hpp    BandedSystem :38-196 (factorizeLU, solve, solveAdj),
//                                                         MINCO_S3NU::setConditions / setParameters / getEnergy / getEnergyPartialGradBy* /
//                                                         propogateGrad / getTrajectory :397-655
// against oracle/_shim_dyn/Eigen/Eigen (eager element-wise arithmetic, left-to-right reductions — see that header for what this does and does
// not pin) and oracle/_shim_dyn/utils/trajectory.hpp (a container, nothing else). The built oracle/_ref/libref_minco.so is kind "reference":
// it pins orc::Minco (oracle_minco.hpp) and the product's host MINCO (host/isdf_minco.hpp), to which the device kernels are tested bit-identical.
// Same C signatures as orc_minco_forward / orc_minco_backward (oracle_capi.cpp): head / tail 3 x 3 column-major with columns pos, vel, acc;
// inPs 3 x (N-1) column-major; coefficient blocks 6N x 3 column-major.
#include <utils/minco.hpp>

static void load(int N, const double *head, const double *tail, const double *inPs, const double *T, minco::MINCO_S3NU &m) {
    Eigen::Matrix3d h, t;
    for (int k = 0; k < 9; k++) { h(k) = head[k]; t(k) = tail[k]; }
    Eigen::Matrix3Xd P(3, N - 1);
    for (int k = 0; k < 3 * (N - 1); k++) P(k) = inPs[k];
    Eigen::VectorXd ts(N);
    for (int k = 0; k < N; k++) ts(k) = T[k];
    m.setConditions(h, t, N);
    m.setParameters(P, ts);
}

extern "C" {
void ref_minco_forward(int N, const double *head, const double *tail, const double *inPs, const double *T, double *coeffs, double *energy,
                       double *gdC, double *gdT) {
    minco::MINCO_S3NU m;
    load(N, head, tail, inPs, T, m);
    const Eigen::MatrixX3d &b = m.getCoeffs();
    for (int k = 0; k < 18 * N; k++) coeffs[k] = b(k);
    m.getEnergy(*energy);
    Eigen::MatrixX3d gc; Eigen::VectorXd gt;
    m.getEnergyPartialGradByCoeffs(gc);
    m.getEnergyPartialGradByTimes(gt);
    for (int k = 0; k < 18 * N; k++) gdC[k] = gc(k);
    for (int k = 0; k < N; k++) gdT[k] = gt(k);
}
void ref_minco_backward(int N, const double *head, const double *tail, const double *inPs, const double *T, const double *gradC, const double *gradT,
                        double *gradP, double *gradTout) {
    minco::MINCO_S3NU m;
    load(N, head, tail, inPs, T, m);
    Eigen::MatrixX3d gc(6 * N, 3); Eigen::VectorXd gt(N);
    for (int k = 0; k < 18 * N; k++) gc(k) = gradC[k];
    for (int k = 0; k < N; k++) gt(k) = gradT[k];
    Eigen::Matrix3Xd gp; Eigen::VectorXd go;
    m.propogateGrad(gc, gt, gp, go);
    for (int k = 0; k < 3 * (N - 1); k++) gradP[k] = gp(k);
    for (int k = 0; k < N; k++) gradTout[k] = go(k);
}
// getTrajectory: piece durations and 3 x 6 coefficient matrices (highest power first), column-major per piece
void ref_minco_trajectory(int N, const double *head, const double *tail, const double *inPs, const double *T, double *durations, double *cmats) {
    minco::MINCO_S3NU m;
    load(N, head, tail, inPs, T, m);
    Trajectory<5> traj;
    m.getTrajectory(traj);
    for (int i = 0; i < N; i++) { durations[i] = traj.durations[i]; for (int k = 0; k < 18; k++) cmats[18 * i + k] = traj.coeffs[i](k); }
}
}
