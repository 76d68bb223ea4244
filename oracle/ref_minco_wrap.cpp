// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the reference's own MINCO (s = 3, non-uniform time), compiled UNMODIFIED from where it lies:
//   /root/reference/src/utils/include/utils/minco.hpp    BandedSystem :38-196 (factorizeLU, solve, solveAdj),
//                                                         MINCO_S3NU::setConditions / setParameters / getEnergy / getEnergyPartialGradBy* /
//                                                         propogateGrad / getTrajectory :397-655
// against oracle/_shim_dyn/Eigen/Eigen (eager element-wise arithmetic, left-to-right reductions — see that header for what this does and does
// not pin); its #include <utils/trajectory.hpp> is the reference's own header too, whose #include <utils/root_finder.hpp> resolves to the
// declarations-only stand-in oracle/_shim_dyn/utils/root_finder.hpp (listed first on the include path). The built oracle/_ref/libref_minco.so is kind "reference":
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
// getTrajectory (minco.hpp:515-529) into the reference's own Trajectory<5> (utils/trajectory.hpp, compiled unmodified): piece durations and
// 3 x 6 coefficient matrices (highest power first), column-major per piece
void ref_minco_trajectory(int N, const double *head, const double *tail, const double *inPs, const double *T, double *durations, double *cmats) {
    minco::MINCO_S3NU m;
    load(N, head, tail, inPs, T, m);
    Trajectory<5> traj;
    m.getTrajectory(traj);
    for (int i = 0; i < N; i++) {
        durations[i] = traj.getPiece(i).getDuration();
        const Trajectory<5>::CoefficientMat &cm = traj.getindexCoeffMat(i);
        for (int k = 0; k < 18; k++) cmats[18 * i + k] = cm(k);
    }
}
// MINCO -> Trajectory<5> -> Trajectory::getPos_Vel_Acc_Jerk (trajectory.hpp:565-570: locatePieceIdx :545-563 + Piece::getPos_Vel_Acc_Jerk
// :105-149) at nt absolute times; out nt x 12 (pos, vel, acc, jerk), piece nt (index locatePieceIdx returns), tloc nt (the local time it leaves)
void ref_traj_eval(int N, const double *head, const double *tail, const double *inPs, const double *T, int nt, const double *times, double *out,
                   int *piece, double *tloc, double *total_duration) {
    minco::MINCO_S3NU m;
    load(N, head, tail, inPs, T, m);
    Trajectory<5> traj;
    m.getTrajectory(traj);
    *total_duration = traj.getTotalDuration();
    for (int q = 0; q < nt; q++) {
        Eigen::Vector3d p, v, a, j;
        traj.getPos_Vel_Acc_Jerk(times[q], p, v, a, j);
        for (int k = 0; k < 3; k++) { out[12 * q + k] = p(k); out[12 * q + 3 + k] = v(k); out[12 * q + 6 + k] = a(k); out[12 * q + 9 + k] = j(k); }
        double t = times[q];
        piece[q] = traj.locatePieceIdx(t);
        tloc[q] = t;
    }
}
}
