// Host-side adapter presenting the device SDF behind the reference's shape plug-in surface:
//   class BasicShape { virtual double getonlySDF(const RowVector3d&); virtual double getonlySDF(const RowVector3d&, const Matrix3d& R_obj);
//                      virtual Vector3d getonlyGrad1(const RowVector3d&); virtual double getSDFwithGrad1(const RowVector3d&, Vector3d&); }
//   (utils/include/utils/Shape.hpp:469-472). Input is a body-frame point; gradients are unit length except for Box.
// Two faces: raw-pointer methods that compile anywhere, and — when Eigen is available (ISDF_WITH_EIGEN) — the exact
// reference signatures, so `class DeviceShape : public shape::BasicShape` in the reference tree is a three-line forwarder
// (INTEGRATION.md). One-point virtual calls cost a kernel launch each (they exist for interface parity, e.g. the
// visualisation and A* callers); the hot loops never go through them — they call isdf_eval_* once per optimiser step.
#pragma once
#include "isdf.h"
#include <stdexcept>
#include <string>
#include <vector>
#ifdef ISDF_WITH_EIGEN
#include <Eigen/Core>
#endif

namespace isdf_host {

class DeviceShape {
public:
    explicit DeviceShape(isdf_ctx *ctx) : ctx_(ctx) {}

    double getonlySDF(const double p[3]) const { double s; query(p, 1, &s, nullptr, ISDF_QUERY_SDF); return s; }
    // getonlySDF(pos, R_obj) evaluates the SDF at pos * R_obj (row vector times matrix; Shape.hpp:847-848) — R_obj row-major
    double getonlySDF(const double p[3], const double R_obj[9]) const {
        const double q[3] = {p[0] * R_obj[0] + p[1] * R_obj[3] + p[2] * R_obj[6], p[0] * R_obj[1] + p[1] * R_obj[4] + p[2] * R_obj[7],
                             p[0] * R_obj[2] + p[1] * R_obj[5] + p[2] * R_obj[8]};
        return getonlySDF(q);
    }
    void getonlyGrad1(const double p[3], double grad[3]) const { query(p, 1, nullptr, grad, ISDF_QUERY_GRAD); }
    double getSDFwithGrad1(const double p[3], double grad[3]) const { double s; query(p, 1, &s, grad, ISDF_QUERY_SDF_GRAD); return s; }

    // batched forms (n x 3 row-major) — what a caller with many points should use
    void getSDFwithGrad1(const double *p, int n, double *sdf, double *grad) const { query(p, n, sdf, grad, ISDF_QUERY_SDF_GRAD); }

#ifdef ISDF_WITH_EIGEN
    double getonlySDF(const Eigen::RowVector3d &p) const { return getonlySDF(p.data()); }
    double getonlySDF(const Eigen::RowVector3d &p, const Eigen::Matrix3d &R_obj) const { const Eigen::RowVector3d q = p * R_obj; return getonlySDF(q.data()); }
    Eigen::Vector3d getonlyGrad1(const Eigen::RowVector3d &p) const { Eigen::Vector3d g; getonlyGrad1(p.data(), g.data()); return g; }
    double getSDFwithGrad1(const Eigen::RowVector3d &p, Eigen::Vector3d &g) const { return getSDFwithGrad1(p.data(), g.data()); }
#endif

private:
    void query(const double *p, int n, double *sdf, double *grad, int what) const {
        const int r = isdf_shape_query(ctx_, p, n, sdf, grad, what);
        if (r != ISDF_OK) throw std::runtime_error(std::string("isdf_shape_query: ") + isdf_last_error());
    }
    isdf_ctx *ctx_;
};

}  // namespace isdf_host
