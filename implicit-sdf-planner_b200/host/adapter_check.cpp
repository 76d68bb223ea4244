// Compiles the host-side adapters (syntax + ABI check in build()) and exports flat hooks so the python tests can drive
// them: the MINCO port against the oracle / numpy, and the full LMBM-signature callback on the GPU box.
#include "isdf_shape_adapter.hpp"
#include "isdf_cost_callback.hpp"
#include "isdf_lbfgs.hpp"
#include <cstring>

using namespace isdf_host;

// the adapters must be assignable to the reference's callback typedefs
typedef double (*lmbm_evaluate_t)(void *instance, const double *x, double *g, const int n);   // lmbm/lmbm.h:206-209
static lmbm_evaluate_t check_lmbm_signature = &BackEnd::costFunctionLmbm;

extern "C" {

int isdf_host_minco_forward(int N, const double *head, const double *tail, const double *inPs, const double *T, double *coeffs,
                            double *energy, double *gradC_energy, double *gradT_energy) {
    MincoS3 m; m.setConditions(head, tail, N); m.setParameters(inPs, T);
    std::memcpy(coeffs, m.coeffs().data(), sizeof(double) * 18 * N);
    if (energy) *energy = m.getEnergy();
    if (gradC_energy) m.getEnergyPartialGradByCoeffs(gradC_energy);
    if (gradT_energy) m.getEnergyPartialGradByTimes(gradT_energy);
    return check_lmbm_signature ? 0 : 1;
}
int isdf_host_minco_backward(int N, const double *head, const double *tail, const double *inPs, const double *T, const double *gradC,
                             const double *gradT, double *gradP, double *gradTout) {
    MincoS3 m; m.setConditions(head, tail, N); m.setParameters(inPs, T);
    m.propagateGrad(gradC, gradT, gradP, gradTout);
    return 0;
}
void *isdf_host_backend_create(isdf_ctx *ctx, int N, const double *head, const double *tail, double rho, int use_swept, int use_time_integral) {
    BackEnd *b = new BackEnd();
    b->setup(ctx, N, head, tail);
    b->rho = rho; b->use_swept = use_swept != 0; b->use_time_integral = use_time_integral != 0;
    return b;
}
double isdf_host_backend_cost(void *be, const double *x, double *g, int n) { return BackEnd::costFunctionLmbm(be, x, g, n); }
void isdf_host_backend_last(void *be, double *cost_pos, double *cost_other, double *cost_total, int *status) {
    BackEnd *b = (BackEnd *)be;
    *cost_pos = b->cost_pos; *cost_other = b->cost_other; *cost_total = b->cost_total; *status = b->last_status;
}
void isdf_host_backend_destroy(void *be) { delete (BackEnd *)be; }
void isdf_host_tau_maps(const double *tau, int n, double *T, double *tau_back) { BackEnd::forwardT(tau, T, n); BackEnd::backwardT(T, tau_back, n); }
double isdf_host_shape_sdf_grad(isdf_ctx *ctx, const double *p, double *grad) { DeviceShape s(ctx); return s.getSDFwithGrad1(p, grad); }

// L-BFGS on any raw callback (known-answer tests, CPU baseline with the oracle as the callback) ...
int isdf_host_lbfgs_generic(int n, double *x, double *f, lbfgs_eval_raw_t eval, void *instance, int mem_size, int past, double delta,
                            double g_epsilon, int max_iterations, int *iterations, int *evaluations) {
    LbfgsParams pr; pr.mem_size = mem_size; pr.past = past; pr.delta = delta; pr.g_epsilon = g_epsilon; pr.max_iterations = max_iterations;
    pr.min_step = 1.0e-32;                                   // config_CappedCone.yaml:101
    LbfgsStats st; double fx = 0;
    const int r = lbfgs_optimize(n, x, fx, eval, nullptr, nullptr, instance, pr, &st);
    *f = fx; if (iterations) *iterations = st.iterations; if (evaluations) *evaluations = st.evaluations;
    return r;
}
// ... and on the GPU-backed callback (BackEnd::costFunctionLbfgs)
int isdf_host_lbfgs_backend(void *be, double *x, int n, double *f, int mem_size, int past, double delta, double g_epsilon, int max_iterations,
                            int *iterations, int *evaluations) {
    return isdf_host_lbfgs_generic(n, x, f, &BackEnd::costFunctionLbfgs, be, mem_size, past, delta, g_epsilon, max_iterations, iterations, evaluations);
}

}  // extern "C"
