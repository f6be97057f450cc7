// Host build of the per-sample math header (csrc/mppi_math.cuh) — TEST INFRASTRUCTURE ONLY.
// The product never executes this on the CPU; the `-m "not gpu"` suite uses it to check, without a GPU,
// that the very source the kernels are compiled from restates the reference's arithmetic
// (tests/test_emu_math.py compares it with the oracle).  Built with g++ -ffp-contract=off.
#include "../../pytorch_mppi_b200/csrc/mppi_math.cuh"

#include <string.h>

using namespace mppi;

template <class Model, typename real>
static void rollout(const double* model_params, const double* ext, int n_ext, int T, const double* x0, const double* v,
                    double u_scale, double* states_out, double* cost_out) {
    typename Model::template P<real> mp;
    Model::template load<real>(mp, model_params, ext, n_ext);
    real x[Model::NX];
    for (int i = 0; i < Model::NX; ++i) x[i] = (real)x0[i];
    real roll = (real)0;
    for (int t = 0; t < T; ++t) {
        real u[Model::NU];
        for (int n = 0; n < Model::NU; ++n) u[n] = Ops<real>::mul((real)u_scale, (real)v[t * Model::NU + n]);
        Model::template step<real>(mp, x, u);
        roll = Ops<real>::add(roll, Model::template cost<real>(mp, x, u));
        for (int i = 0; i < Model::NX; ++i) states_out[t * Model::NX + i] = (double)x[i];
    }
    if (Model::template has_terminal<real>(mp)) roll = Ops<real>::add(roll, Model::template terminal<real>(mp, x));
    *cost_out = (double)roll;
}

extern "C" {

void emu_philox(unsigned long long seed, unsigned long long subseq, unsigned long long offset, unsigned int* out4) {
    U4 r = philox4x32_10(seed, subseq, offset);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
void emu_normals_f32(unsigned long long seed, unsigned long long subseq, unsigned long long offset, float* out4) {
    Normals<float>::draw(seed, subseq, offset, out4);
}
void emu_normals_f64(unsigned long long seed, unsigned long long subseq, unsigned long long offset, double* out2) {
    Normals<double>::draw(seed, subseq, offset, out2);
}

// model: 1 pendulum, 2 linear point, 3 pendulum MLP ; dtype: 0 f32, 1 f64
int emu_rollout(int model, int dtype, const double* model_params, const double* ext, int n_ext, int T, const double* x0,
                const double* v, double u_scale, double* states_out, double* cost_out) {
    if (model == 1 && dtype == 0) rollout<PendulumModel, float>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else if (model == 1) rollout<PendulumModel, double>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else if (model == 2 && dtype == 0) rollout<LinearPointModel, float>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else if (model == 2) rollout<LinearPointModel, double>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else if (model == 3 && dtype == 0) rollout<PendulumMLPModel, float>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else if (model == 3) rollout<PendulumMLPModel, double>(model_params, ext, n_ext, T, x0, v, u_scale, states_out, cost_out);
    else return -1;
    return 0;
}

// colour + action-cost term for one (k,t): eps_raw = colour(z); ac = sum_n U_n (lambda g(eps) Sigma^-1)_n
double emu_colour_and_action_cost(int nu, int diag, int abs_cost, double lambda_, const double* mu, const double* L16,
                                  const double* Sinv16, const double* z, const double* eps, const double* Urow, double* eps_raw_out) {
    NoiseModel<double> nm;
    memset(&nm, 0, sizeof(nm));
    for (int i = 0; i < 4; ++i) nm.mu[i] = mu[i];
    for (int i = 0; i < 16; ++i) { nm.L[i] = L16[i]; nm.Sinv[i] = Sinv16[i]; }
    nm.lambda_ = lambda_;
    nm.diag = diag;
    nm.abs_cost = abs_cost;
    double e[4] = {0, 0, 0, 0};
    if (nu == 1) colour<double, 1>(nm, z, e);
    else if (nu == 2) colour<double, 2>(nm, z, e);
    else if (nu == 3) colour<double, 3>(nm, z, e);
    else colour<double, 4>(nm, z, e);
    for (int i = 0; i < nu; ++i) eps_raw_out[i] = e[i];
    if (nu == 1) return action_cost_term<double, 1>(nm, eps, Urow);
    if (nu == 2) return action_cost_term<double, 2>(nm, eps, Urow);
    if (nu == 3) return action_cost_term<double, 3>(nm, eps, Urow);
    return action_cost_term<double, 4>(nm, eps, Urow);
}

double emu_remainder(double a, double b) { return remainder<double>(a, b); }
float emu_remainderf(float a, float b) { return remainder<float>(a, b); }

}  // extern "C"
