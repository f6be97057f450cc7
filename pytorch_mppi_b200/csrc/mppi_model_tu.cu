// mppi_model_tu.cu — one translation unit per (registered model, dtype): the fused command kernels
// (fused_command_kernel<Model, real, V, ...>, all variants and instantiations), the resident command kernels and the
// states kernel of that model, with the host code that selects and launches them (mppi_model_host.cuh).
//
// Compiled several times by pytorch_mppi_b200/build.py:
//     nvcc ... -DMPPI_TU_MODEL=<1 pendulum | 2 linear point | 3 pendulum MLP | 100 user> -DMPPI_TU_F64=<0|1> -c
// (user model: additionally -DMPPI_USER_MODEL_HEADER="<generated header>", see models.CudaModel).
// Each unit exports ONE symbol, mppi_host::model_ops_<model>_<dtype>(), which mppi_b200.cu's dispatch calls.
#include "mppi_model_host.cuh"

#ifndef MPPI_TU_MODEL
#error "MPPI_TU_MODEL must be defined (see pytorch_mppi_b200/build.py)"
#endif
#ifndef MPPI_TU_F64
#error "MPPI_TU_F64 must be 0 or 1"
#endif

#if MPPI_TU_F64
typedef double TuReal;
#define MPPI_TU_SUFFIX f64
#else
typedef float TuReal;
#define MPPI_TU_SUFFIX f32
#endif

#if MPPI_TU_MODEL == 1
typedef mppi::PendulumModel TuModel;
#define MPPI_TU_NAME pendulum
#elif MPPI_TU_MODEL == 2
typedef mppi::LinearPointModel TuModel;
#define MPPI_TU_NAME linear_point
#elif MPPI_TU_MODEL == 3
typedef mppi::PendulumMLPModel TuModel;
#define MPPI_TU_NAME pendulum_mlp
#elif MPPI_TU_MODEL == 100
#ifndef MPPI_USER_MODEL_HEADER
#error "the user-model unit needs -DMPPI_USER_MODEL_HEADER"
#endif
typedef mppi::UserModel TuModel;
#define MPPI_TU_NAME user
#else
#error "unknown MPPI_TU_MODEL"
#endif

#define MPPI_TU_CAT2(a, b, c) model_ops_##a##_##b
#define MPPI_TU_CAT(a, b) MPPI_TU_CAT2(a, b, )
#define MPPI_TU_GETTER MPPI_TU_CAT(MPPI_TU_NAME, MPPI_TU_SUFFIX)

namespace {

int tu_run_fused(const MppiFusedParams* p, cudaStream_t s, MppiLaunchInfo* info) { return run_fused_variant<TuModel, TuReal>(p, s, info); }
int tu_build_plan(const MppiFusedParams* p, Plan* pl) { return build_plan_variant<TuModel, TuReal>(p, pl); }
int tu_run_states(const MppiFusedParams* p, const void* pa, void* states, cudaStream_t s) { return run_states<TuModel, TuReal>(p, pa, states, s); }
int tu_rollout_states(const MppiFusedParams* p, const void* x0, const void* act, long long stride, int n, int T, void* out, cudaStream_t s) {
    return run_rollout_states<TuModel, TuReal>(p, x0, act, stride, n, T, out, s);
}

const ModelOps g_ops = {tu_run_fused, tu_build_plan, tu_run_states, tu_rollout_states};

}  // namespace

namespace mppi_host {
const ModelOps* MPPI_TU_GETTER() { return &g_ops; }
}  // namespace mppi_host
