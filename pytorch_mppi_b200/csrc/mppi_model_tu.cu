// mppi_model_tu.cu — one translation unit per (registered model, dtype): the fused command kernels
// (fused_command_kernel<Model, real, V, ...>, all variants and instantiations), the resident command kernels and the
// states kernel of that model — and, for the fp32 MLP, the tcgen05 kernels.  The unit contains no launch logic: it
// exports ONE symbol, mppi_host::model_kernels_<model>_<dtype>(), a table of kernel handles plus the parameter packer,
// which the generic host code in mppi_b200.cu (mppi_fused_host.cuh) works on.
//
// Compiled several times by pytorch_mppi_b200/build.py:
//     nvcc ... -DMPPI_TU_MODEL=<1 pendulum | 2 linear point | 3 pendulum MLP | 100 user> -DMPPI_TU_F64=<0|1> -c
// (user model built with nvcc: additionally -DMPPI_USER_MODEL_HEADER="<generated header>"; the default route for user
// models is NVRTC at run time, mppi_user_model_register — same kernels, same table).
#include "mppi_host.cuh"
#include "mppi_mlp_tc.cuh"

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

#define MPPI_TU_CAT2(a, b, c) model_kernels_##a##_##b
#define MPPI_TU_CAT(a, b) MPPI_TU_CAT2(a, b, )
#define MPPI_TU_GETTER MPPI_TU_CAT(MPPI_TU_NAME, MPPI_TU_SUFFIX)

namespace {

constexpr bool kIsMlp = MPPI_TU_MODEL == 3;
typedef TuModel::P<TuReal> TuParams;
static_assert(sizeof(TuParams) <= MPPI_MODEL_BLOCK_BYTES, "model parameter block too large");

void tu_load(const ModelKernels*, void* dst, const double* blob, const double* ext, int n_ext) {
    TuModel::load<TuReal>(*reinterpret_cast<TuParams*>(dst), blob, ext, n_ext);
}

template <int V> const void* split_kernel() {
    if constexpr (kIsMlp) return nullptr;                     // its step is the network; the cost is nothing
    else return (const void*)fused_command_kernel<TuModel, TuReal, V, false, true>;
}
template <int V> const void* resident_kernel() {
    if constexpr (kIsMlp) return nullptr;
    else return (const void*)resident_command_kernel<TuModel, TuReal, V>;
}
const void* stamped_resident_kernel() {
    if constexpr (MPPI_TU_MODEL == 1 && !MPPI_TU_F64) return (const void*)resident_command_kernel<TuModel, TuReal, V_MPPI, true>;
    else return nullptr;
}
template <int V, int SPLIT, int FAST> const void* tc_kernel() {
    if constexpr (kIsMlp && !MPPI_TU_F64) return (const void*)mlp_tc_command_kernel<V, SPLIT, FAST>;
    else return nullptr;
}

const ModelKernels g_kernels = {
    TuModel::NX, TuModel::NU, MPPI_TU_F64, kIsMlp ? 1 : 0, 0, (int)sizeof(TuParams), tu_load,
    {(const void*)fused_command_kernel<TuModel, TuReal, V_MPPI, false>, (const void*)fused_command_kernel<TuModel, TuReal, V_SMPPI, false>,
     (const void*)fused_command_kernel<TuModel, TuReal, V_KMPPI, false>},
    {split_kernel<V_MPPI>(), split_kernel<V_SMPPI>(), split_kernel<V_KMPPI>()},
    (const void*)fused_command_kernel<TuModel, TuReal, V_MPPI, true>,
    {resident_kernel<V_MPPI>(), resident_kernel<V_SMPPI>(), resident_kernel<V_KMPPI>()},
    stamped_resident_kernel(),
    (const void*)states_kernel<TuModel, TuReal>,
    {{{tc_kernel<V_MPPI, 1, 0>(), tc_kernel<V_MPPI, 1, 1>()}, {tc_kernel<V_MPPI, 0, 0>(), tc_kernel<V_MPPI, 0, 1>()}},
     {{tc_kernel<V_SMPPI, 1, 0>(), tc_kernel<V_SMPPI, 1, 1>()}, {tc_kernel<V_SMPPI, 0, 0>(), tc_kernel<V_SMPPI, 0, 1>()}},
     {{tc_kernel<V_KMPPI, 1, 0>(), tc_kernel<V_KMPPI, 1, 1>()}, {tc_kernel<V_KMPPI, 0, 0>(), tc_kernel<V_KMPPI, 0, 1>()}}},
    nullptr,
};

}  // namespace

namespace mppi_host {
const ModelKernels* MPPI_TU_GETTER() { return &g_kernels; }
}  // namespace mppi_host
