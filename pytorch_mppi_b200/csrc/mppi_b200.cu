// mppi_b200.cu — C-ABI entry points (include/mppi_b200.h) and kernel dispatch.
// Built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC
// No torch types cross this boundary; the Python side binds it with ctypes.
#include <cuda_runtime.h>
#include <new>
#include <stdio.h>
#include <string.h>

#include "../../include/mppi_b200.h"
#include <type_traits>
#include "mppi_fused.cuh"
#include "mppi_mlp_tc.cuh"
#include "mppi_resident.cuh"
#include "mppi_resident_host.h"

using namespace mppi;

namespace {

thread_local char g_cuda_err[512] = "";

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return MPPI_ERR_CUDA;
}
int unsupported_at(const char* why, int line) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (mppi_b200.cu:%d)", why, line);
    return MPPI_ERR_UNSUPPORTED;
}
#define UNSUPPORTED(why) unsupported_at(why, __LINE__)
#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)

struct DevInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
};
int get_dev_info(DevInfo& d) {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo cached;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        CK(cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CK(cudaDeviceGetAttribute(&cached.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cached_dev = dev;
    }
    d = cached;
    return MPPI_OK;
}

int validate(const MppiFusedParams* p, bool fused = false) {
    if (p == nullptr) return MPPI_ERR_BAD_ARG;
    if (p->struct_size != sizeof(MppiFusedParams)) return MPPI_ERR_ABI;
    if (p->K <= 0 || p->T <= 0 || p->nu <= 0 || p->nu > MPPI_MAX_NU || p->nx <= 0) return MPPI_ERR_BAD_ARG;
    if (fused && p->nx > MPPI_MAX_NX) return MPPI_ERR_BAD_ARG;      // state by value; the per-step entry points never touch the state
    if (p->variant < 0 || p->variant > 2) return MPPI_ERR_BAD_ARG;
    if (p->dtype != MPPI_F32 && p->dtype != MPPI_F64) return MPPI_ERR_BAD_ARG;
    if (p->variant == MPPI_VARIANT_KMPPI && (p->S <= 0 || p->W == nullptr || p->theta == nullptr)) return MPPI_ERR_BAD_ARG;
    if (p->variant == MPPI_VARIANT_SMPPI && (p->A == nullptr || p->T < 2)) return MPPI_ERR_BAD_ARG;
    if ((p->flags & MPPI_FLAG_SHIFT) && p->variant == MPPI_VARIANT_KMPPI && p->Wshift == nullptr) return MPPI_ERR_BAD_ARG;
    if (p->lambda_ <= 0.0) return MPPI_ERR_BAD_ARG;
    if (p->world < 0 || p->world > MPPI_MAX_RANKS) return MPPI_ERR_BAD_ARG;
    if (p->u_per_command < 1 || p->u_per_command > p->T) return MPPI_ERR_BAD_ARG;
    return MPPI_OK;
}

inline int rows_of(const MppiFusedParams* p) { return (p->variant == MPPI_VARIANT_KMPPI ? p->S : p->T) * p->nu; }

template <typename real> void fill_noise_model(const MppiFusedParams* p, NoiseModel<real>& nm) {
    for (int i = 0; i < MPPI_MAX_NU; ++i) {
        nm.mu[i] = (real)p->noise_mu[i];
        nm.u_min[i] = (real)p->u_min[i];
        nm.u_max[i] = (real)p->u_max[i];
        nm.a_min[i] = (real)p->action_min[i];
        nm.a_max[i] = (real)p->action_max[i];
    }
    for (int i = 0; i < MPPI_MAX_NU * MPPI_MAX_NU; ++i) {
        nm.L[i] = (real)p->chol[i];
        nm.Sinv[i] = (real)p->sigma_inv[i];
    }
    nm.lambda_ = (real)p->lambda_;
    nm.neg_inv_lambda = (real)(-(1.0 / p->lambda_));   // mppi.py:256: -factor * (cost - beta), factor = 1/lambda
    nm.u_scale = (real)p->u_scale;
    nm.w_smooth = (real)p->w_action_seq_cost;
    nm.delta_t = (real)p->delta_t;
    nm.diag = (p->flags & MPPI_FLAG_DIAG_SIGMA) ? 1 : 0;
    nm.abs_cost = (p->flags & MPPI_FLAG_ABS_COST) ? 1 : 0;
}

inline uint64_t ws_bytes(int nb, int R, int es) {
    return 16 + 2 * (uint64_t)align_up(nb * es, 16) + (uint64_t)align_up(nb * R * es, 16);
}

template <typename real> int fill_kargs(const MppiFusedParams* p, KArgs<real>& a, int BS, int nb, int tps = 1) {
    memset(&a, 0, sizeof(a));
    fill_noise_model<real>(p, a.nm);
    for (int i = 0; i < MPPI_MAX_NU; ++i) a.u_init[i] = (real)p->u_init[i];
    for (int i = 0; i < MPPI_MAX_NX; ++i) a.x0[i] = (real)p->state[i];
    a.state_dev = (p->flags & MPPI_FLAG_STATE_DEVICE) ? (const real*)p->state_dev : nullptr;
    a.state_per_sample = (p->flags & MPPI_FLAG_STATE_PER_SAMPLE) ? 1 : 0;
    a.U = (real*)p->U;
    a.A = (real*)p->A;
    a.theta = (real*)p->theta;
    a.W = (const real*)p->W;
    a.Wshift = (const real*)p->Wshift;
    a.cost_total = (real*)p->cost_total;
    a.action_out = (real*)p->action_out;
    a.nominal_used = (real*)p->nominal_used;
    a.stats = (double*)p->stats;
    a.z = (const real*)p->z;
    a.z_out = (real*)p->z_out;
    a.K = p->K;
    a.T = p->T;
    a.S = p->S;
    a.R = rows_of(p);
    a.TN = p->T * p->nu;
    a.upc = p->u_per_command;
    a.n_tiles = (p->K + BS - 1) / BS;
    a.tps = tps;
    a.k_offset = p->k_offset;
    a.seed = p->seed;
    a.offset = p->offset;
    a.shift = (p->flags & MPPI_FLAG_SHIFT) ? 1 : 0;
    a.null_action = (p->flags & MPPI_FLAG_NULL_ACTION) ? 1 : 0;
    a.pdl = (p->flags & MPPI_FLAG_PDL) ? 1 : 0;
    const int es = (int)sizeof(real);
    const bool padded = (p->flags & MPPI_FLAG_NOMINAL_PADDED) || ((a.TN * es) % 16 == 0);
    a.tma_ok = padded && ((uintptr_t)p->U % 16 == 0) && (p->variant != MPPI_VARIANT_SMPPI || (uintptr_t)p->A % 16 == 0);
    // workspace carve
    if (p->workspace != nullptr) {
        unsigned char* w = (unsigned char*)p->workspace;
        a.ticket = (unsigned int*)w;
        a.betaP = (real*)(w + 16);
        a.etaP = (real*)(w + 16 + align_up(nb * es, 16));
        a.VP = (real*)(w + 16 + 2 * align_up(nb * es, 16));
    }
    a.rank = p->rank;
    a.world = p->world <= 0 ? 1 : p->world;
    a.epoch = p->epoch;
    a.export_partial = (p->flags & MPPI_FLAG_EXPORT_PARTIAL) ? 1 : 0;
    a.partial_out = (double*)p->partial_out;
    a.torch_total = p->torch_rng_total;
    a.offset_dev = (unsigned long long*)p->offset_dev;
    a.offset_inc = p->offset_inc;
    a.n_env = p->n_env > 1 ? p->n_env : 1;
    a.env_u_stride = p->env_u_stride;
    a.env_ws_stride = (long long)p->env_ws_stride;
    a.dbg = (unsigned long long*)p->debug_clocks;
    a.host_mailbox = (unsigned long long*)p->host_mailbox;
    a.host_epoch = p->host_epoch;
    bool any_peer = false;
    for (int g = 0; g < MPPI_MAX_RANKS; ++g) {
        a.peers[g] = (unsigned long long*)p->peer_slots[g];
        any_peer = any_peer || p->peer_slots[g] != nullptr;
    }
    if (!any_peer || a.export_partial) a.world = a.export_partial ? a.world : 1;
    return MPPI_OK;
}

inline cudaError_t launch_raw(const void* kernel, int nb, int BD, int smem, cudaStream_t stream, void** argv, bool pdl, int ny = 1) {
    if (!pdl) return cudaLaunchKernel(kernel, dim3(nb, ny), dim3(BD), argv, (size_t)smem, stream);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nb, ny);
    cfg.blockDim = dim3(BD);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, kernel, argv);
}

template <typename... Args>
int launch_kernel(void (*kernel)(Args...), int nb, int BD, int smem, cudaStream_t stream, int ny, Args... args) {
    void* argv[] = {(void*)&args...};
    cudaError_t e = cudaLaunchKernel((const void*)kernel, dim3(nb, ny), dim3(BD), argv, (size_t)smem, stream);
    if (e != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "launch grid=%d block=%d smem=%d: %s (%s)", nb, BD, smem,
                 cudaGetErrorName(e), cudaGetErrorString(e));
        return MPPI_ERR_CUDA;
    }
    return MPPI_OK;
}

struct Geometry {
    int BD, BS, tps, nb, smem, occ, regs;   // BD = BS * tps threads per CTA, BS samples per tile
};

struct GeomKey {
    const void* kernel;
    int dev, variant, K, T, nu, S, bt, tp, gb, r2, single, ne;
    bool operator==(const GeomKey& o) const {
        return kernel == o.kernel && dev == o.dev && variant == o.variant && K == o.K && T == o.T && nu == o.nu && S == o.S &&
               bt == o.bt && tp == o.tp && gb == o.gb && r2 == o.r2 && single == o.single && ne == o.ne;
    }
};

// Launch geometry for (kernel, dimensions).  The occupancy / attribute queries cost microseconds, so
// the last few results are cached per thread: a steady-state command() pays only the lookup.
static thread_local int g_tc_kernel = 0;   // set around plan_geometry() for the tcgen05 kernels (see below)
template <typename KernelT>
int plan_geometry(KernelT kernel, const MppiFusedParams* p, int es, int need_rows2, bool single_partial_grid, Geometry& g,
                  SmemLayout (*layout)(int, int, int, int, int, int, int, int, int)) {
    static thread_local GeomKey keys[8];
    static thread_local Geometry vals[8];
    static thread_local int n_cached = 0, next_slot = 0;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    const GeomKey key{(const void*)kernel, dev, p->variant, p->K, p->T, p->nu, p->S, p->block_threads, p->threads_per_sample,
                      p->grid_blocks, need_rows2, single_partial_grid ? 1 : 0, p->n_env > 1 ? p->n_env : 1};
    for (int i = 0; i < n_cached; ++i)
        if (keys[i] == key) {
            g = vals[i];
            return MPPI_OK;
        }
    DevInfo di;
    int rc = get_dev_info(di);
    if (rc) return rc;
    const int R = rows_of(p);
    // BS samples per tile; tps threads share one sample's sampling/transform work.
    //
    // Automatic geometry (block_threads == 0): blocks are statically assigned tiles (determinism:
    // the reduction order must not depend on scheduling), so the finish time follows the most
    // loaded SM.  Enumerate BS in steps of a warp and j = resident CTAs per SM, size the grid as
    // min(n_tiles, SMs*j), and keep the candidate with the smallest worst-case samples per SM
    // (ties: fewer passes, then more threads).  Measured on B200 (scripts/geom_sweep.py) this picks
    // the winners of an exhaustive sweep within ~3 %: e.g. K=131072 -> BS=448, 293 CTAs, one pass.
    int BS = p->block_threads;
    int tps = p->threads_per_sample;
    int grid_hint = 0;
    if (BS <= 0) {
        cudaFuncAttributes fa0;
        CK(cudaFuncGetAttributes(&fa0, kernel));
        const int dyn0 = di.max_smem_optin - (int)fa0.sharedSizeBytes;
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn0));
        long long best_load = -1;
        int best_bs = 128, best_passes = 0, best_grid = 0;
        // latency hiding needs ~24 resident warps per SM when the problem is large enough to supply them
        const long long per_sm = ((long long)p->K + di.sm_count - 1) / di.sm_count;
        const long long want_threads = per_sm < 768 ? per_sm : 768;
        for (int bs = 128; bs <= 512; bs += 32) {
            SmemLayout Lc = layout(p->variant, p->T, p->nu, p->S, R, bs, bs, single_partial_grid ? 1 : di.sm_count * 4, need_rows2);
            if (Lc.total > dyn0) continue;
            int occ_c = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, kernel, bs, Lc.total));
            if (occ_c < 1) continue;
            const int nt = (p->K + bs - 1) / bs;
            for (int j = 1; j <= occ_c && j <= 8; ++j) {
                const int nbc = nt < di.sm_count * j ? nt : di.sm_count * j;
                const int passes = (nt + nbc - 1) / nbc;
                const int bps = (nbc + di.sm_count - 1) / di.sm_count;
                long long load = (long long)bps * passes * bs;
                const long long resident = (long long)bps * bs;
                if (resident < want_threads) load = load * want_threads / resident;   // under-occupied: proportionally slower
                const bool better = best_load < 0 || load < best_load ||
                                    (load == best_load && (passes < best_passes || (passes == best_passes && bs > best_bs)));
                if (better) {
                    best_load = load;
                    best_bs = bs;
                    best_passes = passes;
                    best_grid = nbc;
                }
            }
        }
        BS = best_bs;
        grid_hint = best_grid;
    }
    if (BS % 32 != 0 || BS < 32 || BS > 512) return MPPI_ERR_BAD_ARG;
    const int n_tiles = (p->K + BS - 1) / BS;
    if (tps <= 0) {
        // helper threads only pay off while an SM hosts a single small CTA
        tps = 1;
        const int envs = p->n_env > 1 ? p->n_env : 1;
        if ((long long)n_tiles * envs <= di.sm_count) tps = 512 / BS >= 4 ? 4 : (512 / BS >= 2 ? 2 : 1);
    }
    while (tps > 1 && BS * tps > 512) tps >>= 1;
    if (tps != 1 && tps != 2 && tps != 4) return MPPI_ERR_BAD_ARG;
    const int BD = BS * tps;
    const int cap = di.sm_count * 16;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kernel));
    const int dyn_limit = di.max_smem_optin - (int)fa.sharedSizeBytes;   // static + dynamic <= opt-in maximum
    // The attribute is a per-kernel LIMIT (setting a smaller value later lowers it), so raise it
    // once to the device maximum; the carve-out actually used follows each launch's request.
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_limit));
    // The layout depends on the grid (rescale factors of nb partials live in shared memory) and the
    // grid on the occupancy the layout allows: iterate from an optimistic guess to a fixed point.
    int nb = n_tiles < cap ? n_tiles : cap;
    int occ = 0;
    SmemLayout L;
    for (int it = 0; it < 4; ++it) {
        L = layout(p->variant, p->T, p->nu, p->S, R, BD, BS, single_partial_grid ? 1 : nb, need_rows2);
        if (L.total > dyn_limit) return UNSUPPORTED("shared-memory tile does not fit");
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, BD, L.total));
        if (occ < 1) return UNSUPPORTED("kernel does not fit on an SM with this block size");
        if (g_tc_kernel) {
            // The occupancy API answers 1 CTA/SM for kernels that allocate tensor memory; measured on B200 the
            // 128-thread tcgen05 CTAs do co-reside (K=131072, T=30: 585 us at 1 CTA/SM, 379 at 2, 311 at 3), so
            // size the grid from the real limits: shared memory, registers, and 64 of 512 TMEM columns per CTA.
            int smem_sm = 0;
            CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
            const int by_smem = smem_sm / ((int)fa.sharedSizeBytes + L.total + 1024);
            const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * BD);
            int o = by_smem < by_regs ? by_smem : by_regs;
            if (o > 512 / 64) o = 512 / 64;
            const char* e = getenv("MPPI_TC_OCC");
            if (e != nullptr && atoi(e) > 0) o = atoi(e);
            if (o > occ) occ = o;
        }
        int nb2 = n_tiles < di.sm_count * occ ? n_tiles : di.sm_count * occ;
        if (nb2 > cap) nb2 = cap;
        if (p->grid_blocks > 0 && p->grid_blocks < nb2) nb2 = p->grid_blocks;
        if (grid_hint > 0 && grid_hint < nb2) nb2 = grid_hint;
        if (nb2 == nb) break;
        nb = nb2;
    }
    L = layout(p->variant, p->T, p->nu, p->S, R, BD, BS, single_partial_grid ? 1 : nb, need_rows2);
    g.BS = BS;
    g.tps = tps;
    g.BD = BD;
    g.nb = nb;
    g.smem = L.total;
    g.occ = occ;
    g.regs = fa.numRegs;
    (void)es;
    keys[next_slot] = key;
    vals[next_slot] = g;
    next_slot = (next_slot + 1) % 8;
    if (n_cached < 8) ++n_cached;
    return MPPI_OK;
}

template <typename real> SmemLayout layout_fn(int v, int T, int nu, int S, int R, int BD, int BS, int nb, int r2) {
    return make_layout<real>(v, T, nu, S, R, BD, BS, nb, r2);
}

// ---- fused command ----------------------------------------------------------------------------
// PENDULUM_MLP in fp32 with model_params[3] != 0: the tcgen05/TMEM kernel (mppi_mlp_tc.cuh).  One CTA = 128
// threads = 128 samples = the 128 lanes of an M=128 accumulator tile; it does not take part in PDL.
template <class Model, typename real, int V, typename KernelT>
bool select_tensor_core_route(const MppiFusedParams* p, MppiFusedParams& p_tc, KernelT& kernel) {
    if constexpr (std::is_same<Model, PendulumMLPModel>::value && std::is_same<real, float>::value) {
        const int mode = (int)p->model_params[3];      // 1: hi/lo-split bf16 operands ("3 x bf16"), 2: plain bf16
        if ((mode == 1 || mode == 2) && p->n_env <= 1) {
            p_tc = *p;
            // 128-thread CTAs, tiles of 128 samples (every thread rolls one).  The kernel also supports tiles of 64
            // samples + 64 helper threads (MPPI_TC_TILE=64; twice the CTAs to spread over the SMs), but a CTA's step
            // time is set by the three MMA round trips, not by its tanh work: measured K=32768, T=30: 132 us with
            // half tiles against 113 us with full ones, so full tiles are the default at every K.
            int bs = 128;
            const char* e_bs = getenv("MPPI_TC_TILE");
            if (e_bs != nullptr && (atoi(e_bs) == 64 || atoi(e_bs) == 128)) bs = atoi(e_bs);
            p_tc.block_threads = bs;
            p_tc.threads_per_sample = 128 / bs;
            p_tc.flags &= ~(uint32_t)MPPI_FLAG_PDL;
            const bool fast = p->model_params[2] != 0.0;
            kernel = mode == 1 ? (fast ? mlp_tc_command_kernel<V, 1, 1> : mlp_tc_command_kernel<V, 1, 0>)
                               : (fast ? mlp_tc_command_kernel<V, 0, 1> : mlp_tc_command_kernel<V, 0, 0>);
            g_tc_kernel = 1;
            // co-residency is bounded by shared memory: ask for the largest carve-out
            cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            return true;
        }
    }
    (void)p_tc;
    (void)kernel;
    return false;
}

// MPPI_FLAG_SPLIT_COST: problems small enough to run with helper threads (threads_per_sample > 1) may take the
// split-cost rollout (fused_command_kernel<..., SPLIT = true>) if its per-step state buffer fits in shared memory;
// the geometry stays the one chosen for the plain kernel.
template <class Model, typename real, int V, typename KernelT>
void select_split_cost_rollout(const MppiFusedParams* p, bool eligible, KernelT& kernel, Geometry& g, int& split) {
    split = 0;
    if constexpr (!std::is_same<Model, PendulumMLPModel>::value) {      // its step is the network; the cost is nothing
        if (!eligible || !(p->flags & MPPI_FLAG_SPLIT_COST) || g.tps <= 1) return;
        KernelT k2 = fused_command_kernel<Model, real, V, false, true>;
        MppiFusedParams p2 = *p;
        p2.block_threads = g.BS;
        p2.threads_per_sample = g.tps;
        p2.grid_blocks = g.nb;
        Geometry g2;
        if (plan_geometry(k2, &p2, (int)sizeof(real), Model::NX << 8, false, g2, layout_fn<real>) != MPPI_OK) return;
        if (g2.BS != g.BS || g2.tps != g.tps) return;
        kernel = k2;
        g = g2;
        split = 1;
    }
}

// MPPI_FLAG_WIDE_REGS: a launch that puts at most one CTA on an SM can afford the instantiation compiled without the
// 64-register cap (fused_command_kernel<..., SPLIT = false, MINB = 1>: no spills in the last-CTA tail).
template <class Model, typename real, int V, typename KernelT>
void select_wide_register_kernel(const MppiFusedParams* p, bool eligible, KernelT& kernel, Geometry& g, int& wide) {
    wide = 0;
    if (!eligible || !(p->flags & MPPI_FLAG_WIDE_REGS)) return;
    DevInfo di;
    if (get_dev_info(di) != MPPI_OK || g.nb > di.sm_count) return;
    KernelT k2 = fused_command_kernel<Model, real, V, false, false, 1>;
    MppiFusedParams p2 = *p;
    p2.block_threads = g.BS;
    p2.threads_per_sample = g.tps;
    p2.grid_blocks = g.nb;
    Geometry g2;
    if (plan_geometry(k2, &p2, (int)sizeof(real), 0, false, g2, layout_fn<real>) != MPPI_OK) return;
    if (g2.BS != g.BS || g2.tps != g.tps || g2.nb != g.nb) return;
    kernel = k2;
    g = g2;
    wide = 1;
}

template <class Model, typename real, int V>
int run_fused(const MppiFusedParams* p, cudaStream_t stream, MppiLaunchInfo* info) {
    if (p->nx != Model::NX || p->nu != Model::NU) return MPPI_ERR_BAD_ARG;
    const bool batched = p->n_env > 1;
    if (batched && (V != V_MPPI || p->world > 1)) return UNSUPPORTED("batched environments: MPPI variant, single GPU only");
    if (batched && info == nullptr && !(p->flags & MPPI_FLAG_STATE_DEVICE)) return MPPI_ERR_BAD_ARG;   // states are (n_env, nx) on the device
    auto kernel = (batched && V == V_MPPI) ? fused_command_kernel<Model, real, V_MPPI, true> : fused_command_kernel<Model, real, V, false>;
    MppiFusedParams p_tc;
    const bool tc_route = select_tensor_core_route<Model, real, V>(p, p_tc, kernel);
    if (tc_route) p = &p_tc;
    Geometry g;
    int rc = plan_geometry(kernel, p, (int)sizeof(real), 0, false, g, layout_fn<real>);
    g_tc_kernel = 0;
    if (rc) return rc;
    int split = 0, wide = 0;
    select_split_cost_rollout<Model, real, V>(p, !tc_route && !batched, kernel, g, split);
    select_wide_register_kernel<Model, real, V>(p, !tc_route && !batched && !split, kernel, g, wide);
    const uint64_t need_ws = ws_bytes(g.nb, rows_of(p), (int)sizeof(real));
    KArgs<real> a;
    fill_kargs<real>(p, a, g.BS, g.nb, g.tps);
    if (info != nullptr) {
        DevInfo di;
        get_dev_info(di);
        info->block_threads = g.BD;
        info->threads_per_sample = g.tps;
        info->grid_blocks = g.nb;
        info->smem_bytes = g.smem;
        info->regs_per_thread = g.regs;
        info->max_blocks_per_sm = g.occ;
        info->sm_count = di.sm_count;
        // report the worst case so one allocation serves any later geometry for these dimensions
        info->workspace_bytes = ws_bytes(di.sm_count * 16, rows_of(p), (int)sizeof(real));
        info->tma_staging = a.tma_ok;
        info->split_cost = split;
        info->wide_regs = wide;
        return MPPI_OK;
    }
    if (p->U == nullptr || p->cost_total == nullptr || p->action_out == nullptr || p->nominal_used == nullptr ||
        p->stats == nullptr || p->workspace == nullptr)
        return MPPI_ERR_BAD_ARG;
    if (p->workspace_bytes < need_ws) return MPPI_ERR_WORKSPACE;
    if (batched && (p->env_ws_stride < need_ws || p->workspace_bytes < p->env_ws_stride * (uint64_t)p->n_env ||
                    p->env_u_stride < p->T * p->nu))
        return MPPI_ERR_WORKSPACE;
    if (a.world > 1 && rows_of(p) > MPPI_XCHG_MAX_R) return UNSUPPORTED("T*nu exceeds the peer mailbox record size");
    if (a.export_partial && p->partial_out == nullptr) return MPPI_ERR_BAD_ARG;
    typename Model::template P<real> mp;
    Model::template load<real>(mp, p->model_params, p->model_params_ext, p->n_model_params_ext);
    void* argv2[2] = {(void*)&a, (void*)&mp};
    cudaError_t e = launch_raw((const void*)kernel, g.nb, g.BD, g.smem, stream, argv2, a.pdl != 0, a.n_env);
    if (e != cudaSuccess) return cuda_fail(e, "fused launch");
    return MPPI_OK;
}

template <class Model, typename real>
int run_fused_variant(const MppiFusedParams* p, cudaStream_t s, MppiLaunchInfo* info) {
    switch (p->variant) {
        case MPPI_VARIANT_MPPI: return run_fused<Model, real, V_MPPI>(p, s, info);
        case MPPI_VARIANT_SMPPI: return run_fused<Model, real, V_SMPPI>(p, s, info);
        case MPPI_VARIANT_KMPPI: return run_fused<Model, real, V_KMPPI>(p, s, info);
    }
    return MPPI_ERR_BAD_ARG;
}

template <class Model> int run_fused_dtype(const MppiFusedParams* p, cudaStream_t s, MppiLaunchInfo* info) {
    return p->dtype == MPPI_F32 ? run_fused_variant<Model, float>(p, s, info) : run_fused_variant<Model, double>(p, s, info);
}

// ---- plans ---------------------------------------------------------------------------------------
// Resident mode: the host side of the protocol (struct Resident, res_*) is csrc/mppi_resident_host.h; this file supplies its
// backend (cooperative launch / stream synchronise / stream query).
struct ResidentDevice {
    unsigned long long* host_box = nullptr;
    unsigned long long* board = nullptr;
    void* action_dev = nullptr;
    cudaStream_t stream = nullptr;
    unsigned long long idle_ns = 0;
};

struct Plan {
    MppiFusedParams p;
    const void* kernel;
    const void* res_kernel;            // resident_command_kernel<Model, real, V, sharded>, or nullptr when this plan cannot run resident
    int res_xchg;                      // the plan is one shard of a multi-GPU controller: records carry the exchange epoch
    Resident res;
    ResidentDevice resdev;
    Geometry g;
    int is_double, nx, upc_nu, pdl;
    unsigned long long epoch, host_epoch;
    alignas(16) unsigned char kargs[sizeof(KArgs<double>)];
    alignas(16) unsigned char mparams[12288];
};

template <class Model, typename real, int V> int build_plan(const MppiFusedParams* p, Plan* pl) {
    if (p->nx != Model::NX || p->nu != Model::NU) return MPPI_ERR_BAD_ARG;
    const bool batched = p->n_env > 1;
    if (batched && (V != V_MPPI || p->world > 1)) return UNSUPPORTED("batched environments: MPPI variant, single GPU only");
    auto kernel = (batched && V == V_MPPI) ? fused_command_kernel<Model, real, V_MPPI, true> : fused_command_kernel<Model, real, V, false>;
    MppiFusedParams p_tc;
    const bool tc_route = select_tensor_core_route<Model, real, V>(p, p_tc, kernel);
    if (tc_route) p = &p_tc;
    int rc = plan_geometry(kernel, p, (int)sizeof(real), 0, false, pl->g, layout_fn<real>);
    g_tc_kernel = 0;
    if (rc) return rc;
    int split = 0, wide = 0;
    select_split_cost_rollout<Model, real, V>(p, !tc_route && !batched, kernel, pl->g, split);
    select_wide_register_kernel<Model, real, V>(p, !tc_route && !batched && !split, kernel, pl->g, wide);
    if (p->U == nullptr || p->cost_total == nullptr || p->nominal_used == nullptr || p->stats == nullptr || p->workspace == nullptr)
        return MPPI_ERR_BAD_ARG;
    if (p->workspace_bytes < ws_bytes(pl->g.nb, rows_of(p), (int)sizeof(real))) return MPPI_ERR_WORKSPACE;
    static_assert(sizeof(typename Model::template P<real>) <= sizeof(pl->mparams), "model parameter block too large");
    KArgs<real>* a = reinterpret_cast<KArgs<real>*>(pl->kargs);
    fill_kargs<real>(p, *a, pl->g.BS, pl->g.nb, pl->g.tps);
    if (a->world > 1 && rows_of(p) > MPPI_XCHG_MAX_R) return UNSUPPORTED("T*nu exceeds the peer mailbox record size");
    typename Model::template P<real>* mp = reinterpret_cast<typename Model::template P<real>*>(pl->mparams);
    Model::template load<real>(*mp, p->model_params, p->model_params_ext, p->n_model_params_ext);
    pl->kernel = (const void*)kernel;
    pl->res_kernel = nullptr;
    pl->res_xchg = 0;
    if constexpr (!std::is_same<Model, PendulumMLPModel>::value) {
        // resident mode runs the split-cost rollout with one tile per CTA: exactly the plans that took it; a sharded
        // controller (in-kernel NVLink exchange) gets the instantiation whose records carry the exchange epoch
        if (split && !batched && !a->export_partial && pl->g.nb == a->n_tiles)
            pl->res_kernel = a->world > 1 ? (const void*)resident_command_kernel<Model, real, V, true>
                                          : (const void*)resident_command_kernel<Model, real, V, false>;
        pl->res_xchg = a->world > 1 ? 1 : 0;
        // the one instantiation with %globaltimer stamps (a profiling aid, scripts/resident_timeline.py)
        if constexpr (std::is_same<Model, PendulumModel>::value && std::is_same<real, float>::value && V == V_MPPI) {
            if (pl->res_kernel != nullptr && !pl->res_xchg && p->debug_clocks != nullptr)
                pl->res_kernel = (const void*)resident_command_kernel<Model, real, V, false, true>;
        }
    }
    pl->is_double = sizeof(real) == 8;
    pl->nx = Model::NX;
    pl->upc_nu = p->u_per_command * p->nu;
    pl->pdl = (p->flags & MPPI_FLAG_PDL) ? 1 : 0;
    pl->epoch = p->epoch;
    pl->host_epoch = 0;
    pl->p = *p;
    return MPPI_OK;
}

template <class Model, typename real> int build_plan_variant(const MppiFusedParams* p, Plan* pl) {
    switch (p->variant) {
        case MPPI_VARIANT_MPPI: return build_plan<Model, real, V_MPPI>(p, pl);
        case MPPI_VARIANT_SMPPI: return build_plan<Model, real, V_SMPPI>(p, pl);
        case MPPI_VARIANT_KMPPI: return build_plan<Model, real, V_KMPPI>(p, pl);
    }
    return MPPI_ERR_BAD_ARG;
}
template <class Model> int build_plan_dtype(const MppiFusedParams* p, Plan* pl) {
    return p->dtype == MPPI_F32 ? build_plan_variant<Model, float>(p, pl) : build_plan_variant<Model, double>(p, pl);
}

template <typename real>
inline void plan_update(Plan* pl, const double* state, const void* state_dev, uint32_t flags, uint64_t seed, uint64_t offset,
                        const void* z, void* action_out, void* host_mailbox) {
    KArgs<real>* a = reinterpret_cast<KArgs<real>*>(pl->kargs);
    if (state != nullptr)
        for (int i = 0; i < pl->nx; ++i) a->x0[i] = (real)state[i];
    a->state_dev = (flags & MPPI_FLAG_STATE_DEVICE) ? (const real*)state_dev : nullptr;
    a->state_per_sample = (flags & MPPI_FLAG_STATE_PER_SAMPLE) ? 1 : 0;
    a->shift = (flags & MPPI_FLAG_SHIFT) ? 1 : 0;
    a->seed = seed;
    a->offset = offset;
    a->z = (const real*)z;
    a->action_out = (real*)action_out;
    if (a->world > 1 || a->export_partial) a->epoch = ++pl->epoch;
    a->host_mailbox = (unsigned long long*)host_mailbox;
    if (host_mailbox != nullptr) a->host_epoch = ++pl->host_epoch;
}

inline int plan_launch(Plan* pl, cudaStream_t stream) {
    void* argv[2] = {(void*)pl->kargs, (void*)pl->mparams};
    cudaError_t e = launch_raw(pl->kernel, pl->g.nb, pl->g.BD, pl->g.smem, stream, argv, pl->pdl != 0, pl->p.n_env > 1 ? pl->p.n_env : 1);
    if (e != cudaSuccess) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "plan launch grid=%d block=%d smem=%d: %s (%s)", pl->g.nb, pl->g.BD, pl->g.smem,
                 cudaGetErrorName(e), cudaGetErrorString(e));
        return MPPI_ERR_CUDA;
    }
    return MPPI_OK;
}

// ---- resident mode: the device backend of csrc/mppi_resident_host.h ---------------------------------
template <typename real>
int resident_launch_t(Plan* pl, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq_start, uint32_t gen) {
    ResidentDevice& d = pl->resdev;
    KArgs<real> a = *reinterpret_cast<KArgs<real>*>(pl->kargs);
    a.seed = seed;
    a.z = nullptr;
    a.z_out = nullptr;
    a.state_dev = nullptr;
    a.action_out = (real*)d.action_dev;
    a.host_mailbox = d.host_box + RES_BOX_ACTION;
    a.pdl = 0;
    ResidentArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.host_cmd = d.host_box + RES_BOX_RECORD;
    ra.host_status = d.host_box + RES_BOX_DONE;
    ra.board = d.board;
    ra.seq_start = seq_start;
    ra.offset_pred = offset_pred;
    ra.idle_ns = d.idle_ns;
    ra.gen = gen;
    ra.shift_pred = shift_pred;
    ra.n_words = 3 + pl->nx * (pl->is_double ? 2 : 1) + (pl->res_xchg ? 2 : 0);
    DevInfo di;
    int rc = get_dev_info(di);
    if (rc) return rc;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, pl->res_kernel));
    const int dyn_limit = di.max_smem_optin - (int)fa.sharedSizeBytes;
    if (pl->g.smem > dyn_limit) return UNSUPPORTED("resident kernel: shared-memory tile does not fit");
    CK(cudaFuncSetAttribute(pl->res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_limit));
    CK(cudaMemsetAsync(d.board, 0, MPPI_RES_BOARD_WORDS * sizeof(unsigned long long), d.stream));
    void* argv[3] = {(void*)&a, (void*)pl->mparams, (void*)&ra};
    // cooperative: every CTA is resident or the launch fails — the CTAs wait for each other through the board
    cudaError_t e = cudaLaunchCooperativeKernel(pl->res_kernel, dim3(pl->g.nb), dim3(pl->g.BD), argv, (size_t)pl->g.smem, d.stream);
    if (e != cudaSuccess) return cuda_fail(e, "resident launch");
    return MPPI_OK;
}
int resident_be_launch(void* ctx, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq_start, uint32_t gen) {
    Plan* pl = reinterpret_cast<Plan*>(ctx);
    return pl->is_double ? resident_launch_t<double>(pl, seed, offset_pred, shift_pred, seq_start, gen)
                         : resident_launch_t<float>(pl, seed, offset_pred, shift_pred, seq_start, gen);
}
int resident_be_drain(void* ctx) {
    Plan* pl = reinterpret_cast<Plan*>(ctx);
    CK(cudaStreamSynchronize(pl->resdev.stream));
    return MPPI_OK;
}
int resident_be_health(void* ctx) {
    Plan* pl = reinterpret_cast<Plan*>(ctx);
    cudaError_t q = cudaStreamQuery(pl->resdev.stream);
    if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery while waiting for the resident grid");
    return MPPI_OK;
}
static_assert(MPPI_RES_CMD_SHIFT == (unsigned)RES_CMD_SHIFT && MPPI_RES_CMD_STOP == (unsigned)RES_CMD_STOP, "record flags: host and device disagree");
static_assert((int)RES_ERR_BAD_ARG == (int)MPPI_ERR_BAD_ARG && (int)RES_ERR_TIMEOUT == (int)MPPI_ERR_TIMEOUT, "status codes of the protocol header");

int dispatch_fused(const MppiFusedParams* p, cudaStream_t s, MppiLaunchInfo* info) {
    int rc = validate(p, true);
    if (rc) return rc;
    switch (p->model) {
#ifndef MPPI_ONLY_USER_MODEL
        case MPPI_MODEL_PENDULUM: return run_fused_dtype<PendulumModel>(p, s, info);
        case MPPI_MODEL_LINEAR_POINT: return run_fused_dtype<LinearPointModel>(p, s, info);
        case MPPI_MODEL_PENDULUM_MLP:
            if (p->n_model_params_ext < PendulumMLPModel::N_EXT || p->model_params_ext == nullptr) return MPPI_ERR_BAD_ARG;
            return run_fused_dtype<PendulumMLPModel>(p, s, info);
#endif
#ifdef MPPI_USER_MODEL_HEADER
        case MPPI_MODEL_USER: return run_fused_dtype<UserModel>(p, s, info);
#endif
    }
    return MPPI_ERR_UNSUPPORTED;
}

// ---- generic path: sample / softmin -------------------------------------------------------------
template <typename real, int V, int NU>
int run_sample(const MppiFusedParams* p, KArgs<real>& a_extra, cudaStream_t stream) {
    auto kernel = sample_kernel<real, V, NU>;
    Geometry g;
    int rc = plan_geometry(kernel, p, (int)sizeof(real), V == V_KMPPI, true, g, layout_fn<real>);
    if (rc) return rc;
    KArgs<real> a;
    fill_kargs<real>(p, a, g.BS, g.nb, g.tps);
    a.out_pa = a_extra.out_pa;
    a.out_noise = a_extra.out_noise;
    a.out_noise_theta = a_extra.out_noise_theta;
    a.out_cost_init = a_extra.out_cost_init;
    a.override_rows = a_extra.override_rows;
    a.n_override = a_extra.n_override;
    a.override_start = a_extra.override_start;
    if (a_extra.shift < 0) {   // materialise from nominal_used: no shift, plain loads
        a.U = (real*)p->nominal_used;
        a.A = a.U + a.TN;
        a.theta = a.U + 2 * a.TN;
        a.shift = 0;
        a.tma_ok = 0;
        a.nominal_used = nullptr;
    }
    return launch_kernel(kernel, g.nb, g.BD, g.smem, stream, a.n_env, a);
}

template <typename real, int V> int run_sample_nu(const MppiFusedParams* p, KArgs<real>& e, cudaStream_t s) {
    switch (p->nu) {
        case 1: return run_sample<real, V, 1>(p, e, s);
        case 2: return run_sample<real, V, 2>(p, e, s);
        case 3: return run_sample<real, V, 3>(p, e, s);
        case 4: return run_sample<real, V, 4>(p, e, s);
    }
    return MPPI_ERR_UNSUPPORTED;
}

template <typename real>
int run_sample_any(const MppiFusedParams* p, void* pa, void* noise, void* noise_theta, void* cost_init, const void* ovr,
                   int n_ovr, int ovr_start, bool from_nominal_used, cudaStream_t s) {
    KArgs<real> e;
    memset(&e, 0, sizeof(e));
    e.out_pa = (real*)pa;
    e.out_noise = (real*)noise;
    e.out_noise_theta = (real*)noise_theta;
    e.out_cost_init = (real*)cost_init;
    e.override_rows = (const real*)ovr;
    e.n_override = n_ovr;
    e.override_start = ovr_start;
    e.shift = from_nominal_used ? -1 : 0;
    switch (p->variant) {
        case MPPI_VARIANT_MPPI: return run_sample_nu<real, V_MPPI>(p, e, s);
        case MPPI_VARIANT_SMPPI: return run_sample_nu<real, V_SMPPI>(p, e, s);
        case MPPI_VARIANT_KMPPI: return run_sample_nu<real, V_KMPPI>(p, e, s);
    }
    return MPPI_ERR_BAD_ARG;
}

template <typename real, int V, int NU>
int run_softmin(const MppiFusedParams* p, const void* cost, const void* eps, cudaStream_t stream) {
    auto kernel = softmin_update_kernel<real, V, NU>;
    Geometry g;
    int rc = plan_geometry(kernel, p, (int)sizeof(real), 0, false, g, layout_fn<real>);
    if (rc) return rc;
    const uint64_t need_ws = ws_bytes(g.nb, rows_of(p), (int)sizeof(real));
    if (p->workspace == nullptr || p->workspace_bytes < need_ws) return MPPI_ERR_WORKSPACE;
    if (p->n_env > 1 && (p->env_ws_stride < need_ws || p->workspace_bytes < p->env_ws_stride * (uint64_t)p->n_env)) return MPPI_ERR_WORKSPACE;
    KArgs<real> a;
    fill_kargs<real>(p, a, g.BS, g.nb, g.tps);
    a.in_cost = (const real*)cost;
    a.in_eps = (const real*)eps;
    if (a.world > 1 && rows_of(p) > MPPI_XCHG_MAX_R) return UNSUPPORTED("T*nu exceeds the peer mailbox record size");
    return launch_kernel(kernel, g.nb, g.BD, g.smem, stream, a.n_env, a);
}

template <typename real, int V> int run_softmin_nu(const MppiFusedParams* p, const void* c, const void* e, cudaStream_t s) {
    switch (p->nu) {
        case 1: return run_softmin<real, V, 1>(p, c, e, s);
        case 2: return run_softmin<real, V, 2>(p, c, e, s);
        case 3: return run_softmin<real, V, 3>(p, c, e, s);
        case 4: return run_softmin<real, V, 4>(p, c, e, s);
    }
    return MPPI_ERR_UNSUPPORTED;
}
template <typename real> int run_softmin_any(const MppiFusedParams* p, const void* c, const void* e, cudaStream_t s) {
    switch (p->variant) {
        case MPPI_VARIANT_MPPI: return run_softmin_nu<real, V_MPPI>(p, c, e, s);
        case MPPI_VARIANT_SMPPI: return run_softmin_nu<real, V_SMPPI>(p, c, e, s);
        case MPPI_VARIANT_KMPPI: return run_softmin_nu<real, V_KMPPI>(p, c, e, s);
    }
    return MPPI_ERR_BAD_ARG;
}

template <class Model, typename real>
int run_states(const MppiFusedParams* p, const void* pa, void* states, cudaStream_t stream) {
    if (p->nx != Model::NX || p->nu != Model::NU) return MPPI_ERR_BAD_ARG;
    KArgs<real> a;
    fill_kargs<real>(p, a, 128, 1);
    typename Model::template P<real> mp;
    Model::template load<real>(mp, p->model_params, p->model_params_ext, p->n_model_params_ext);
    states_kernel<Model, real><<<(p->K + 127) / 128, 128, 0, stream>>>((const real*)pa, (real*)states, a, mp,
                                                                        (long long)p->T * p->nu);
    CK(cudaGetLastError());
    return MPPI_OK;
}

// get_rollouts (mppi.py:425-448): n start states, each rolled through an action sequence
template <class Model, typename real>
int run_rollout_states(const MppiFusedParams* p, const void* start_states, const void* actions, long long stride, int n, int T,
                       void* states, cudaStream_t stream) {
    if (p->nx != Model::NX || p->nu != Model::NU) return MPPI_ERR_BAD_ARG;
    KArgs<real> a;
    memset(&a, 0, sizeof(a));
    a.nm.u_scale = (real)p->u_scale;
    a.K = n;
    a.T = T;
    a.state_dev = (const real*)start_states;
    a.state_per_sample = 1;
    typename Model::template P<real> mp;
    Model::template load<real>(mp, p->model_params, p->model_params_ext, p->n_model_params_ext);
    states_kernel<Model, real><<<(n + 127) / 128, 128, 0, stream>>>((const real*)actions, (real*)states, a, mp, stride);
    CK(cudaGetLastError());
    return MPPI_OK;
}
template <class Model>
int run_rollout_states_dtype(const MppiFusedParams* p, const void* x0, const void* act, long long stride, int n, int T, void* out,
                             cudaStream_t s) {
    return p->dtype == MPPI_F32 ? run_rollout_states<Model, float>(p, x0, act, stride, n, T, out, s)
                                : run_rollout_states<Model, double>(p, x0, act, stride, n, T, out, s);
}

template <typename real, int V>
int run_apply(const MppiFusedParams* p, const void* partials, cudaStream_t stream) {
    KArgs<real> a;
    fill_kargs<real>(p, a, 128, 1);
    a.world = p->world <= 0 ? 1 : p->world;
    const int smem = align_up((2 * a.TN + a.R + a.T * a.S) * (int)sizeof(real), 16) + (a.R + 2) * 8 + 16;
    apply_partials_kernel<real, V><<<1, 128, smem, stream>>>(a, (const double*)partials, p->nu);
    CK(cudaGetLastError());
    return MPPI_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int mppi_b200_abi_version(void) { return MPPI_B200_ABI_VERSION; }

uint64_t mppi_abi_layout(int which) {
    switch (which) {
        case 0: return sizeof(MppiFusedParams);
        case 1: return offsetof(MppiFusedParams, seed);
        case 2: return offsetof(MppiFusedParams, state);
        case 3: return offsetof(MppiFusedParams, U);
        case 4: return offsetof(MppiFusedParams, rank);
        case 5: return offsetof(MppiFusedParams, partial_out);
        case 6: return sizeof(MppiLaunchInfo);
    }
    return 0;
}

const char* mppi_status_string(int s) {
    switch (s) {
        case MPPI_OK: return "ok";
        case MPPI_ERR_BAD_ARG: return "bad argument";
        case MPPI_ERR_UNSUPPORTED: return "unsupported (model, variant, dtype, size) combination";
        case MPPI_ERR_WORKSPACE: return "workspace too small";
        case MPPI_ERR_CUDA: return "CUDA error";
        case MPPI_ERR_ABI: return "MppiFusedParams size mismatch (ABI)";
        case MPPI_ERR_TIMEOUT: return "peer exchange timed out";
    }
    return "unknown status";
}

const char* mppi_last_cuda_error(void) { return g_cuda_err; }

int mppi_fused_query(const MppiFusedParams* p, MppiLaunchInfo* out) {
    if (out == nullptr) return MPPI_ERR_BAD_ARG;
    memset(out, 0, sizeof(*out));
    return dispatch_fused(p, nullptr, out);
}

int mppi_fused_command(const MppiFusedParams* p, void* stream) { return dispatch_fused(p, (cudaStream_t)stream, nullptr); }

int mppi_plan_create(const MppiFusedParams* p, void** plan_out) {
    if (plan_out == nullptr) return MPPI_ERR_BAD_ARG;
    int rc = validate(p, true);
    if (rc) return rc;
    Plan* pl = new (std::nothrow) Plan();
    if (pl == nullptr) return MPPI_ERR_BAD_ARG;
    switch (p->model) {
#ifndef MPPI_ONLY_USER_MODEL
        case MPPI_MODEL_PENDULUM: rc = build_plan_dtype<PendulumModel>(p, pl); break;
        case MPPI_MODEL_LINEAR_POINT: rc = build_plan_dtype<LinearPointModel>(p, pl); break;
        case MPPI_MODEL_PENDULUM_MLP:
            rc = (p->n_model_params_ext < PendulumMLPModel::N_EXT || p->model_params_ext == nullptr)
                     ? (int)MPPI_ERR_BAD_ARG : build_plan_dtype<PendulumMLPModel>(p, pl);
            break;
#endif
#ifdef MPPI_USER_MODEL_HEADER
        case MPPI_MODEL_USER: rc = build_plan_dtype<UserModel>(p, pl); break;
#endif
        default: rc = MPPI_ERR_UNSUPPORTED;
    }
    if (rc) {
        delete pl;
        return rc;
    }
    *plan_out = pl;
    return MPPI_OK;
}

int mppi_plan_destroy(void* plan) {
    if (plan == nullptr) return MPPI_ERR_BAD_ARG;
    Plan* pl = reinterpret_cast<Plan*>(plan);
    if (pl->res.armed) mppi_resident_stop(plan);      // the resident grid reads this controller's buffers
    delete pl;
    return MPPI_OK;
}

int mppi_plan_command(void* plan, const double* state, const void* state_dev, uint32_t flags, uint64_t seed, uint64_t offset,
                      const void* z, void* action_out, void* stream) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    if (pl == nullptr || action_out == nullptr) return MPPI_ERR_BAD_ARG;
    if (state == nullptr && !(flags & MPPI_FLAG_STATE_DEVICE)) return MPPI_ERR_BAD_ARG;
    if (pl->is_double) plan_update<double>(pl, state, state_dev, flags, seed, offset, z, action_out, nullptr);
    else plan_update<float>(pl, state, state_dev, flags, seed, offset, z, action_out, nullptr);
    return plan_launch(pl, (cudaStream_t)stream);
}

int mppi_plan_command_host(void* plan, const double* state, uint32_t flags, uint64_t seed, uint64_t offset, const void* z,
                           void* action_out_dev, void* host_mailbox, void* action_host_out, void* stream) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    if (pl == nullptr || state == nullptr || host_mailbox == nullptr || action_host_out == nullptr || action_out_dev == nullptr)
        return MPPI_ERR_BAD_ARG;
    if (pl->is_double) plan_update<double>(pl, state, nullptr, flags, seed, offset, z, action_out_dev, host_mailbox);
    else plan_update<float>(pl, state, nullptr, flags, seed, offset, z, action_out_dev, host_mailbox);
    int rc = plan_launch(pl, (cudaStream_t)stream);
    if (rc) return rc;
    // every action value arrives as self-validating 8-byte word(s): payload32 | (epoch & 0xffffffff) << 32
    volatile unsigned long long* box = reinterpret_cast<volatile unsigned long long*>(host_mailbox);
    const unsigned long long want = pl->host_epoch & 0xffffffffull;
    const int nwords = pl->upc_nu * (pl->is_double ? 2 : 1);
    unsigned long long spins = 0;
    for (int w = 0; w < nwords; ++w) {
        while ((box[w] >> 32) != want) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((++spins & 0xFFFFF) == 0) {       // every ~1M spins: has the stream died or finished without publishing?
                cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery while waiting for the mailbox");
                if (q == cudaSuccess && (box[w] >> 32) != want) return MPPI_ERR_TIMEOUT;
            }
        }
    }
    // the action is returned in the controller's own dtype (f32 words or f64 double-words)
    if (pl->is_double) {
        unsigned long long* out = reinterpret_cast<unsigned long long*>(action_host_out);
        for (int i = 0; i < pl->upc_nu; ++i) out[i] = (box[2 * i] & 0xffffffffull) | (box[2 * i + 1] << 32);
    } else {
        uint32_t* out = reinterpret_cast<uint32_t*>(action_host_out);
        for (int i = 0; i < pl->upc_nu; ++i) out[i] = (uint32_t)box[i];
    }
    return MPPI_OK;
}

int mppi_resident_start(void* plan, void* host_box, void* board_dev, void* action_out_dev, uint64_t idle_us, void* stream) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    if (pl == nullptr || host_box == nullptr || board_dev == nullptr || action_out_dev == nullptr || idle_us == 0)
        return MPPI_ERR_BAD_ARG;
    if (pl->res_kernel == nullptr)
        return UNSUPPORTED("resident mode needs a plan on the split-cost rollout (one tile per SM, registered analytic model, in-kernel exchange if sharded)");
    if (pl->p.z_out != nullptr) return UNSUPPORTED("resident mode does not record the noise it draws (z_out)");
    if (pl->res.launched) {                 // re-arming: the old grid leaves first (it polls the box that is about to be cleared)
        int rc = res_halt(pl->res);
        if (rc) return rc;
    }
    ResidentDevice& d = pl->resdev;
    d.host_box = reinterpret_cast<unsigned long long*>(host_box);
    d.board = reinterpret_cast<unsigned long long*>(board_dev);
    d.action_dev = action_out_dev;
    d.idle_ns = idle_us * 1000ull;
    d.stream = (cudaStream_t)stream;
    const ResidentBackend be{pl, resident_be_launch, resident_be_drain, resident_be_health};
    return res_arm(pl->res, host_box, pl->nx, pl->upc_nu, pl->is_double, be, pl->res_xchg);
}

int mppi_resident_command(void* plan, const double* state, uint32_t flags, uint64_t seed, uint64_t offset, void* action_host_out) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    if (pl == nullptr) return MPPI_ERR_BAD_ARG;
    // a sharded controller advances its exchange epoch once per command, on whichever route the command takes
    const uint64_t epoch = pl->res_xchg ? ++pl->epoch : 0;
    return res_command(pl->res, state, (flags & MPPI_FLAG_SHIFT) ? 1 : 0, seed, offset, action_host_out, epoch);
}

int mppi_resident_sync(void* plan) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    return pl == nullptr ? (int)MPPI_ERR_BAD_ARG : res_sync(pl->res);
}

int mppi_resident_stop(void* plan) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    return pl == nullptr ? (int)MPPI_ERR_BAD_ARG : res_stop(pl->res);
}

uint64_t mppi_resident_launches(void* plan) {
    Plan* pl = reinterpret_cast<Plan*>(plan);
    return pl == nullptr ? 0 : pl->res.launches;
}

int mppi_apply_partials(const MppiFusedParams* p, const void* partials, void* stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (partials == nullptr || p->nominal_used == nullptr || p->stats == nullptr || p->U == nullptr || p->action_out == nullptr)
        return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (p->dtype == MPPI_F32) {
        switch (p->variant) {
            case MPPI_VARIANT_MPPI: return run_apply<float, V_MPPI>(p, partials, s);
            case MPPI_VARIANT_SMPPI: return run_apply<float, V_SMPPI>(p, partials, s);
            case MPPI_VARIANT_KMPPI: return run_apply<float, V_KMPPI>(p, partials, s);
        }
    } else {
        switch (p->variant) {
            case MPPI_VARIANT_MPPI: return run_apply<double, V_MPPI>(p, partials, s);
            case MPPI_VARIANT_SMPPI: return run_apply<double, V_SMPPI>(p, partials, s);
            case MPPI_VARIANT_KMPPI: return run_apply<double, V_KMPPI>(p, partials, s);
        }
    }
    return MPPI_ERR_BAD_ARG;
}

uint64_t mppi_xchg_bytes(void) { return (uint64_t)2 * MPPI_MAX_RANKS * MPPI_XCHG_MAX_WORDS * sizeof(unsigned long long); }

int mppi_xchg_create(void** mailbox, void* ipc_handle_out_64B) {
    if (mailbox == nullptr || ipc_handle_out_64B == nullptr) return MPPI_ERR_BAD_ARG;
    void* ptr = nullptr;
    CK(cudaMalloc(&ptr, mppi_xchg_bytes()));
    CK(cudaMemset(ptr, 0, mppi_xchg_bytes()));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out_64B, &h, 64);
    *mailbox = ptr;
    return MPPI_OK;
}

int mppi_xchg_open(const void* ipc_handle_64B, void** peer_mailbox) {
    if (ipc_handle_64B == nullptr || peer_mailbox == nullptr) return MPPI_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64B, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_mailbox = ptr;
    return MPPI_OK;
}

int mppi_xchg_close(void* peer_mailbox) {
    if (peer_mailbox == nullptr) return MPPI_ERR_BAD_ARG;
    CK(cudaIpcCloseMemHandle(peer_mailbox));
    return MPPI_OK;
}

int mppi_xchg_destroy(void* mailbox) {
    if (mailbox == nullptr) return MPPI_ERR_BAD_ARG;
    CK(cudaFree(mailbox));
    return MPPI_OK;
}

int mppi_materialize(const MppiFusedParams* p, void* perturbed_action, void* noise, void* noise_theta, void* states,
                     void* stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (p->nominal_used == nullptr) return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (states != nullptr && perturbed_action == nullptr) return MPPI_ERR_BAD_ARG;
    rc = p->dtype == MPPI_F32
             ? run_sample_any<float>(p, perturbed_action, noise, noise_theta, nullptr, nullptr, 0, 0, true, s)
             : run_sample_any<double>(p, perturbed_action, noise, noise_theta, nullptr, nullptr, 0, 0, true, s);
    if (rc || states == nullptr) return rc;
    switch (p->model) {
#ifndef MPPI_ONLY_USER_MODEL
        case MPPI_MODEL_PENDULUM:
            return p->dtype == MPPI_F32 ? run_states<PendulumModel, float>(p, perturbed_action, states, s)
                                        : run_states<PendulumModel, double>(p, perturbed_action, states, s);
        case MPPI_MODEL_LINEAR_POINT:
            return p->dtype == MPPI_F32 ? run_states<LinearPointModel, float>(p, perturbed_action, states, s)
                                        : run_states<LinearPointModel, double>(p, perturbed_action, states, s);
#endif
#ifdef MPPI_USER_MODEL_HEADER
        case MPPI_MODEL_USER:
            return p->dtype == MPPI_F32 ? run_states<UserModel, float>(p, perturbed_action, states, s)
                                        : run_states<UserModel, double>(p, perturbed_action, states, s);
#endif
    }
    return MPPI_ERR_UNSUPPORTED;
}

int mppi_rollout_states(const MppiFusedParams* p, const void* start_states, const void* actions, int64_t actions_stride,
                        int32_t n_rollouts, int32_t T, void* states_out, void* stream) {
    if (p == nullptr || start_states == nullptr || actions == nullptr || states_out == nullptr) return MPPI_ERR_BAD_ARG;
    if (p->struct_size != sizeof(MppiFusedParams)) return MPPI_ERR_ABI;
    if (n_rollouts < 1 || T < 1 || actions_stride < 0) return MPPI_ERR_BAD_ARG;
    if (p->dtype != MPPI_F32 && p->dtype != MPPI_F64) return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    switch (p->model) {
#ifndef MPPI_ONLY_USER_MODEL
        case MPPI_MODEL_PENDULUM:
            return run_rollout_states_dtype<PendulumModel>(p, start_states, actions, actions_stride, n_rollouts, T, states_out, s);
        case MPPI_MODEL_LINEAR_POINT:
            return run_rollout_states_dtype<LinearPointModel>(p, start_states, actions, actions_stride, n_rollouts, T, states_out, s);
        case MPPI_MODEL_PENDULUM_MLP:
            if (p->n_model_params_ext < PendulumMLPModel::N_EXT || p->model_params_ext == nullptr) return MPPI_ERR_BAD_ARG;
            return run_rollout_states_dtype<PendulumMLPModel>(p, start_states, actions, actions_stride, n_rollouts, T, states_out, s);
#endif
#ifdef MPPI_USER_MODEL_HEADER
        case MPPI_MODEL_USER:
            return run_rollout_states_dtype<UserModel>(p, start_states, actions, actions_stride, n_rollouts, T, states_out, s);
#endif
    }
    return MPPI_ERR_UNSUPPORTED;
}

int mppi_sample_perturb(const MppiFusedParams* p, void* perturbed_action, void* noise, void* noise_theta, void* cost_init,
                        const void* override_rows, int32_t n_override, int32_t override_start, void* stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (p->U == nullptr || perturbed_action == nullptr || noise == nullptr || p->nominal_used == nullptr) return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    return p->dtype == MPPI_F32 ? run_sample_any<float>(p, perturbed_action, noise, noise_theta, cost_init, override_rows,
                                                        n_override, override_start, false, s)
                                : run_sample_any<double>(p, perturbed_action, noise, noise_theta, cost_init, override_rows,
                                                         n_override, override_start, false, s);
}

int mppi_cost_accumulate(void* cost, const void* c, void* var_acc, int32_t M, int32_t K, double discount, int32_t dtype,
                         void* stream) {
    if (cost == nullptr || c == nullptr || M < 1 || K < 1) return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int BD = 256, nb = (K + BD - 1) / BD;
    if (dtype == MPPI_F32)
        cost_accumulate_kernel<float><<<nb, BD, 0, s>>>((float*)cost, (const float*)c, (float*)var_acc, M, K, (float)discount);
    else if (dtype == MPPI_F64)
        cost_accumulate_kernel<double><<<nb, BD, 0, s>>>((double*)cost, (const double*)c, (double*)var_acc, M, K, discount);
    else
        return MPPI_ERR_BAD_ARG;
    CK(cudaGetLastError());
    return MPPI_OK;
}

int mppi_softmin_update(const MppiFusedParams* p, const void* cost_total, const void* eps, void* stream) {
    int rc = validate(p);
    if (rc) return rc;
    if (cost_total == nullptr || eps == nullptr || p->nominal_used == nullptr || p->U == nullptr || p->action_out == nullptr ||
        p->stats == nullptr)
        return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    return p->dtype == MPPI_F32 ? run_softmin_any<float>(p, cost_total, eps, s) : run_softmin_any<double>(p, cost_total, eps, s);
}

int mppi_omega(const void* cost_total, void* omega_out, const void* stats, double lambda_, int32_t K, int32_t dtype,
               void* stream) {
    if (cost_total == nullptr || omega_out == nullptr || stats == nullptr || K < 1 || lambda_ <= 0) return MPPI_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int BD = 256, nb = (K + BD - 1) / BD;
    if (dtype == MPPI_F32)
        omega_kernel<float><<<nb, BD, 0, s>>>((const float*)cost_total, (float*)omega_out, (const double*)stats,
                                              (float)(-(1.0 / lambda_)), K);
    else if (dtype == MPPI_F64)
        omega_kernel<double><<<nb, BD, 0, s>>>((const double*)cost_total, (double*)omega_out, (const double*)stats,
                                               -(1.0 / lambda_), K);
    else
        return MPPI_ERR_BAD_ARG;
    CK(cudaGetLastError());
    return MPPI_OK;
}

}  // extern "C"
