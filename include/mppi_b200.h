/*
 * mppi_b200.h — C ABI of the B200-native MPPI rollout-and-reweight engine.
 *
 * The reference (UM-ARM-Lab/pytorch_mppi) is pure Python and has no FFI; the drop-in boundary is
 * its Python class API.  This header is the boundary *below* that API: the Python
 * controllers in pytorch_mppi_b200/ bind these symbols with ctypes (INTEGRATION.md shows the stub a
 * reference maintainer would add to mppi.py to do the same).  Every entry point
 *   - takes plain pointers / sizes / a cudaStream_t passed as void*  (no torch types),
 *   - never allocates or frees caller-visible memory (buffers are caller-owned device memory),
 *   - launches asynchronously on the given stream and returns 0 or a negative MppiStatus,
 *   - is not re-entrant per controller (one controller <-> one stream, as the reference).
 *
 * Each entry point cites the reference code it replaces (file:line into
 * /root/reference/src/pytorch_mppi/mppi.py).
 */
#ifndef MPPI_B200_H
#define MPPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_B200_ABI_VERSION 1

#define MPPI_MAX_NU 4            /* control dimension supported by the fused registry            */
#define MPPI_MAX_NX 8            /* state dimension supported by the fused registry              */
#define MPPI_MODEL_PARAM_DOUBLES 48
#define MPPI_MAX_RANKS 8         /* GPUs of one NVSwitch box                                     */

typedef enum MppiStatus {
    MPPI_OK = 0,
    MPPI_ERR_BAD_ARG = -1,        /* null pointer / out-of-range dimension / unknown enum        */
    MPPI_ERR_UNSUPPORTED = -2,    /* (model, variant, dtype) not in the compiled registry        */
    MPPI_ERR_WORKSPACE = -3,      /* workspace too small (see mppi_fused_query)                  */
    MPPI_ERR_CUDA = -4,           /* a CUDA runtime call failed; see mppi_last_cuda_error()      */
    MPPI_ERR_ABI = -5,            /* struct_size does not match this library                     */
    MPPI_ERR_TIMEOUT = -6         /* peer exchange timed out (reported via the status word)      */
} MppiStatus;

typedef enum MppiVariant { MPPI_VARIANT_MPPI = 0, MPPI_VARIANT_SMPPI = 1, MPPI_VARIANT_KMPPI = 2 } MppiVariant;
typedef enum MppiDType { MPPI_F32 = 0, MPPI_F64 = 1 } MppiDType;

/* Registry of analytic models whose dynamics + running cost are compiled into the fused kernel.
 * (The reference takes arbitrary Python callables, mppi.py:147-157; arbitrary callables use the
 * per-step entry points further down.) */
typedef enum MppiModel {
    MPPI_MODEL_PENDULUM = 1,      /* /root/reference/tests/pendulum.py:30-60                      */
    MPPI_MODEL_LINEAR_POINT = 2,  /* tests/test_mppi.py:24-51 and tests/smooth_mppi.py:29-142    */
    MPPI_MODEL_PENDULUM_MLP = 3,  /* tests/pendulum_approximate.py:47-67 (3-32-32-2 tanh residual MLP) */
    MPPI_MODEL_USER = 100         /* a user-written model compiled into a variant of this library (build with
                                     -DMPPI_USER_MODEL_HEADER=\"file\"; see pytorch_mppi_b200.models.CudaModel): the
                                     reference accepts arbitrary Python callables (mppi.py:63-64) — this is the
                                     compiled-in equivalent for analytic dynamics/costs                        */
} MppiModel;

/* model_params layouts (doubles):
 *  PENDULUM     : [0]=g [1]=m [2]=l [3]=dt [4]=max_torque [5]=max_speed [6]=w_thdot(0.1)
 *  LINEAR_POINT : [0..3]=B(2x2 row-major) [4..5]=goal [6..9]=Q [10]=has_R [11..14]=R
 *                 [15]=terminal_scale [16]=n_hills(<=3) then per hill h at 17+7h:
 *                 Qh(4) centre(2) height(1)
 *  PENDULUM_MLP : [0]=max_torque [1]=w_thdot [2]=tanh_mode(0 exp-based, 1 MUFU.TANH) [3]=tensor_cores (0 = FFMA kernel;
 *                 tcgen05/TMEM kernel, fp32 accumulate: 1 = hi/lo-split bf16 operands, 2 = plain bf16; 3 = automatic:
 *                 mode 1 wherever that kernel exists — fp32, one environment — else the FFMA kernel); the 1,250 weights go in
 *                 model_params_ext: W1 (32x3 row-major), b1 (32), W2 (32x32), b2 (32), W3 (2x32), b3 (2)
 */

enum {
    MPPI_FLAG_SHIFT = 1u << 0,            /* shift_nominal_trajectory before sampling (mppi.py:232-238, 249-250) */
    MPPI_FLAG_NULL_ACTION = 1u << 1,      /* sample_null_action: global sample 0 is zeroed before the clamp (:390-392) */
    MPPI_FLAG_ABS_COST = 1u << 2,         /* noise_abs_cost (:190-191, 196-197)                                  */
    MPPI_FLAG_DIAG_SIGMA = 1u << 3,       /* diagonal covariance fast path (:131-136, 188-193, 204-205)          */
    MPPI_FLAG_STATE_DEVICE = 1u << 4,     /* read the state from state_dev instead of the by-value copy          */
    MPPI_FLAG_STATE_PER_SAMPLE = 1u << 5, /* state_dev is (K,nx): one start state per sample (:302-303)          */
    MPPI_FLAG_EXPORT_PARTIAL = 1u << 6,   /* multi-GPU, library collective: write this rank's (beta,eta,V) to
                                             partial_out and do NOT update U (mppi_apply_partials finishes)    */
    MPPI_FLAG_NOMINAL_PADDED = 1u << 7,   /* U (and A) are 16-byte aligned allocations padded to a multiple of
                                             16 bytes: lets the kernel stage them with one TMA bulk copy        */
    MPPI_FLAG_PDL = 1u << 8,              /* programmatic dependent launch: the kernel may start while the previous
                                             kernel on the stream is finishing; it draws its Philox normals (shared
                                             memory only) and touches global memory only after griddepcontrol.wait */
    MPPI_FLAG_WIDE_REGS = 1u << 10,       /* retired (accepted and ignored; the bit stays reserved): round 1's opt-in
                                             instantiation without the 64-register cap.  The split-cost kernels have the
                                             whole register file anyway, the single-loop kernels of large K need the cap. */
    MPPI_FLAG_SPLIT_COST = 1u << 9        /* small problems (threads_per_sample > 1): the rollout thread runs the bare state
                                             recurrence, the sample's helper threads evaluate the T running costs in
                                             parallel from the stored states, summed in the reference's order (same
                                             operations and rounding as the single-loop rollout; see mppi_fused.cuh)  */
};

/* One `command()` of a registered analytic model: replaces, in one launch,
 *   shift_nominal_trajectory            mppi.py:232-238 (SMPPI :489-493, KMPPI :617-619)
 *   _compute_perturbed_action_and_noise mppi.py:375-385 (SMPPI :539-552, KMPPI :657-670)
 *   _compute_rollout_costs_single       mppi.py:297-332
 *   _compute_total_cost_batch           mppi.py:407-417 (SMPPI :554-570)
 *   _compute_weighting                  mppi.py:254-259
 *   the einsum update in _command       mppi.py:268-270 (SMPPI :527-531, KMPPI :679-682)
 * All pointers are device pointers of element type `dtype` unless noted. */
typedef struct MppiFusedParams {
    uint32_t struct_size;        /* sizeof(MppiFusedParams), ABI check                                  */
    int32_t variant;             /* MppiVariant                                                         */
    int32_t model;               /* MppiModel                                                           */
    int32_t dtype;               /* MppiDType                                                           */
    int32_t K;                   /* samples rolled out by THIS rank                                     */
    int32_t T;                   /* horizon                                                             */
    int32_t nx, nu;              /* must match the model                                                */
    int32_t S;                   /* KMPPI: number of support points (else 0)                            */
    int32_t u_per_command;       /* rows of the action sequence copied to action_out                    */
    uint32_t flags;              /* MPPI_FLAG_*                                                         */
    int32_t block_threads;       /* samples per CTA tile (multiple of 32, <= 512); 0 = library default  */
    int32_t grid_blocks;         /* 0 = library default (<= SMs * resident CTAs)                        */
    int32_t threads_per_sample;  /* 1/2/4 threads share one sample's sampling work; 0 = library default */
    int64_t k_offset;            /* global index of this rank's first sample (keys the RNG, null action)*/
    uint64_t seed, offset;       /* Philox4x32-10 key / counter base (see oracle/philox_oracle.py)      */
    double lambda_;              /* temperature                                                         */
    double u_scale;
    double noise_mu[MPPI_MAX_NU];
    double chol[MPPI_MAX_NU * MPPI_MAX_NU];      /* row-major lower Cholesky factor L of noise_sigma; diag: sqrt on the diagonal */
    double sigma_inv[MPPI_MAX_NU * MPPI_MAX_NU]; /* row-major inverse covariance                                         */
    double u_min[MPPI_MAX_NU], u_max[MPPI_MAX_NU];
    double u_init[MPPI_MAX_NU];
    double action_min[MPPI_MAX_NU], action_max[MPPI_MAX_NU];   /* SMPPI (mppi.py:464-477)              */
    double w_action_seq_cost, delta_t;                         /* SMPPI (mppi.py:456-459)              */
    double model_params[MPPI_MODEL_PARAM_DOUBLES];
    double state[MPPI_MAX_NX];   /* start state by value: no host->device copy on the hot path          */
    const void* state_dev;       /* optional device state, (nx) or (K,nx)                               */
    void* U;                     /* (T,nu) in/out nominal controls (SMPPI: control derivative)          */
    void* A;                     /* SMPPI: (T,nu) in/out action_sequence                                */
    void* theta;                 /* KMPPI: (S,nu) in/out control points                                 */
    const void* W;               /* KMPPI: (T,S)  k(Hs,Tk) k(Tk,Tk)^-1                                  */
    const void* Wshift;          /* KMPPI: (S,S)  k(Tk+1,Tk) k(Tk,Tk)^-1                                */
    void* cost_total;            /* (K) out                                                             */
    void* action_out;            /* (u_per_command,nu) out                                              */
    void* nominal_used;          /* out, 3*(T*nu) elements: U | A | theta as used for sampling (post-shift,
                                    pre-update); lets noise/perturbed_action be materialised lazily     */
    void* stats;                 /* out, 4 doubles: beta, eta, (reserved), status word                  */
    const void* z;               /* optional injected standard normals (K,T,nu) / KMPPI (K,S,nu); NULL = Philox */
    void* z_out;                 /* optional: the standard normals actually used, same shape            */
    void* workspace;             /* caller-owned scratch, zero-initialised ONCE by the caller           */
    uint64_t workspace_bytes;
    /* ---- multi-GPU (K sharded over ranks) ---- */
    int32_t rank, world;
    uint64_t epoch;              /* strictly increasing per command; keys the peer-exchange flags        */
    void* peer_slots[MPPI_MAX_RANKS];   /* in-kernel exchange: pointer to every rank's mailbox (own included),
                                           from mppi_xchg_*; all NULL = no in-kernel exchange            */
    void* partial_out;           /* MPPI_FLAG_EXPORT_PARTIAL: (2 + R) doubles out                        */
    /* ---- batched environments (MPPI_Batched, mppi.py:691-873): n_env independent problems sharing one noise
     * stream, launched as gridDim.y = n_env.  Buffers are (n_env, ...) with these strides; state_dev is
     * (n_env, nx) on the device.  n_env <= 1 means a single problem.  MPPI variant only. */
    int32_t n_env;
    int32_t env_u_stride;        /* ELEMENTS between consecutive environments' U (>= T*nu, 16-byte multiple for TMA) */
    uint64_t env_ws_stride;      /* BYTES between consecutive environments' workspace slices                       */
    void* host_mailbox;          /* optional PINNED HOST memory (device-visible under UVA), u_per_command*nu 8-byte
                                    words (x2 for f64): the kernel stores every action value as a self-validating
                                    word, payload32 | (host_epoch & 0xffffffff) << 32, so the host spins on the
                                    words themselves instead of issuing a D2H copy or a stream synchronise      */
    uint64_t host_epoch;
    uint64_t torch_rng_total;    /* 0: the engine's own stream (one Philox subsequence per sample).  > 0: reproduce the
                                    stream of `torch.randn(K,T,nu, device="cuda")` (ATen/native/cuda/
                                    DistributionTemplates.h:50-100): element li of the flat (K,T,nu) tensor comes from
                                    thread idx = li % total, loop slot (li / total), with total = 256 * grid — so a
                                    reference controller on device="cuda" and this engine draw identical noise from
                                    the same (seed, offset).  `offset` is then the generator offset / 4.            */
    void* offset_dev;            /* optional DEVICE u64: when set, the Philox counter base is read from here instead of
                                    `offset`, and the finishing kernel adds `offset_inc` to it — lets a captured CUDA
                                    graph of a whole command draw fresh noise on every replay                       */
    uint64_t offset_inc;
    const double* model_params_ext;   /* HOST pointer: extra model parameters (e.g. MLP weights), read at call /
                                         plan-creation time; NULL if the model needs none                    */
    int32_t n_model_params_ext;
    int32_t K_geom;              /* 0, or the sample count the launch geometry is planned for (>= K): every shard of a
                                    multi-GPU controller passes the LARGEST shard so that all ranks launch the same
                                    grid and publish the same number of records per command                */
    void* debug_clocks;          /* optional profiling aid: (grid_blocks, 16) uint64 %globaltimer stamps (ns) at
                                    phase boundaries of each CTA's first tile; NULL on the product path   */
    void* xchg_status_host;      /* optional pinned HOST int64: the kernel stores MPPI_ERR_TIMEOUT there when a peer
                                    exchange timed out (the caller checks it before the next command)      */
    void* user_model;            /* model == MPPI_MODEL_USER: handle from mppi_user_model_register (a model compiled at
                                    run time), or NULL for a model linked into a variant library            */
} MppiFusedParams;

typedef struct MppiLaunchInfo {
    int32_t block_threads, grid_blocks;
    int32_t smem_bytes, regs_per_thread;
    int32_t max_blocks_per_sm, sm_count;
    uint64_t workspace_bytes;    /* minimum workspace for these dimensions                               */
    int32_t tma_staging;         /* 1 if the nominal sequence is staged with cp.async.bulk (TMA)         */
    int32_t threads_per_sample;
    int32_t split_cost;          /* 1 if MPPI_FLAG_SPLIT_COST was honoured for these dimensions          */
    int32_t wide_regs;           /* always 0 (MPPI_FLAG_WIDE_REGS is retired)                            */
    int32_t cluster_size;        /* thread-block-cluster size of the launch (1 = no cluster): the CTAs of a cluster
                                    reduce their softmin partials through distributed shared memory        */
    int32_t xchg_records;        /* sharded controllers: records each rank publishes per command (its cluster records
                                    in direct mode, 1 = the rank's combined record); 0 = not sharded       */
} MppiLaunchInfo;

int mppi_b200_abi_version(void);
/* Layout probe for foreign-language mirrors of MppiFusedParams: which=0 sizeof, 1 offsetof(seed),
 * 2 offsetof(state), 3 offsetof(U), 4 offsetof(rank), 5 offsetof(partial_out), 6 sizeof(MppiLaunchInfo). */
uint64_t mppi_abi_layout(int which);
const char* mppi_status_string(int status);
const char* mppi_last_cuda_error(void);

/* Launch geometry / scratch needs for these parameters (no launch). */
int mppi_fused_query(const MppiFusedParams* p, MppiLaunchInfo* out);
/* The fused command (see struct comment).  `stream` is a cudaStream_t. */
int mppi_fused_command(const MppiFusedParams* p, void* stream);

/* ---- Plans: the steady-state fast path ----------------------------------------------------------
 * A plan freezes everything `mppi_fused_command` derives from an MppiFusedParams (kernel selection,
 * launch geometry, the pre-cast kernel argument block) so that one command costs one short call.
 * Per command only the start state, the RNG counter / injected-noise pointer, the flags that vary
 * (SHIFT, STATE_DEVICE, STATE_PER_SAMPLE) and the action destination change.  The multi-GPU epoch is
 * advanced inside the plan (every rank issues the same sequence of commands). */
int mppi_plan_create(const MppiFusedParams* p, void** plan_out);
int mppi_plan_destroy(void* plan);
/* One command(); `state` = nx host doubles (by value into the launch) or NULL with `state_dev` set. */
int mppi_plan_command(void* plan, const double* state, const void* state_dev, uint32_t flags, uint64_t seed,
                      uint64_t offset, const void* z, void* action_out, void* stream);
/* One command() for a host-resident control loop: launches, then spins (in C) on the pinned-host
 * mailbox the kernel stores the flagged action words into, and returns the action in
 * `action_host_out` (u_per_command*nu elements of the controller dtype).  No D2H memcpy call, no stream synchronise.
 * `host_mailbox`: u_per_command*nu*8 bytes (x2 for f64) of zero-initialised pinned host memory. */
int mppi_plan_command_host(void* plan, const double* state, uint32_t flags, uint64_t seed, uint64_t offset,
                           const void* z, void* action_out_dev, void* host_mailbox, void* action_host_out, void* stream);

/* ---- Resident mode: command() without a kernel launch ---------------------------------------------
 * For a host control loop (mppi.py:240-252 called once per control step).  One cooperative launch leaves the
 * command's grid on the SMs; each following command is a record (flags, Philox counter, start state) written into
 * pinned host memory that the grid polls, and the action comes back the way mppi_plan_command_host returns it.
 * The grid prepares the next command's noise and perturbed actions while the host turns around, and leaves by
 * itself after `idle_us` without a command (the next command relaunches it), so a device-wide synchronise is
 * never blocked for longer than that.  Results are bit-identical to mppi_plan_command_host with the same
 * (state, flags, seed, offset).  Available for plans that run the split-cost rollout (one tile per SM;
 * MPPI_FLAG_SPLIT_COST honoured — MppiLaunchInfo.split_cost) of the analytic registry models; others: MPPI_ERR_UNSUPPORTED.
 * A plan that is one shard of a multi-GPU controller with the in-kernel exchange (peer_slots set) runs the instantiation
 * whose records also carry the exchange epoch: every rank's host posts its own record, the finishers exchange over
 * NVLink inside the resident grids (validated on one GPU in round 1; the sharded instantiation is to be run on two).
 *
 *   host_box      pinned host memory, zero-initialised, 64 + u_per_command*nu (x2 for f64) 8-byte words:
 *                 [0,32) command record | [32] sequence number of the last finished command | [33] exit word |
 *                 [64,..) action words (payload32 | seq32 << 32)
 *   board_dev     device memory, 128 8-byte words (the record re-published for all CTAs + the done word)
 *   action_out_dev device (u_per_command,nu) of the controller dtype (the action as on the launch route)
 *   stream        a NON-BLOCKING stream used for nothing else: work on other streams is not ordered against the
 *                 resident grid; call mppi_resident_sync before reading U / cost_total / stats, mppi_resident_stop
 *                 before writing U or launching mppi_plan_command on the same controller.
 * While resident, the controller's parameters are frozen (they are kernel arguments of the resident launch). */
int mppi_resident_start(void* plan, void* host_box, void* board_dev, void* action_out_dev, uint64_t idle_us, void* stream);
/* One command(): `state` = nx host doubles; flags: MPPI_FLAG_SHIFT is read; returns the action in action_host_out
 * (u_per_command*nu elements of the controller dtype). */
int mppi_resident_command(void* plan, const double* state, uint32_t flags, uint64_t seed, uint64_t offset,
                          void* action_host_out);
/* Wait until everything the last command wrote to device memory (U, cost_total, nominal_used, stats) is complete. */
int mppi_resident_sync(void* plan);
/* Send the grid away and wait for it; mppi_plan_destroy does this too. */
int mppi_resident_stop(void* plan);
/* Kernel launches made by resident mode so far (the first command, and each wake-up after an idle exit). */
uint64_t mppi_resident_launches(void* plan);

/* Multi-GPU, library-collective route: after every rank exported its partial and the caller
 * all-gathered them (NCCL), finish the update on each rank: beta=min, rescale, U += sum/eta
 * (mppi.py:254-259, 268-270 across shards).  `partials` is (world, 2+R) doubles on device. */
int mppi_apply_partials(const MppiFusedParams* p, const void* partials, void* stream);

/* Peer mailboxes for the in-kernel NVLink exchange.  The library owns these small buffers
 * (cudaMalloc + cudaIpc), because IPC handles must cover a whole allocation. */
/* A user-written analytic model compiled at run time (the reference's plugin surface is arbitrary Python callables,
 * mppi.py:63-64; analytic ones can be given as CUDA C++ bodies — pytorch_mppi_b200.models.CudaModel — and are compiled
 * with NVRTC in process).  `cubin` is the compiled module, `names` the lowered names of its kernels:
 * [0] fused_command_kernel<UserModel, real, variant, false, false>   [1] ... split-cost (or NULL)
 * [2] ... batched, MPPI variant (or NULL)   [3] resident_command_kernel<UserModel, real, variant> (or NULL)
 * [4] states_kernel<UserModel, real>.  The handle goes into MppiFusedParams.user_model with model = MPPI_MODEL_USER. */
int mppi_user_model_register(const void* cubin, uint64_t cubin_bytes, int32_t nx, int32_t nu, int32_t n_params, int32_t dtype,
                             int32_t variant, const char* const* names, void** handle_out);
int mppi_user_model_release(void* handle);
/* Highest command epoch this plan has used (the tag of its reduction records, shared by the launch and resident routes);
 * pass it as MppiFusedParams.epoch when re-creating the plan so that stale records can never match. */
uint64_t mppi_plan_epoch(void* plan);
int mppi_xchg_create(void** mailbox, void* ipc_handle_out_64B);       /* local mailbox + its IPC handle */
int mppi_xchg_open(const void* ipc_handle_64B, void** peer_mailbox);  /* map a peer's mailbox           */
int mppi_xchg_close(void* peer_mailbox);
int mppi_xchg_destroy(void* mailbox);
uint64_t mppi_xchg_bytes(void);

/* ---- API-visible intermediates, produced on demand (off the hot path) -------------------------
 * The reference stores noise / perturbed_action (mppi.py:383-385) and, with a terminal cost,
 * states / actions (mppi.py:307-322) on every command.  The fused kernel keeps them in shared
 * memory only; these entry points regenerate them from (seed, offset | z) and nominal_used. */
int mppi_materialize(const MppiFusedParams* p, void* perturbed_action /*(K,T,nu)*/, void* noise /*(K,T,nu)*/,
                     void* noise_theta /*KMPPI (K,S,nu) or NULL*/, void* states /*(K,T,nx) or NULL*/, void* stream);

/* get_rollouts (mppi.py:425-448) for a registered model: `n_rollouts` start states (n_rollouts,nx) on the device, each
 * rolled T steps through u_scale * actions[t] — one (T,nu) sequence replayed by every rollout (actions_stride = 0) or a
 * sequence per rollout (actions_stride = elements between them, e.g. T*nu) — into states_out (n_rollouts,T,nx).
 * Only model, dtype, nx, nu, u_scale and the model parameters of `p` are read. */
int mppi_rollout_states(const MppiFusedParams* p, const void* start_states, const void* actions, int64_t actions_stride,
                        int32_t n_rollouts, int32_t T, void* states_out, void* stream);

/* ---- Per-step entry points for arbitrary Python dynamics/cost callables ------------------------
 * The T-loop stays in Python (mppi.py:312-322); sampling, cost accumulation and the softmin update
 * are kernels. */

/* mppi.py:375-385 + 186-199 + 415 (SMPPI :539-562, KMPPI :657-670): fills perturbed_action, noise
 * (and noise_theta) and initialises cost_init[k] = sum_t U . action_cost (+ SMPPI smoothness).
 * `override` (n_override,T,nu) rows replace samples [override_start, ...) before the clamp
 * (SpecificActionSampler, mppi.py:393-399). */
int mppi_sample_perturb(const MppiFusedParams* p, void* perturbed_action, void* noise, void* noise_theta,
                        void* cost_init, const void* override_rows, int32_t n_override, int32_t override_start,
                        void* stream);
/* mppi.py:318-319 / 362-364: cost[m,k] += c[m,k]; M>1 also var_acc[k] += var_m(c) * discount. */
int mppi_cost_accumulate(void* cost /*(M,K)*/, const void* c /*(M*K)*/, void* var_acc /*(K) or NULL*/,
                         int32_t M, int32_t K, double discount, int32_t dtype, void* stream);
/* mppi.py:254-259 + 268-270 from materialised tensors: cost_total (K) and eps = noise (K,T,nu)
 * (KMPPI: noise_theta (K,S,nu)).  Reads the post-shift nominal from p->nominal_used (written by
 * mppi_sample_perturb), writes p->U / A / theta, p->action_out and p->stats. */
int mppi_softmin_update(const MppiFusedParams* p, const void* cost_total, const void* eps, void* stream);
/* omega (mppi.py:256-258) from cost_total and the (beta, eta) a command left in `stats`. */
int mppi_omega(const void* cost_total, void* omega_out, const void* stats, double lambda_, int32_t K, int32_t dtype,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H */
