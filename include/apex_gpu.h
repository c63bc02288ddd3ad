/*
 * apex_gpu.h — C ABI of the B200-native bundle-adjustment Levenberg–Marquardt path.
 *
 * This is the drop-in boundary a Rust shim in apex-solver's reserved module
 * `src/linearizer/gpu/mod.rs:1-5` (and a new `linalg::gpu`) would bind with
 * `extern "C"`; INTEGRATION.md shows that shim. Every entry point cites the
 * reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers + sizes, `repr(C)`-compatible PODs, no callbacks, no C++ or
 *     torch types; never throws across the boundary.
 *   - every function returns an apex_status: 0 = ok, negative = error (mapped 1:1
 *     on the reference's error enums, see below). apex_last_error() gives the text.
 *   - a context is single-threaded (`&mut self` semantics of `LinearSolver`,
 *     src/linalg/mod.rs:143-180); one context per GPU / rank.
 *   - all arithmetic is FP64; indices are u32 (the reference uses usize).
 *   - THERE IS NO CPU FALLBACK: without a CUDA device every compute entry point
 *     returns APEX_ERR_NO_DEVICE.
 *
 * The same declarations (with the prefix `oracle_` instead of `apex_`) are
 * implemented on the CPU by oracle/apex_oracle.cpp, the test oracle.
 */
#ifndef APEX_GPU_H
#define APEX_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t apex_status;

/* ---- status codes ------------------------------------------------------- */
#define APEX_OK 0
/* LinAlgError (src/linalg/mod.rs:77-101) */
#define APEX_ERR_FACTORIZATION_FAILED (-1) /* LinAlgError::FactorizationFailed  */
#define APEX_ERR_SINGULAR_MATRIX (-2)      /* LinAlgError::SingularMatrix       */
#define APEX_ERR_INVALID_INPUT (-5)        /* LinAlgError::InvalidInput / LinearizerError::InvalidInput (src/linearizer/mod.rs:65-85) */
#define APEX_ERR_INVALID_STATE (-6)        /* LinAlgError::InvalidState         */
/* OptimizerError (src/optimizer/mod.rs:66-141) */
#define APEX_ERR_LINEAR_SOLVE_FAILED (-10) /* OptimizerError::LinearSolveFailed */
#define APEX_ERR_INVALID_PARAMETERS (-11)  /* OptimizerError::InvalidParameters */
#define APEX_ERR_NUMERICAL_INSTABILITY (-12)
#define APEX_ERR_EMPTY_PROBLEM (-13)       /* OptimizerError::EmptyProblem      */
#define APEX_ERR_NO_RESIDUAL_BLOCKS (-14)  /* OptimizerError::NoResidualBlocks  */
/* boundary-only */
#define APEX_ERR_CUDA (-20)        /* CUDA runtime error                         */
#define APEX_ERR_NCCL (-21)        /* NCCL error                                 */
#define APEX_ERR_NO_DEVICE (-22)   /* no CUDA device / library built without one */
#define APEX_ERR_UNSUPPORTED (-23) /* valid in the reference, not on this path   */
/* IoError (crates/apex-io/src/lib.rs:52-82), BAL loader */
#define APEX_ERR_IO (-30)             /* IoError::Io (file cannot be read / written)            */
#define APEX_ERR_PARSE (-31)          /* IoError::Parse{line,message}                           */
#define APEX_ERR_INVALID_NUMBER (-32) /* IoError::InvalidNumber{line,value}                     */
#define APEX_ERR_MISSING_FIELDS (-33) /* IoError::MissingFields{line}                           */

/* ---- enums ---------------------------------------------------------------- */

/* Camera model of the projection factor (crates/apex-camera-models/src/, one file per model).
 * Intrinsic vector layouts follow the reference's `From<&Camera> for DVector`. */
enum apex_camera_model {
  APEX_CAM_BAL = 0,            /* BALPinholeCameraStrict [f,k1,k2]             bal_pinhole.rs:242 */
  APEX_CAM_PINHOLE = 1,        /* PinholeCamera [fx,fy,cx,cy]                   pinhole.rs:202     */
  APEX_CAM_KANNALA_BRANDT = 2, /* [fx,fy,cx,cy,k1,k2,k3,k4]                     kannala_brandt.rs:358 */
  APEX_CAM_DOUBLE_SPHERE = 3,  /* [fx,fy,cx,cy,xi,alpha]                        double_sphere.rs:332 */
  APEX_CAM_RADTAN = 4,         /* [fx,fy,cx,cy,k1,k2,p1,p2,k3]                  rad_tan.rs:333     */
  APEX_CAM_UCM = 5,            /* [fx,fy,cx,cy,alpha]                           ucm.rs:300         */
  APEX_CAM_EUCM = 6,           /* [fx,fy,cx,cy,alpha,beta]                      eucm.rs:320        */
  APEX_CAM_FOV = 7,            /* [fx,fy,cx,cy,w]                               fov.rs:291         */
  APEX_CAM_FTHETA = 8          /* [cx,cy,k1..k4]                                ftheta.rs:224      */
};

/* OptimizeParams<POSE,LANDMARK,INTRINSIC> (src/factors/mod.rs:83-101). Only the two
 * modes that are live in bin/bundle_adjustment.rs are supported:
 * BundleAdjustment = POSE|LANDMARK, SelfCalibration = POSE|LANDMARK|INTRINSIC. */
#define APEX_OPT_POSE 1u
#define APEX_OPT_LANDMARK 2u
#define APEX_OPT_INTRINSIC 4u
/* With APEX_OPT_INTRINSIC: ALL cameras share ONE intrinsics variable - the graph of the reference's calibration tests
 * (tests/camera_*_integration.rs: one multi-observation ProjectionFactor per camera over [pose_k, landmarks, intrinsics],
 * src/factors/projection_factor.rs:184-364; an n-observation block is n rows of this SoA description). intr[0] is the shared
 * value (rows 1.. are ignored on upload and returned equal to row 0); explicit Schur + Cholesky only, and the loss must be
 * NONE or L2 (a robust loss on an n-observation block weighs the block's whole squared norm, not each observation). */
#define APEX_OPT_SHARED_INTRINSICS 8u

/* LossFunction implementations (src/core/loss_functions.rs). params[] meaning per id. */
enum apex_loss {
  APEX_LOSS_NONE = 0,          /* residual block without a loss (loss_func = None)             */
  APEX_LOSS_L2 = 1,            /* L2Loss                        :174                            */
  APEX_LOSS_L1 = 2,            /* L1Loss                        :236                            */
  APEX_LOSS_HUBER = 3,         /* HuberLoss{scale=p0}           :353                            */
  APEX_LOSS_CAUCHY = 4,        /* CauchyLoss{scale=p0}          :486                            */
  APEX_LOSS_FAIR = 5,          /* FairLoss{scale=p0}            :585                            */
  APEX_LOSS_GEMAN_MCCLURE = 6, /* GemanMcClureLoss{scale=p0}    :674                            */
  APEX_LOSS_WELSCH = 7,        /* WelschLoss{scale=p0}          :759                            */
  APEX_LOSS_TUKEY = 8,         /* TukeyBiweightLoss{scale=p0}   :848                            */
  APEX_LOSS_ANDREWS = 9,       /* AndrewsWaveLoss{scale=p0}     :949                            */
  APEX_LOSS_RAMSAY_EA = 10,    /* RamsayEaLoss{scale=p0}        :1037                           */
  APEX_LOSS_TRIMMED_MEAN = 11, /* TrimmedMeanLoss{scale=p0}     :1132                           */
  APEX_LOSS_LP_NORM = 12,      /* LpNormLoss{p=p0}              :1207                           */
  APEX_LOSS_BARRON = 13,       /* BarronGeneralLoss{alpha=p0,scale=p1} :1316 (AdaptiveBarron delegates, :1569) */
  APEX_LOSS_T_DISTRIBUTION = 14 /* TDistributionLoss{nu=p0}     :1445                           */
};

/* LinearSolverType::{ExplicitSchur, ImplicitSchur} as spelled in README.md:256-257;
 * today's code spells them SparseSchurComplement + SchurVariant::{Sparse,Iterative}
 * (src/linalg/sparse/explicit_schur.rs:58-65). */
enum apex_schur_variant {
  /* SchurVariant::Sparse: explicit S, direct Cholesky (explicit_schur.rs:539-634). */
  APEX_SCHUR_EXPLICIT = 0,
  /* Matrix-free PCG with block preconditioner = the math of IterativeSchurSolver
   * (src/linalg/sparse/implicit_schur.rs:163-946), wired with the +J^T r gradient. */
  APEX_SCHUR_IMPLICIT = 1,
  /* SchurVariant::Iterative exactly as dispatched today: explicit S + scalar-Jacobi PCG
   * (explicit_schur.rs:639-756, reached from :1224-1225). */
  APEX_SCHUR_EXPLICIT_PCG = 2
};

/* SchurPreconditioner (explicit_schur.rs:67-78); used by APEX_SCHUR_IMPLICIT only. */
enum apex_schur_preconditioner {
  APEX_PRECOND_NONE = 0,
  APEX_PRECOND_BLOCK_DIAGONAL = 1,
  APEX_PRECOND_SCHUR_JACOBI = 2
};

/* OptimizationStatus (src/optimizer/mod.rs:189-216), same order. */
enum apex_optimization_status {
  APEX_STATUS_CONVERGED = 0,
  APEX_STATUS_MAX_ITERATIONS_REACHED = 1,
  APEX_STATUS_COST_TOLERANCE_REACHED = 2,
  APEX_STATUS_PARAMETER_TOLERANCE_REACHED = 3,
  APEX_STATUS_GRADIENT_TOLERANCE_REACHED = 4,
  APEX_STATUS_NUMERICAL_FAILURE = 5,
  APEX_STATUS_USER_TERMINATED = 6,
  APEX_STATUS_TIMEOUT = 7,
  APEX_STATUS_TRUST_REGION_RADIUS_TOO_SMALL = 8,
  APEX_STATUS_MIN_COST_THRESHOLD_REACHED = 9,
  APEX_STATUS_ILL_CONDITIONED_JACOBIAN = 10,
  APEX_STATUS_INVALID_NUMERICAL_VALUES = 11,
  APEX_STATUS_FAILED = 12
};

/* ---- PODs ------------------------------------------------------------------ */

typedef struct apex_ctx apex_ctx; /* opaque */

/* One LossFunction instance: enum apex_loss + its parameters (params[] meaning per id above). */
typedef struct apex_loss_spec {
  int32_t loss_id;
  int32_t reserved;
  double params[4];
} apex_loss_spec;

typedef struct apex_ctx_desc {
  int32_t device;           /* CUDA device ordinal                                       */
  int32_t rank;             /* 0..nranks-1                                               */
  int32_t nranks;           /* 1 = single GPU                                            */
  int32_t reserved;
  const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks, or NULL when nranks==1 */
} apex_ctx_desc;

/* The factor graph bin/bundle_adjustment.rs:212-441 builds (one ProjectionFactor with a single
 * 2x1 observation per residual block, Problem::add_residual_block src/core/problem.rs:575-598),
 * flattened to SoA. Host pointers, caller-owned, copied by apex_problem_upload. With nranks>1
 * every rank passes the SAME full problem; the library keeps its own shard of points. */
typedef struct apex_problem_desc {
  int32_t camera_model;  /* enum apex_camera_model                                        */
  uint32_t opt_flags;    /* APEX_OPT_*                                                    */
  int32_t intr_dim;      /* K = CameraModel::INTRINSIC_DIM of camera_model                */
  /* 1 if an "intr_XXXX" variable exists per camera even though no factor references it
   * (bin/bundle_adjustment.rs:243-246 always inserts them): they then own K all-zero columns
   * (diagonal = lambda) and enter the parameter norm. Ignored when APEX_OPT_INTRINSIC is set. */
  int32_t intr_vars_present;
  uint32_t ncam;
  uint32_t npts;
  uint64_t nobs;
  const double* pose;       /* [ncam][7] = tx,ty,tz,qw,qx,qy,qz (se3.rs:200-222), world->camera */
  const double* intr;       /* [ncam][K]                                                 */
  const double* pt;         /* [npts][3]                                                 */
  const uint32_t* obs_cam;  /* [nobs] camera index of each residual block, insertion order */
  const uint32_t* obs_pt;   /* [nobs] landmark index                                     */
  const double* obs_uv;     /* [nobs][2] measured pixel                                  */
  int32_t loss_id;          /* enum apex_loss, uniform over blocks (unless obs_loss is given) */
  int32_t reserved0;
  double loss_params[4];
  /* Problem::fix_variable (src/core/problem.rs:609-616): bit d set = tangent DOF d is fixed
   * (step zeroed at update time only, src/core/problem.rs:185-289). NULL = nothing fixed. */
  const uint8_t* pose_fixed;  /* [ncam] bits 0..5                                        */
  const uint16_t* intr_fixed; /* [ncam] bits 0..K-1                                      */
  const uint8_t* pt_fixed;    /* [npts] bits 0..2                                        */
  /* Per-block loss functions: every ResidualBlock owns its own Option<Box<dyn LossFunction>>
   * (src/core/residual_block.rs:97-123). obs_loss[o] indexes loss_table (at most 256 distinct instances - the shim
   * de-duplicates boxes by type and parameters); NULL = every block uses loss_id / loss_params above. */
  const uint8_t* obs_loss;          /* [nobs] index into loss_table, or NULL                */
  const apex_loss_spec* loss_table; /* [n_losses]                                           */
  int32_t n_losses;
  int32_t reserved1;
} apex_problem_desc;

/* LevenbergMarquardtConfig (src/optimizer/levenberg_marquardt.rs:213-317), field for field, plus the
 * Schur solver's CG parameters (explicit_schur.rs:211-212). Dead fields of the reference
 * (damping_increase_factor ... min_relative_decrease) are carried but unused. */
typedef struct apex_lm_config {
  int32_t schur_variant;        /* enum apex_schur_variant                                */
  int32_t schur_preconditioner; /* enum apex_schur_preconditioner                         */
  int32_t max_iterations;       /* default 50; for_bundle_adjustment(): 20 (:519-530)     */
  int32_t cg_max_iterations;    /* 200                                                    */
  double cost_tolerance;        /* 1e-6  */
  double parameter_tolerance;   /* 1e-8  */
  double gradient_tolerance;    /* 1e-10 */
  double timeout_seconds;       /* <= 0: None */
  double damping;               /* 1e-3  */
  double damping_min;           /* 1e-12 */
  double damping_max;           /* 1e12  */
  double damping_increase_factor; /* dead */
  double damping_decrease_factor; /* dead */
  double damping_nu;            /* 2.0   */
  double trust_region_radius;   /* 1e4 (only forwarded to check_convergence) */
  double min_step_quality;      /* dead */
  double good_step_quality;     /* dead */
  double min_diagonal;          /* dead */
  double max_diagonal;          /* dead */
  double min_cost_threshold;    /* NaN: None */
  double min_trust_region_radius; /* 1e-32 */
  double max_condition_number;  /* dead; NaN: None */
  double min_relative_decrease; /* dead */
  double cg_tolerance;          /* 1e-6  */
  int32_t use_jacobi_scaling;   /* column scaling 1/(1+||J col||) fixed at iteration 0 (:871-879; default false, :352) */
  int32_t compute_covariances;  /* accepted, no effect: the Schur solvers return no covariance matrix (linalg/mod.rs:170-172) */
} apex_lm_config;

/* SolverResult + ConvergenceInfo (src/optimizer/mod.rs:163-172,250-273). */
typedef struct apex_lm_result {
  int32_t status;     /* enum apex_optimization_status                                    */
  int32_t iterations; /* = iteration + 1 at termination (levenberg_marquardt.rs:1015)     */
  double initial_cost;
  double final_cost;
  double elapsed_seconds;
  double final_gradient_norm;
  double final_parameter_update_norm;
  int32_t cost_evaluations;
  int32_t jacobian_evaluations;
  int32_t successful_steps;
  int32_t unsuccessful_steps;
  double final_damping;     /* config.damping is mutated by a solve (:706-714)             */
  double final_damping_nu;
  int64_t linear_iterations; /* total PCG iterations (0 for the direct variant)            */
} apex_lm_result;

/* IterationStats (src/optimizer/mod.rs:375-398) + linear-solver iterations. */
typedef struct apex_iter_trace {
  int32_t iteration;
  int32_t accepted;
  int32_t ls_iter; /* PCG iterations of this LM iteration                                  */
  int32_t reserved;
  double cost;          /* state.current_cost after accept/reject                          */
  double cost_change;
  double gradient_norm;
  double step_norm;
  double tr_ratio;      /* rho                                                             */
  double tr_radius;     /* damping after update_damping                                    */
  double new_cost;      /* cost at the trial point                                         */
  double predicted_reduction;
  double parameter_norm;
  double iter_time_ms;
} apex_iter_trace;

/* Sizes of the current problem as seen by the solver. */
typedef struct apex_dims {
  uint32_t ncam, npts;
  uint64_t nobs;
  int32_t intr_dim;    /* K                                                               */
  int32_t dc;          /* camera-side DOF per camera inside the reduced system (6 or 6+K)  */
  uint64_t cam_dof;    /* rows of S in the reference layout (includes unreferenced intr columns) */
  uint64_t lm_dof;     /* 3*npts                                                          */
  uint32_t npts_local; /* points owned by this rank                                        */
  uint32_t flags;      /* bit 0: the NVLink peer-memory all-reduce is active (nranks > 1)  */
  uint64_t nobs_local;
} apex_dims;

/* ---- entry points ---------------------------------------------------------- */

/* Library/ABI version (major*100+minor) and whether a CUDA device is usable. */
int32_t apex_abi_version(void);
int32_t apex_device_count(void);

/* Fill `cfg` with LevenbergMarquardtConfig::default() (levenberg_marquardt.rs:319-359) or the
 * for_bundle_adjustment() preset (:519-530). */
void apex_lm_config_default(apex_lm_config* cfg);
void apex_lm_config_for_bundle_adjustment(apex_lm_config* cfg);

/* ncclGetUniqueId for the host to broadcast (128 bytes). */
apex_status apex_nccl_unique_id(void* out128);

apex_status apex_ctx_create(const apex_ctx_desc* desc, apex_ctx** out);
void apex_ctx_destroy(apex_ctx* ctx);
const char* apex_last_error(const apex_ctx* ctx);

/* Replaces Problem construction + initialize_optimization_state (src/optimizer/mod.rs:522-563):
 * copies the SoA problem to HBM, builds the point-major and camera-major observation orders,
 * normalises pose quaternions as SE3::from(DVector) does (se3.rs:200-206). */
apex_status apex_problem_upload(apex_ctx* ctx, const apex_problem_desc* desc);
apex_status apex_get_dims(const apex_ctx* ctx, apex_dims* out);
/* Overwrite / read back the current variable values (host arrays, full problem layout). */
apex_status apex_params_upload(apex_ctx* ctx, const double* pose, const double* intr, const double* pt);
apex_status apex_params_download(apex_ctx* ctx, double* pose, double* intr, double* pt);

/* AssemblyBackend::assemble (src/linearizer/mod.rs:191-227 -> cpu/sparse.rs:119-184) followed by the
 * block J^T J / J^T r accumulation that solve_augmented_equation starts with
 * (explicit_schur.rs:1146-1166): residuals, loss-corrected Jacobian blocks, H_cc/H_pp blocks, g.
 * Results stay on the device. `lambda` is the LM damping used for the point-block inverses. */
apex_status apex_linearize(apex_ctx* ctx, double lambda);
/* compute_residual_sparse + compute_cost (src/core/problem.rs:864-899, src/optimizer/mod.rs:358-361):
 * 0.5*||r~||^2 at the current values. */
apex_status apex_cost(apex_ctx* ctx, double* cost);

/* Debug / parity read-backs of the last apex_linearize, in the caller's observation order.
 * Any pointer may be NULL. r[nobs][2]; jc[nobs][2][dc] (columns: pose 6 then intrinsics K);
 * jp[nobs][2][3]. Single-rank contexts only. */
apex_status apex_get_linearization(apex_ctx* ctx, double* r, double* jc, double* jp);
/* hcc[ncam][dc][dc] (undamped), gc[ncam][dc], hpp[npts][9] (undamped), gp[npts][3],
 * hpp_inv[npts][9] (of the damped, guarded block: explicit_schur.rs:365-442 / implicit_schur.rs:685-778). */
apex_status apex_get_blocks(apex_ctx* ctx, double* hcc, double* gc, double* hpp, double* gp, double* hpp_inv);

/* apply_schur_operator_fast (implicit_schur.rs:163-251): y = (H_cc + lambda I - H_cp H_pp^-1 H_cp^T) x for
 * the last linearization. x, y are HOST vectors of ncam*dc doubles in camera-major order
 * [cam0: pose6, intrK | cam1: ...]. For parity tests. */
apex_status apex_schur_matvec(apex_ctx* ctx, const double* x, double* y);
/* Same operator on device-resident vectors, `reps` back-to-back launches, for roofline timing;
 * returns the mean CUDA-event duration of one application in milliseconds. */
apex_status apex_schur_matvec_bench(apex_ctx* ctx, int32_t reps, int32_t flush_l2, double* ms_per_call);

/* LinearSolver::solve_augmented_equation (src/linalg/mod.rs:143-180; explicit_schur.rs:1129-1234 /
 * implicit_schur.rs:1035-1085) on the last linearization. Outputs (host, may be NULL):
 * step_cam[ncam][dc], step_intr_unref[ncam][K] is implicit zero, step_pt[npts][3];
 * grad_norm = ||J^T r||_2 (get_gradient, +J^T r convention); pcg_iters. */
apex_status apex_solve_augmented(apex_ctx* ctx, int32_t schur_variant, int32_t preconditioner,
                                 int32_t cg_max_iterations, double cg_tolerance, double lambda,
                                 double* step_cam, double* step_pt, double* grad_norm, int32_t* pcg_iters);

/* Debug / parity read-back of the step the last solve produced (apex_solve_augmented, or the LAST iteration of
 * apex_lm_solve: the `step` of levenberg_marquardt.rs:888-899 before apply_parameter_step): step_cam[ncam][dc],
 * step_pt[npts][3] in the caller's landmark order. Either pointer may be NULL. Works on every rank of a sharded context
 * (collective: all ranks must call it). */
apex_status apex_get_step(apex_ctx* ctx, double* step_cam, double* step_pt);

/* LevenbergMarquardt::optimize (levenberg_marquardt.rs:1034-1083 -> optimize_with_mode :823-1028):
 * the whole loop, device resident. `trace` (may be NULL) receives up to trace_cap rows. */
apex_status apex_lm_solve(apex_ctx* ctx, const apex_lm_config* cfg, apex_lm_result* result,
                          apex_iter_trace* trace, int32_t trace_cap);

/* Observers: the reference's OptObserver (src/observers/mod.rs:201-330). The LM loop feeds them once per iteration, after the
 * accept / reject decision and before the convergence test, exactly like notify_observers_generic (src/optimizer/mod.rs:728-743,
 * called at levenberg_marquardt.rs:930-940): set_iteration_metrics(cost, gradient_norm, Some(damping), step_norm, Some(rho)) and
 * on_step(values, iteration) arrive here as ONE call of `on_step` with the metric tuple; the variable values of that moment are not
 * pushed (they live on the device) - an observer that wants them calls apex_params_download from inside the callback (the loop
 * is at a synchronisation point; on a sharded context that call is collective, so either every rank's observer downloads or
 * none). set_matrix_data is not fed on this path (the reference's generic path skips it as well, mod.rs:740-741).
 * `on_optimization_complete` is called once when the loop ends with a status, with iterations = the SolverResult's
 * (notify_complete, levenberg_marquardt.rs:1010-1011); not on an error return. Either pointer may be NULL. Observers stay
 * registered across solves and uploads until apex_clear_observers; callbacks run on the calling thread. */
typedef struct apex_observer_metrics {
  int32_t iteration;      /* index of the iteration just finished (0-based): the `iteration` of on_step */
  int32_t accepted;       /* step_eval.accepted */
  double cost;            /* state.current_cost after the decision */
  double gradient_norm;   /* step_result.gradient_norm */
  double damping;         /* Some(config.damping), already updated by update_damping */
  double step_norm;
  double step_quality;    /* Some(step_eval.rho) */
} apex_observer_metrics;
typedef struct apex_observer {
  void (*on_step)(void* user, apex_ctx* ctx, const apex_observer_metrics* metrics);
  void (*on_optimization_complete)(void* user, apex_ctx* ctx, int32_t iterations);
  void* user;
} apex_observer;
apex_status apex_add_observer(apex_ctx* ctx, const apex_observer* observer);
apex_status apex_clear_observers(apex_ctx* ctx);

/* Number of kernel launches issued by this context since creation (bench bookkeeping). */
int64_t apex_kernel_launches(const apex_ctx* ctx);

/* Device-side timing for bench.py (CUDA events on the context's own stream; the reference has no equivalent:
 * its only timers are the wall-clock IterationStats of src/optimizer/mod.rs:375-398). With profiling on, every
 * launch of the Schur-operator kernel inside a solve is bracketed by an event pair. apex_profile_read waits for
 * the stream, returns the totals since the last read and resets them. */
typedef struct apex_profile {
  double lm_device_ms;        /* first-to-last kernel of the last apex_lm_solve, on the device timeline   */
  double matvec_ms;           /* summed duration of the Schur-operator kernel launches                    */
  int64_t matvec_launches;
  double linearize_ms;        /* summed duration of apex_linearize's kernel group (K1+K2+K3)               */
  int64_t linearize_launches;
  double schur_form_ms;       /* explicit variants: formation of the dense reduced camera system S (K7)    */
  int64_t schur_forms;
  double cholesky_ms;         /* explicit variant: dense FP64 tensor-core factorisations (K8), summed      */
  int64_t cholesky_factorizations;
  uint64_t cholesky_n;        /* order of the factored matrix (padded)                                     */
  uint64_t upload_h2d_bytes;  /* host-to-device bytes of the last apex_problem_upload (structure + values) */
} apex_profile;
/* Measurement aid for K8: factor a synthetic n x n SPD matrix with the dense FP64 tensor-core Cholesky of the explicit
 * Schur path `reps` times (after one warm-up); average milliseconds of the factorisation alone. Needs no problem. */
apex_status apex_dense_cholesky_bench(apex_ctx* ctx, uint32_t n, int32_t reps, double* ms_per_factorization);
apex_status apex_profile_enable(apex_ctx* ctx, int32_t on);
apex_status apex_profile_read(apex_ctx* ctx, apex_profile* out);

/* Host-only: the sharding rule apex_problem_upload applies. Landmarks are owned block-cyclically: landmark p belongs to
 * rank (p / block) % nranks (so every rank sees every camera neighbourhood of a locality-ordered reconstruction); every
 * observation lives with its landmark. Returns the block size and what `rank` owns. Needs no device. */
apex_status apex_shard_info(uint32_t npts, uint64_t nobs, const uint32_t* obs_pt, int32_t nranks, int32_t rank,
                            uint32_t* block, uint32_t* npts_local, uint64_t* nobs_local);

/* Host-only: build the static observation layout apex_problem_upload would build for (nranks, rank) - landmark
 * shard, 256-slot point-major chunks with their camera-sorted lane order, camera-major work items - check its invariants
 * and report its size. Needs no device. */
typedef struct apex_layout_stats {
  uint32_t shard_block, npts_local; /* block-cyclic ownership: block size, landmarks owned by the rank */
  uint64_t nobs_local;
  uint32_t ntiles, nlong_tiles;    /* tiles; tiles holding one landmark with more than 256 observations */
  uint32_t nchunks, nnormal_chunks;
  uint32_t ncam_items, max_segments_per_chunk;
  uint64_t nsegments;              /* (chunk, camera) runs of the camera-sorted lanes, cut at warp boundaries */
  uint64_t slots_used;             /* = nobs_local when consistent                                 */
  int32_t consistent;              /* 1 when every structural invariant holds                      */
  int32_t reserved;
  double build_ms;
  /* work distribution of the Schur operator's chunk kernel: ranges (one per resident CTA of a 148-SM device), cameras per window,
   * windows, and the sum of the windows' camera counts (rows flushed per operator application) */
  uint32_t mv_ranges, mv_window, mv_nwindows, reserved2;
  uint64_t mv_rows;
} apex_layout_stats;
apex_status apex_layout_stats_compute(const apex_problem_desc* desc, int32_t nranks, int32_t rank, apex_layout_stats* out);

/* ---- BAL files and the CLI's problem construction (host only, no device needed) --------------------
 * apex_bal_load          <- BalLoader::load (crates/apex-io/src/bal.rs:138-400): blank lines skipped (:144-149),
 *                           header "ncam npts nobs", nobs lines "cam pt x y", 9 lines per camera (rx ry rz tx ty tz
 *                           f k1 k2), 3 lines per point; a focal length that is not positive and finite becomes
 *                           500.0 (:99-113). Errors map on IoError; apex_bal_last_error() holds the reference's
 *                           message text ("Parse error at line N: ...").
 * apex_bal_build_problem <- run_bundle_adjustment / add_factors (bin/bundle_adjustment.rs:212-441): axis-angle ->
 *                           quaternion (angle < 1e-10 -> identity, :200-208), pose storage [t, qw qx qy qz],
 *                           intrinsics [f,k1,k2], landmarks 0..num_points-1, observations with
 *                           point_index < num_points in file order, HuberLoss(1.0), pose_0000 fixed (6 DOF).
 *                           optimization_type: 0 = bundle-adjustment (pose + landmarks), 1 = self-calibration; the
 *                           CLI's other three types are not functional in the reference and give
 *                           APEX_ERR_UNSUPPORTED. The arrays behind the returned desc belong to the dataset object
 *                           and live until apex_bal_free / the next apex_bal_build_problem on it.
 * apex_bal_write         writes the same format (17 significant digits: values round-trip bit-exactly). */
typedef struct apex_bal_dataset apex_bal_dataset; /* opaque */
typedef struct apex_bal_view {
  uint32_t ncam, npts;
  uint64_t nobs;
  const double* cameras;   /* [ncam][9] rx ry rz tx ty tz f k1 k2 (f normalised)          */
  const double* points;    /* [npts][3]                                                   */
  const uint32_t* obs_cam; /* [nobs]                                                      */
  const uint32_t* obs_pt;  /* [nobs]                                                      */
  const double* obs_uv;    /* [nobs][2]                                                   */
} apex_bal_view;
apex_status apex_bal_load(const char* path, apex_bal_dataset** out);
apex_status apex_bal_from_arrays(uint32_t ncam, uint32_t npts, uint64_t nobs, const double* cameras, const double* points,
                                 const uint32_t* obs_cam, const uint32_t* obs_pt, const double* obs_uv, apex_bal_dataset** out);
apex_status apex_bal_view_get(const apex_bal_dataset* ds, apex_bal_view* out);
apex_status apex_bal_write(const apex_bal_dataset* ds, const char* path);
apex_status apex_bal_build_problem(apex_bal_dataset* ds, uint64_t num_points, int32_t optimization_type, apex_problem_desc* out);
void apex_bal_free(apex_bal_dataset* ds);
const char* apex_bal_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* APEX_GPU_H */
