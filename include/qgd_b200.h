/*
 * qgd_b200.h -- C ABI of the B200-native gradient hot path of QuantumGateDesign.jl.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types.
 * A Julia host binds these with `ccall` exactly the way the reference already binds
 * its only native dependency (`ccall((:bsplvd_, lib), Cvoid, (Ref{Float64}, ...))`,
 * reference src/Controls/FortranBSpline.jl:257-265).  INTEGRATION.md shows the
 * Julia-side stub for every entry point.
 *
 * Conventions (all follow the reference so a Julia `Array{Float64}` can be passed as is):
 *   - every array is column-major Float64 (or Int64 where stated), owned by the caller;
 *   - N = N_tot_levels, 2N = real_system_size, m = order/2, nic = N_initial_conditions,
 *     Nc = N_operators, P = total number of control coefficients;
 *   - the state is the real-split stack w = [u; v] (reference docs/src/index.md:35-48);
 *   - history arrays have the reference layout [2N, 1+m, 1+nsteps/save, nic]
 *     (reference src/forward_evolution.jl:23,43), batched evaluations append a
 *     trailing batch index;
 *   - every function returns 0 on success, a negative QGD_E* code otherwise, never
 *     unwinds across the boundary; the message is available from qgd_last_error().
 *   - one call in flight per handle (the reference's callers are single threaded,
 *     Ipopt invokes its callbacks serially, src/ipopt_optimal_control.jl:243-346).
 */
#ifndef QGD_B200_H
#define QGD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGD_OK 0
#define QGD_EINVAL (-1)      /* bad argument (reference: ArgumentError / @assert)        */
#define QGD_ECUDA (-2)       /* CUDA runtime failure, or no usable sm_100 device          */
#define QGD_ENOMEM (-3)      /* device allocation failed                                  */
#define QGD_EUNSUPPORTED (-4)/* valid in the reference but not built here */
#define QGD_ESTATE (-5)      /* call sequence error (e.g. history_precomputed w/o history) */

/* ---- operator matrices: reference SchrodingerProb{M} with M dense or SparseMatrixCSC
 *      (src/SchrodingerProb.jl:25-41) ------------------------------------------------ */
#define QGD_MAT_DENSE 0
#define QGD_MAT_CSC 1
typedef struct qgd_matrix {
  int32_t kind;          /* QGD_MAT_DENSE | QGD_MAT_CSC                                   */
  int32_t reserved;
  int64_t nrows, ncols;
  const double *dense;   /* [nrows*ncols] column-major, kind == QGD_MAT_DENSE             */
  int64_t nnz;           /* kind == QGD_MAT_CSC: Julia SparseMatrixCSC fields, 1-based    */
  const int64_t *colptr; /* [ncols+1]                                                     */
  const int64_t *rowval; /* [nnz]                                                         */
  const double *nzval;   /* [nnz]                                                         */
} qgd_matrix_t;

/* ---- controls on the hot path (reference src/Controls/) ----------- */
#define QGD_CONTROL_GRAPE 1           /* GRAPEControl(N_amplitudes, tf)      grape_control.jl:18-26   */
#define QGD_CONTROL_BSPLINE2 2        /* BSpline2Control(D1, tf)             bspline_control.jl:21-43 */
#define QGD_CONTROL_FORTRAN_BSPLINE 3 /* FortranBSplineControl(degree, N_basis, tf) FortranBSpline.jl:16-61 */
#define QGD_CONTROL_HOST_TABLE 4      /* any other AbstractControl (src/Controls/Control.jl:6-27: GeneralBSplineControl,
                                       * HermiteControl, HermiteCarrierControl, bcarrier2 ...): the host evaluates the
                                       * control protocol and passes tables (qgd_*_tables below); n_amplitudes = N_coeff */
typedef struct qgd_control {
  int32_t type;          /* base control type, QGD_CONTROL_*                              */
  int32_t reserved;
  double tf;
  int64_t n_amplitudes;  /* GRAPE                                                         */
  int64_t D1;            /* BSpline2                                                      */
  int64_t degree;        /* FortranBSpline                                                */
  int64_t n_basis;       /* FortranBSpline                                                */
  int64_t n_carriers;    /* 0: bare base control; >0: CarrierControl(base, freqs) CarrierControl.jl:5-23 */
  const double *carrier_freqs; /* [n_carriers]                                            */
} qgd_control_t;

#define QGD_PRECOND_IDENTITY 0 /* IdentityPreconditioner            preconditioners.jl:35-40  */
#define QGD_PRECOND_LU 1       /* LUPreconditioner                  preconditioners.jl:44-55  */
#define QGD_PRECOND_DIAGONAL 2 /* DiagonalHamiltonianPreconditioner preconditioners.jl:64-126 */

/* Mirror of the fields of the reference's SchrodingerProb (src/SchrodingerProb.jl:25-41)
 * plus the control collection that the reference passes alongside it. */
typedef struct qgd_problem {
  int64_t N_tot_levels;
  int64_t N_ess_levels;
  int64_t N_initial_conditions;
  int64_t N_operators;
  qgd_matrix_t system_sym;        /* K_s  [N,N]                                           */
  qgd_matrix_t system_asym;       /* S_s  [N,N]                                           */
  const qgd_matrix_t *sym_operators;  /* [Nc] K_c                                         */
  const qgd_matrix_t *asym_operators; /* [Nc] S_c                                         */
  const double *u0;               /* [N, nic]                                             */
  const double *v0;               /* [N, nic]                                             */
  qgd_matrix_t guard_subspace_projector; /* [2N,2N]; nnz==0 / all-zero allowed            */
  double tf;
  int64_t nsteps;
  double gmres_abstol;
  double gmres_reltol;
  int32_t preconditioner;         /* QGD_PRECOND_*                                        */
  int32_t reserved;
  const qgd_control_t *controls;  /* [Nc], control k drives operator k                    */
} qgd_problem_t;

typedef struct qgd_handle qgd_handle_t;

/* Number of coefficients of one control / of the whole collection
 * (reference get_number_of_control_parameters, src/Controls/Control.jl:94-96). */
int64_t qgd_control_n_coeff(const qgd_control_t *c);
int64_t qgd_problem_n_coeff(const qgd_problem_t *p);

/* Build the device-resident problem: validates like the reference's inner constructor
 * (symmetry / antisymmetry / shapes, src/SchrodingerProb.jl:73-154), converts the
 * operators to the on-chip row format, builds the preconditioner factors. `device` is the
 * CUDA ordinal (-1: current device). Fails with QGD_ECUDA when no sm_100 GPU is present;
 * there is no CPU fallback. */
int qgd_create(const qgd_problem_t *prob, int device, qgd_handle_t **out);
int qgd_destroy(qgd_handle_t *h);
const char *qgd_last_error(void);

/* The reference mutates prob.nsteps / prob.gmres_abstol / prob.gmres_reltol in place
 * (examples/cnot3_optimize_gate.jl:51-55); the same three knobs here. */
int qgd_set_nsteps(qgd_handle_t *h, int64_t nsteps);
int qgd_set_gmres_tolerances(qgd_handle_t *h, double abstol, double reltol);

/* Behaviour switches of a handle (they replace the environment variables of the first round; nothing in the library
 * reads the environment).  Unknown keys return QGD_EINVAL.  Changing an option invalidates the resident history. */
#define QGD_OPT_STRICT_MGS 1            /* 1: strict modified Gram-Schmidt in the register-operator sweeps, the
                                         * orthogonalisation of IterativeSolvers' ModifiedGramSchmidt one projection at a
                                         * time (iteration counts equal the reference algorithm's, tests assert equality);
                                         * 0 (default): coefficients of 32 basis vectors at a time from the same vector
                                         * (classical inside a block, modified across blocks)      */
#define QGD_OPT_DISABLE_FAST 2          /* 1: never take the register-operator sweeps (qgd_fast.cuh)                    */
#define QGD_OPT_DISABLE_DENSE_SWEEP 3   /* 1: never take the FP64 tensor-core sweeps (qgd_dense.cu)                     */
#define QGD_OPT_DISABLE_DENSE_DMMA 4    /* 1: qgd_compute_derivatives never takes the tensor-core contraction           */
#define QGD_OPT_DENSE_TERMINAL 5        /* dense terminal condition: 0 sequential columns with the reference's carried
                                         * initial guess (default), 1 all columns in parallel from a zero guess,
                                         * 2 the generic one-warp kernel                                                */
#define QGD_OPT_DISABLE_TMEM 6          /* 1: no tensor-memory tier for the Krylov basis                                */
#define QGD_OPT_SEG_STEPS 7             /* time steps per work-queue ticket (0: automatic, about nsteps / 24)           */
#define QGD_OPT_L2_PERSIST 8            /* 1: persisting L2 access-policy window over the Krylov workspace              */
#define QGD_OPT_LATENCY_WARPS 9         /* warps per CTA of the register-operator sweeps (0: automatic)                 */
#define QGD_OPT_TERMINAL_EXCHANGE 10    /* column sharding over GPUs, exchange between the sweeps: 0 (default) the final
                                         * states of all columns (2N*nic doubles per control vector), every rank then
                                         * solves the terminal condition of all columns in the reference's order with its
                                         * carried initial guess -- results equal the single-GPU ones; 1 only the two
                                         * scalars dot(psi,R), dot(psi,T) per control vector, every rank solves its own
                                         * columns (the first one from a zero guess: lambda_N moves by the GMRES
                                         * tolerance; sparse problems only)                                             */
#define QGD_OPT_LATENCY_TEAM 11         /* the four-warps-per-column latency team of the register-operator sweeps: 0 (default)
                                         * automatic -- taken when no more columns are in flight than the GPU has SMs, e.g.
                                         * ONE gradient evaluation --, 1 always, 2 never                                 */
#define QGD_OPT_MAX 11
int qgd_set_option(qgd_handle_t *h, int32_t key, int64_t value);
int qgd_get_option(qgd_handle_t *h, int32_t key, int64_t *value);

/* Restrict this handle to the initial-condition columns [col_begin, col_begin+col_count)
 * (multi-GPU column sharding: one process per GPU each owning a contiguous block, the
 * GPU counterpart of `Threads.@threads for initial_condition_index`,
 * src/forward_evolution.jl:48,332).  Default: all columns. */
int qgd_set_column_shard(qgd_handle_t *h, int64_t col_begin, int64_t col_count);

/* eval_forward! (src/forward_evolution.jl:33-70, 88-245) for n_batch control vectors.
 *   pcof        [P, n_batch]
 *   history     [2N, 1+m, 1+nsteps/save_every, ncol, n_batch]  or NULL (kept on device only)
 *   final_state [2N, ncol, n_batch] or NULL            (history[:,1,end,:])
 *   gmres_iters [nsteps, ncol, n_batch] Int64 or NULL  (loop count of :211-217 per step)
 * ncol = columns owned by this handle. */
int qgd_eval_forward(qgd_handle_t *h, const double *pcof, int64_t n_batch, int32_t order,
                     int64_t save_every, double *history, double *final_state,
                     int64_t *gmres_iters);

/* The same split in two, so that several handles (each owns a stream) run their sweeps CONCURRENTLY on one GPU: the
 * step-size studies of the reference (get_histories, src/Tests/test_convergence.jl:76-121; estimate_timesteps_per_period,
 * src/calculate_timestep.jl:58-98) solve the same problem at nsteps = base * 2^k, k = 0..5 -- six launches of 8 warps each,
 * which fill a B200 only when they overlap.  _async uploads pcof and enqueues; _collect synchronises and copies out. */
int qgd_eval_forward_async(qgd_handle_t *h, const double *pcof, int64_t n_batch, int32_t order,
                           int64_t save_every, int32_t want_iters);
int qgd_eval_forward_collect(qgd_handle_t *h, double *history, double *final_state, int64_t *gmres_iters);

/* eval_forward! with the `forcing` keyword (src/forward_evolution.jl:33-37, 118-129, 167-206): the forcing enters the
 * Taylor recursion at t_n (compute_derivatives!(...; forcing_matrix), src/hermite.jl:91-95) and, being explicit, the
 * implicit-side combination of the forcing at t_{n+1} is moved to the right-hand side; the Taylor columns stored for
 * the final time are computed WITHOUT forcing, as the reference does (:232-236).
 *   forcing [2N, m, 1+nsteps, ncol, n_batch]; other arguments as qgd_eval_forward. */
int qgd_eval_forward_forced(qgd_handle_t *h, const double *pcof, int64_t n_batch, int32_t order,
                            int64_t save_every, const double *forcing, double *history,
                            double *final_state, int64_t *gmres_iters);

/* eval_grad_forced (src/eval_grad_forced.jl:18-195), cost_type = :Infidelity: the gradient by differentiating
 * Schroedinger's equation w.r.t. each control parameter -- the reference's exactness cross-check of the
 * discrete adjoint (test/GradientTests/compare_gradients.jl:47-66).  One unforced solve, then P forced solves
 * with zero initial state batched on the device (their forcing is formed on the fly from the resident
 * history and the control basis table; no [2N, m, 1+nsteps, nic] forcing array per parameter exists).
 *   pcof [P] (one control vector), target [2N, nic], grad [P]. */
int qgd_eval_grad_forced(qgd_handle_t *h, const double *pcof, const double *target, int32_t order,
                         double *grad);

/* Host-evaluated controls: the sweeps only consume p_k^(j)(t_n)/j!, q_k^(j)(t_n)/j! and, for the gradient,
 * d/dtheta of the same (what fill_p_mat!/fill_q_mat!, src/Controls/Control.jl:99-149, and eval_grad_p_derivative! /
 * eval_grad_q_derivative! return, divided by j!).  For control types the device kernels do not evaluate, or when any
 * control is QGD_CONTROL_HOST_TABLE, the caller fills these tables with the reference's own control code:
 *   cvals [Nc, 1+m, 2, 1+nsteps, n_batch]   cvals[k, j, 0, n, b] = p_k^(j)(t_n)/j!, [.., 1, ..] = q_k^(j)(t_n)/j!
 *   table [P,  1+m, 2, 1+nsteps]            table[theta, j, 0, n] = d/dtheta p_k(theta)^(j)(t_n)/j!  (shared by the
 *                                           batch: controls linear in pcof; nonlinear controls use n_batch = 1)
 * (column-major, first index fastest, t_n = n tf/nsteps).  Outputs as qgd_eval_forward / qgd_discrete_adjoint. */
int qgd_eval_forward_tables(qgd_handle_t *h, int64_t n_batch, int32_t order, int64_t save_every,
                            const double *cvals, double *history, double *final_state,
                            int64_t *gmres_iters);
int qgd_discrete_adjoint_tables(qgd_handle_t *h, int64_t n_batch, int32_t order, const double *cvals,
                                const double *table, const double *target, double *grad,
                                double *infidelity, double *guard_penalty);

/* discrete_adjoint! (src/eval_grad_discrete_adjoint.jl:107-160), cost_type = :Infidelity.
 *   target      [2N, nic] real-stacked vcat(real, imag) of the complex gate (what
 *               complex_to_real produces at :126); ALL nic columns even when sharded
 *   history_precomputed != 0: reuse the history left on the device by the previous
 *               qgd_eval_forward / qgd_discrete_adjoint call of this handle (same pcof batch)
 *   grad        [P, n_batch]                         (sum over this handle's columns)
 *   infidelity  [n_batch]  infidelity_real (src/infidelity.jl:7-18)   (all columns needed:
 *               in sharded mode use the two-phase API below)
 *   guard_penalty [n_batch] guard_penalty_real (src/infidelity.jl:56-96)
 *   history, lambda_history [2N,1+m,1+nsteps,ncol,n_batch] or NULL
 *   adjoint_forcing [2N,1+nsteps,ncol,n_batch] or NULL
 *   iters_fwd, iters_adj [nsteps, ncol, n_batch] or NULL; iters_term [nic, n_batch] or NULL */
int qgd_discrete_adjoint(qgd_handle_t *h, const double *pcof, int64_t n_batch,
                         const double *target, int32_t order, int32_t history_precomputed,
                         double *grad, double *infidelity, double *guard_penalty,
                         double *history, double *lambda_history, double *adjoint_forcing,
                         int64_t *iters_fwd, int64_t *iters_adj, int64_t *iters_term);

/* Same evaluation with every buffer already in device memory (HBM) and nothing copied:
 * d_pcof [P,n_batch], d_target [2N,nic], d_grad [P,n_batch], d_infidelity/d_guard [n_batch].
 * Launches on `stream` (a cudaStream_t passed as void*, NULL = the handle's own stream) and
 * returns without synchronising. */
int qgd_discrete_adjoint_device(qgd_handle_t *h, const double *d_pcof, int64_t n_batch,
                                const double *d_target, int32_t order, double *d_grad,
                                double *d_infidelity, double *d_guard_penalty, void *stream);

/* Synchronise the handle's stream (and `stream`, if not NULL) and report the device error word of the sweeps: the way
 * to complete and check an asynchronous qgd_discrete_adjoint_device call. */
int qgd_synchronize(qgd_handle_t *h, void *stream);

/* ---- multi-GPU inside the library ------------------------------------------------------------------------------------
 * The reference parallelises over the independent initial-condition columns (Threads.@threads,
 * src/forward_evolution.jl:48,332); they couple only through dot(final_state, R), dot(final_state, T) of
 * compute_terminal_condition (src/eval_grad_discrete_adjoint.jl:27-28) and the serial gradient sum (:150-157).  With a
 * communicator attached, qgd_discrete_adjoint / qgd_discrete_adjoint_device on every participating handle evaluate
 * the SAME control vectors on their own block of columns and exchange on the device, on the sweep stream: one NCCL
 * all-reduce between the sweeps (QGD_OPT_TERMINAL_EXCHANGE) and one all-reduce of [grad; guard] at the end; every rank
 * returns the complete gradient, infidelity and guard penalty.  NCCL is bound at run time (libnccl.so.2).
 *
 * (a) one process per GPU: rank 0 calls qgd_comm_get_unique_id, the host program distributes the 128 bytes (MPI,
 *     torch.distributed, a file ...), every rank calls qgd_comm_init_rank on its handle (collective), which also
 *     assigns the rank its contiguous column block [rank*nic/n, (rank+1)*nic/n). */
#define QGD_NCCL_UNIQUE_ID_BYTES 128
int qgd_comm_set_nccl_library(const char *path); /* optional: path of libnccl.so.2 (default: the loader's search)      */
int qgd_comm_get_unique_id(unsigned char *id128);
int qgd_comm_init_rank(qgd_handle_t *h, int32_t n_ranks, int32_t rank, const unsigned char *id128);
int qgd_comm_finalize(qgd_handle_t *h);           /* back to a single-GPU handle owning all columns                     */

/* (b) one process driving n GPUs (what a Julia session does: `qgd_init_multi_gpu` of SURVEY 8b).  The set owns one
 *     handle per device and an NCCL communicator over them (ncclCommInitAll).
 *     shard = QGD_SHARD_COLUMNS: every GPU takes a block of columns of all n_batch control vectors (the exchanges
 *     above, grouped); QGD_SHARD_CONTROL_VECTORS: every GPU takes a block of the control vectors with all columns, no
 *     collective (the batched random-pcof sweep).  Host buffers in and out like qgd_discrete_adjoint. */
typedef struct qgd_multi qgd_multi_t;
#define QGD_SHARD_COLUMNS 0
#define QGD_SHARD_CONTROL_VECTORS 1
int qgd_init_multi_gpu(const qgd_problem_t *prob, int32_t n_gpus, const int32_t *devices /* NULL: 0..n_gpus-1 */,
                       qgd_multi_t **out);
int qgd_multi_n_gpus(qgd_multi_t *mg);
int qgd_multi_handle(qgd_multi_t *mg, int32_t i, qgd_handle_t **out); /* per-device handle, e.g. for qgd_set_option */
int qgd_multi_set_nsteps(qgd_multi_t *mg, int64_t nsteps);
int qgd_multi_set_gmres_tolerances(qgd_multi_t *mg, double abstol, double reltol);
int qgd_multi_discrete_adjoint(qgd_multi_t *mg, const double *pcof, int64_t n_batch, const double *target,
                               int32_t order, int32_t shard, double *grad, double *infidelity,
                               double *guard_penalty);
int qgd_multi_destroy(qgd_multi_t *mg);

/* Two-phase form for column sharding across GPUs: phase 1 runs the forward
 * sweep on the owned columns and returns their final states and guard-penalty partial;
 * the host all-gathers the final states (the only cross-column coupling: <psi,R>, <psi,T>
 * of compute_terminal_condition, src/eval_grad_discrete_adjoint.jl:27-28); phase 2 takes
 * the final states of ALL columns and produces this rank's gradient partial, to be
 * all-reduced by the host.
 *   final_state_local [2N, ncol, n_batch], guard_local [n_batch]
 *   final_state_all   [2N, nic, n_batch] */
int qgd_adjoint_phase1(qgd_handle_t *h, const double *pcof, int64_t n_batch, int32_t order,
                       double *final_state_local, double *guard_local);
int qgd_adjoint_phase2(qgd_handle_t *h, const double *target, const double *final_state_all,
                       double *grad_local, double *infidelity);

/* Objective pieces on their own (src/infidelity.jl:7-18, 56-96). final_state [2N,nic,n_batch]. */
int qgd_infidelity_real(qgd_handle_t *h, const double *final_state, const double *target,
                        int64_t n_batch, double *infidelity);

/* Control evaluation, kernel K1 (fill_p_mat!/fill_q_mat!, src/Controls/Control.jl:125-149):
 *   times [ntimes]; pcof [P]; p_out,q_out [nderiv, Nc, ntimes] with entry (1+j,k,it) =
 *   p_k^(j)(t)/j!  -- the Taylor-scaled values the time stepper consumes.
 * grad tables (eval_grad_{p,q}_derivative!, e.g. FortranBSpline.jl:149-189), un-scaled:
 *   gp_out,gq_out [P, nderiv, ntimes] with entry (theta, 1+r, it) = d p_k(theta)^(r)(t) / d theta
 *   where k(theta) is the control owning coefficient theta; either may be NULL. */
int qgd_eval_controls(qgd_handle_t *h, const double *pcof, const double *times, int64_t ntimes,
                      int32_t nderiv, double *p_out, double *q_out, double *gp_out,
                      double *gq_out);

/* Kernel K2 on its own (compute_derivatives!, src/hermite.jl:56-101 and the transposed
 * recursion W_j(t)^T x of compute_adjoint_derivatives!, :284-305):
 *   uv [2N, 1+m, ncols_in]: column 1 given, columns 2..1+m overwritten;
 *   cvals_re, cvals_im [1+m, Nc]; adjoint != 0 selects Lambda_j = W_j^T x. */
int qgd_compute_derivatives(qgd_handle_t *h, double *uv, int64_t ncols_in, int32_t order,
                            const double *cvals_re, const double *cvals_im, int32_t adjoint);

/* Counters of the last call, for tracing: kernel launches, bytes copied each way. */
typedef struct qgd_stats {
  int64_t kernel_launches;
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  double last_forward_ms;  /* CUDA-event time of the forward sweep kernel  */
  double last_backward_ms; /* CUDA-event time of the backward sweep kernel */
  double last_total_ms;    /* CUDA-event time of the whole device section  */
  int64_t fast_path_launches; /* sweep launches that took the register-operator kernels (qgd_fast.cuh) */
  int64_t collectives;        /* NCCL collectives enqueued by the call (multi-GPU)                      */
} qgd_stats_t;
int qgd_get_stats(qgd_handle_t *h, qgd_stats_t *out);

/* Measured FP64 FMA throughput of this device in TFLOP/s (micro-benchmark kernel); the
 * roofline denominator MEASURED_PEAKS.json does not provide. */
int qgd_measure_fp64_peak(int device, double *tflops);
/* Same for the FP64 tensor cores (DMMA.8x8x4): the denominator for the dense sweeps. */
int qgd_measure_dmma_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* QGD_B200_H */
