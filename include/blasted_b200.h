/** \file blasted_b200.h
 * \brief C ABI of the B200-native asynchronous-preconditioner path (libblasted_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Every entry point
 * names the reference interface it stands in for (paths relative to the BLASTed source tree).
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md and implemented in
 * blasted_b200/host/b200_solverops.hpp (C++ adapters deriving from the reference's own
 * SRPreconditioner / FactoryBase classes).
 *
 * Conventions
 *  - all functions return 0 on success, non-zero on failure; b200_last_error() gives the message
 *    (the C++ adapters re-throw the reference's exception types; nothing throws across this ABI);
 *  - matrices are square sparse (block-)row matrices in the reference layout
 *    (include/srmatrixdefs.hpp:38-79): browptr[nbrows+1], bcolind[nnzb] sorted ascending per row,
 *    vals[nnzb*bs*bs] with contiguous bs x bs blocks, diagind[nbrows]; int32 indices, fp64 values;
 *  - `*_host` variants take host buffers and perform the H2D/D2H copies themselves (what PETSc's
 *    PCApply hands over, src/blasted_petsc.cpp:474-517); un-suffixed compute-path variants taking
 *    `d_` arguments expect device pointers and are asynchronous on the handle's CUDA stream;
 *  - one in-flight operation per handle (the reference objects are not re-entrant either:
 *    `apply` mutates the shared scratch `ytemp`, src/solverops_ilu0.cpp:280).
 *  - There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef BLASTED_B200_H
#define BLASTED_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums: numeric values identical to the reference ---- */

/** include/solvertypes.h:14-26 (BlastedSolverType) */
typedef enum {
	B200_JACOBI = 0,
	B200_GS = 1,
	B200_SGS = 2,
	B200_ILU0 = 3,
	B200_SEQILU0 = 4,
	B200_SFILU0 = 5,
	B200_SAPILU0 = 6,
	B200_CSC_BGS = 7,          /* not provided (out of scope: solverops_sgs.hpp:116-119) */
	B200_LEVEL_SGS = 8,
	B200_ASYNC_LEVEL_ILU0 = 9,
	B200_NO_PREC = 10,
	B200_EXTERNAL = 11         /* not provided */
} b200_solver_type;

/** include/async_initialization_decl.hpp:16-25 (FactInit) */
typedef enum { B200_INIT_F_ZERO = 0, B200_INIT_F_ORIGINAL = 1, B200_INIT_F_SGS = 2,
               B200_INIT_F_NONE = 3 } b200_fact_init;
/** include/async_initialization_decl.hpp:28-35 (ApplyInit) */
typedef enum { B200_INIT_A_ZERO = 0, B200_INIT_A_JACOBI = 1, B200_INIT_A_NONE = 2 } b200_apply_init;

/** Block layout: Eigen::ColMajor = 0, Eigen::RowMajor = 1 (include/blasted_config.hpp:13-14) */
typedef enum { B200_COLMAJOR = 0, B200_ROWMAJOR = 1 } b200_block_storage;

/** How the level-scheduled objects order rows (device-side choice; no reference counterpart).
 *  CONTIGUOUS reproduces src/levelschedule.cpp:12-71 bit-for-bit (levels of consecutive rows);
 *  DAG uses true dependency wavefronts (same numerical result: exact substitution). */
typedef enum { B200_LEVELS_DAG = 0, B200_LEVELS_CONTIGUOUS = 1 } b200_level_mode;

/** Mirrors AsyncSolverSettings : SolverSettings (include/solverfactory.hpp:46-68). */
typedef struct b200_settings {
	int prectype;             /* b200_solver_type */
	int bs;                   /* block size: 1, 4 or 5 (preconditioners); SpMV also 3 and 7 */
	int blockstorage;         /* b200_block_storage */
	int relax;                /* request relaxation instead of preconditioning */
	int thread_chunk_size;    /* accepted for interface parity and IGNORED: the OpenMP dynamic-schedule
	                           * chunk of the reference (solverfactory.hpp:55) has no device counterpart -
	                           * work is mapped by persistent grids sized from the SM count */
	int scale;                /* symmetric scaling of the matrix before ILU */
	int nbuildsweeps;         /* asynchronous factorisation sweeps */
	int napplysweeps;         /* asynchronous triangular-solve / SGS sweeps */
	int fact_inittype;        /* b200_fact_init */
	int apply_inittype;       /* b200_apply_init */
	int compute_precinfo;     /* fill the 6-entry PrecInfo in b200_prec_compute */
	int level_mode;           /* b200_level_mode, for the level-scheduled types */
} b200_settings;

typedef struct b200_mat b200_mat;       /* device-resident CSR/BSR matrix (operator A) */
typedef struct b200_prec b200_prec;     /* preconditioner object */

/* ---- library ---- */
const char *b200_last_error(void);
/** Number of CUDA devices visible, 0 if none (never fails). */
int b200_device_count(void);
/** Select the device used by handles created afterwards on this thread (cudaSetDevice). */
int b200_set_device(int device);
/** Number of kernels launched by this library since load / the last reset (bench evidence). */
long long b200_kernel_launches(void);
void b200_reset_kernel_launches(void);

/** Per-kernel-class device timing (CUDA events on the launching stream), off by default.
 *  Classes: 0 factor lower launch, 1 factor upper launch, 2 factor init, 3 diagonal inversion /
 *  scatter, 4 triangular/SGS lower (forward) sweep, 5 upper (backward) sweep, 6 SpMV, 7 other.
 *  Device-side counterpart of the hand timers in src/blasted_petsc.cpp:416-427,499-510. */
void b200_profile_enable(int on);
void b200_profile_reset(void);
int b200_profile_get(double ms[8], long long count[8]);
/* All classes: the eight above, then 8 BLAS-1 of the Krylov drivers (dots, axpys), 9 halo pack,
 * 10 time the compute stream waited for the halo, 11 all-reduce of the dot products (includes the
 * wait for the slowest rank).  n <= b200_profile_classes(). */
int b200_profile_classes(void);
int b200_profile_get_n(int n, double *ms, long long *count);

/* ---- device-resident matrix: SRMatrixStorage + CSRMatrixView/BSRMatrixView
 *      (include/srmatrixdefs.hpp:38-79, include/blockmatrices.hpp:71-160) ---- */

/** Copies a host matrix to the device.  Stands in for wrapping raw arrays in
 *  SRMatrixStorage<const double,const int> (src/blasted_petsc.cpp:285-297) plus
 *  CSRMatrixView/BSRMatrixView construction (tests/testsolve.cpp:43-56).
 *  diagind may be NULL (it is then located on the device); a missing diagonal is only an error
 *  for preconditioners, not for SpMV. */
int b200_mat_create_host(int nbrows, int bs, int blockstorage, const int *browptr,
                         const int *bcolind, const double *vals, const int *diagind,
                         b200_mat **out);
/** Same, from device arrays that are copied (pattern) / transposed-or-copied (values). */
int b200_mat_create_device(int nbrows, int bs, int blockstorage, const int *d_browptr,
                           const int *d_bcolind, const double *d_vals, b200_mat **out);
/** New values on the same pattern (PETSc rewrites `a` in place between solves,
 *  include/solverops_ilu0.hpp:53-56). */
int b200_mat_update_values_host(b200_mat *m, const double *vals);
int b200_mat_update_values_device(b200_mat *m, const double *d_vals);
void b200_mat_destroy(b200_mat *m);
int b200_mat_dim(const b200_mat *m);            /* nbrows*bs, as MatrixView::dim() */
int b200_mat_nbrows(const b200_mat *m);
long long b200_mat_nnzb(const b200_mat *m);
int b200_mat_set_stream(b200_mat *m, void *cuda_stream);
/** The restarted Krylov drivers keep their basis storage with the operator between solves
 *  (up to 2*restart+2 vectors, allocated eight at a time as the basis grows); this frees it. */
int b200_mat_release_workspace(b200_mat *m);

/* ---- front end: the step before the path.  Every array argument is
 *      a host pointer when on_device == 0 and a device pointer otherwise. ---- */

/** Coordinate triplets (0-based, any order, square matrix of `nrows` scalar rows) to a resident
 *  CSR (bs = 1) or BSR matrix with sorted columns, located diagonals and zero-filled blocks.
 *  Replaces COOMatrix's row/column sort and convertToCSR / convertToBSR<bs,stor>
 *  (src/coomatrix.cpp:222-403); `blockstorage` is the layout b200_mat_get_host hands back. */
int b200_mat_create_coo(int nrows, long long nnz, const int *rowind, const int *colind,
                        const double *vals, int bs, int blockstorage, int on_device, b200_mat **out);
/** Copies the matrix back to host arrays (any may be NULL): browptr[nbrows+1], bcolind[nnzb],
 *  diagind[nbrows], vals[nnzb*bs*bs] in the caller's block layout. */
int b200_mat_get_host(const b200_mat *m, int *browptr, int *bcolind, int *diagind, double *vals);
/** Reordering::applyOrdering on the resident matrix (src/reorderingscaling.cpp:77-205):
 *  forward - new row i is old row rord[i], columns renamed by the inverse of cord and re-sorted;
 *  inverse - old row i moves to row rord[i], columns renamed by cord.  Either may be NULL.
 *  diagind is located again afterwards (the reference leaves it stale).  Preconditioners built on
 *  the matrix before the call must be rebuilt. */
int b200_mat_reorder(b200_mat *m, const int *rord, const int *cord, int inverse, int on_device);
/** ReorderingScaling::applyScaling on the resident matrix (src/reorderingscaling.cpp:282-337):
 *  block row i times (inverse: divided by) rowscale[i], then block column j by colscale[j]. */
int b200_mat_scale(b200_mat *m, const double *rowscale, const double *colscale, int inverse,
                   int on_device);
/** Reordering::applyOrdering on a vector of n blocks of bs (src/reorderingscaling.cpp:211-266):
 *  forward vec[i] <- vec[ord[i]], inverse vec[ord[i]] <- vec[i]; ord is the row or the column
 *  ordering according to the direction wanted.  ord == NULL is a no-op. */
int b200_vec_reorder(double *vec, long long n, int bs, const int *ord, int inverse, int on_device);
/** ReorderingScaling::applyScaling on a vector (src/reorderingscaling.cpp:340-368). */
int b200_vec_scale(double *vec, long long n, int bs, const double *scale, int inverse, int on_device);

/** y = A x  : AbstractLinearOperator::apply (include/linearoperator.hpp:36),
 *  BLAS_CSR/BLAS_BSR::matrix_apply (src/blas/matvecs.cpp:25-48, 78-91). */
int b200_mat_apply(const b200_mat *m, const double *d_x, double *d_y);
int b200_mat_apply_host(const b200_mat *m, const double *x, double *y);
/** z = a A x + b y : MatrixView::gemv3 (include/linearoperator.hpp:125-131, matvecs.cpp:51-108). */
int b200_mat_gemv3(const b200_mat *m, double a, const double *d_x, double b, const double *d_y,
                   double *d_z);
int b200_mat_gemv3_host(const b200_mat *m, double a, const double *x, double b, const double *y,
                        double *z);

/* ---- preconditioners: SRFactory::create_preconditioner + Preconditioner virtuals
 *      (src/solverfactory.cpp:131-228, include/solverops_base.hpp:32-64) ---- */

/** Creates the preconditioner of settings->prectype on the matrix `m` (shared, not copied; `m`
 *  must outlive the preconditioner).  Fails for invalid (prectype, bs, blockstorage) combinations
 *  exactly where the reference throws std::invalid_argument (src/solverfactory.cpp:122, 203-206,
 *  220-227); b200_last_error() carries the reference's message. */
int b200_prec_create(const b200_settings *settings, b200_mat *m, b200_prec **out);
/** Preconditioner::compute(): (re)build from the matrix's current values; first call also builds
 *  the ILU position lists / level schedule on the device (src/solverops_ilu0.cpp:190-202,358-368).
 *  precinfo may be NULL; layout = PrecInfo::f_info (include/preconditioner_diagnostics.hpp:14-58). */
int b200_prec_compute(b200_prec *p, double precinfo[6]);
/** The update a Newton / time-stepping loop performs between solves, in one call: new values of
 *  the matrix from host memory (same pattern - PETSc rewrites `a` in place between compute()
 *  calls, include/solverops_ilu0.hpp:53-56, src/blasted_petsc.cpp:314-327) followed by compute().
 *  The values travel in chunks on a copy stream; the layout conversion and, where it depends on
 *  nothing but the entry itself (ILU0 types on blocks, INIT_F_ORIGINAL, no scaling), the initial
 *  guess of the factor run behind the next chunk's copy.  Returns when `vals` has been read;
 *  the factorisation itself is only enqueued.  vals == NULL: plain b200_prec_compute. */
int b200_prec_compute_host(b200_prec *p, const double *vals, double precinfo[6]);
/** Preconditioner::apply(r, z) : z = M^-1 r. */
int b200_prec_apply(b200_prec *p, const double *d_r, double *d_z);
int b200_prec_apply_host(b200_prec *p, const double *r, double *z);
/** setApplyParams({.., maxits}) + apply_relax(b, x) (src/blasted_petsc.cpp:519-576): x is
 *  updated in place.  Fails with the reference's message where it throws
 *  ("ILU relaxation not implemented!", src/solverops_ilu0.cpp:215). */
/** Preconditioner::setApplyParams(SolveParams{rtol, atol, dtol, ctol, maxits})
 *  (include/solverops_base.hpp:18-26, :58).  With ctol != 0 the Jacobi relaxation stops on
 *  ||x_new - x||_2 < atol, relative decrease < rtol or relative growth > dtol
 *  (src/solverops_jacobi.cpp:86-105); the asynchronous relaxations never check (as the reference).
 *  apply_relax with maxits <= 0 uses the maxits stored here. */
int b200_prec_set_apply_params(b200_prec *p, double rtol, double atol, double dtol, int ctol, int maxits);
int b200_prec_apply_relax(b200_prec *p, const double *d_b, double *d_x, int maxits);
int b200_prec_apply_relax_host(b200_prec *p, const double *b, double *x, int maxits);
/** Deferred errors of the asynchronous device-pointer entry points: synchronises the handle's
 *  stream and fails (once) if a one-launch exact substitution (seqilu0 / sapilu0 / level types)
 *  met a dependency that never arrived since the last check.  The *_host entry points and the
 *  Krylov drivers call it themselves; callers of b200_prec_apply() call it where they next
 *  synchronise (the reference throws synchronously from apply(); there is no asynchronous
 *  counterpart in include/solverops_base.hpp). */
int b200_prec_check(b200_prec *p);
int b200_prec_dim(const b200_prec *p);
int b200_prec_relaxation_available(const b200_prec *p);
void b200_prec_destroy(b200_prec *p);
int b200_prec_set_stream(b200_prec *p, void *cuda_stream);
/** Change sweep counts after creation (bench / convergence studies). */
int b200_prec_set_sweeps(b200_prec *p, int nbuildsweeps, int napplysweeps);

/* ---- setup products and diagnostics, for parity checks ---- */

/** ILUPositions (include/ilu_pattern.hpp:36-48) built on the device: sizes, then copy-out. */
int b200_prec_positions_size(b200_prec *p, long long *npos);
int b200_prec_get_positions(b200_prec *p, int *posptr, int *lowerp, int *upperp);
/** Sizes of the work lists built from the position lists, without copying them out: stats =
 *  {lower entries, upper entries incl. diagonals, upper entries that change between sweeps (have
 *  products or are diagonal), products of lower entries, products of upper entries} - what the
 *  algorithmic byte counts of the factor launches are made of (bench.py). */
int b200_prec_pattern_stats(b200_prec *p, long long stats[5]);
/** Level schedule.  CONTIGUOUS mode: `levels` (nlevels+1 entries) as computeLevels returns.
 *  DAG mode: level pointers (nlevels+1) and the row list ordered by level (nbrows). */
int b200_prec_levels_size(b200_prec *p, int *nlevels);
int b200_prec_get_levels(b200_prec *p, int *level_ptr, int *level_rows_or_null);
/** Factor values in the caller's block layout; diagonal blocks inverted for bs > 1, exactly what
 *  the reference holds in `iluvals` after compute() (src/async_blockilu_factor.cpp:144-146). */
int b200_prec_get_factor(b200_prec *p, double *iluvals);
/** Inverted diagonal (blocks) of the Jacobi-derived objects (`dblocks`, solverops_jacobi.cpp:31). */
int b200_prec_get_dblocks(b200_prec *p, double *dblocks);
/** Symmetric scaling vector (src/rawsrmatrixutils.cpp:343-350), if scaling is on. */
int b200_prec_get_scale(b200_prec *p, double *scale);
/** sum |(A - LU)_S| of the current factor (src/async_ilu_factor.cpp:180-217,
 *  src/async_blockilu_factor.cpp:257-297).  Valid for ILU0-type objects after compute(). */
int b200_prec_ilu_residual(b200_prec *p, double *res);
/** Device time (ms, CUDA events) of the last compute() and the last apply() on this handle:
 *  the device-side counterpart of Blasted_data.{factor,apply}walltime (include/blasted_petsc.h:31-64). */
int b200_prec_last_times(b200_prec *p, double *compute_ms, double *apply_ms);

/* ---- Krylov test drivers on the device (tests/solvers.cpp:90-352) ---- */

typedef struct b200_solve_info {
	int converged;
	int iters;            /* as SolveInfo::iters: BiCGSTAB step+1, GCR/Richardson step */
	double resnorm;       /* 2-norm of the final residual */
	double bnorm;         /* 2-norm of the right-hand side */
	double device_ms;     /* CUDA-event time of the whole solve */
	double prec_ms;       /* of which preconditioner applications */
} b200_solve_info;

/** solver: "bicgstab" (tests/solvers.cpp:140-244), "gcr" (:252-352, == FGMRES in exact
 *  arithmetic, tests/solvers.hpp:108-110), "richardson" (:90-133), and "fgmres": flexible
 *  GMRES(restart) proper (right-preconditioned, classical Gram-Schmidt, Givens rotations; what
 *  PETSc's -ksp_type fgmres runs around the reference; iters = inner iterations).  Host b/x. */
int b200_solve_host(const char *solver, const b200_mat *A, b200_prec *M, const double *b, double *x,
                    double tol, int maxiter, int restart, b200_solve_info *info);
/** Device b/x. */
int b200_solve(const char *solver, const b200_mat *A, b200_prec *M, const double *d_b, double *d_x,
               double tol, int maxiter, int restart, b200_solve_info *info);

/* ---- one subdomain per GPU: row-partitioned operator, NCCL halo exchange and all-reduce ----
 * The reference is the LOCAL preconditioner under PETSc's bjacobi/asm (doc/user-doc.md:36-40;
 * one block per rank, src/blasted_petsc.cpp:604-606); these entry points provide the outer
 * distributed operator and Krylov reductions that PETSc's MatMult/VecDot provide there. */

typedef struct b200_comm b200_comm;
typedef struct b200_dist_mat b200_dist_mat;

/** Load NCCL from `path` (NULL: the libnccl.so.2 already in the process / on the loader path). */
int b200_nccl_load(const char *path);
/** Rank 0 creates the 128-byte NCCL unique id; the caller broadcasts it to the other ranks. */
int b200_comm_unique_id(char id[128]);
int b200_comm_create(const char id[128], int rank, int world, b200_comm **out);
void b200_comm_destroy(b200_comm *c);
int b200_comm_allreduce_sum(b200_comm *c, double *vals, int n);

/** diag: the square local diagonal block (local column numbering, what Mat_SeqAIJ of PCBJACOBI
 *  gives, src/blasted_petsc.cpp:278-298).  offd: the local rows' couplings to other subdomains,
 *  its column indices pointing into the halo buffer (may be NULL).  The halo buffer is the
 *  concatenation, in neighbour order, of recv_counts[k] (block) entries from neigh_ranks[k];
 *  send_idx lists, grouped by neighbour, the local (block) rows sent to it. */
int b200_dist_mat_create(b200_comm *comm, b200_mat *diag, b200_mat *offd, int nhalo, int nneigh,
                         const int *neigh_ranks, const int *send_counts, const int *send_idx,
                         const int *recv_counts, b200_dist_mat **out);
void b200_dist_mat_destroy(b200_dist_mat *d);
/** y_local = (A x)_local with halo exchange (device pointers, collective over the communicator). */
int b200_dist_mat_apply(b200_dist_mat *d, const double *d_x, double *d_y);
/** The same product with a halo the CALLER has filled (nhalo*bs values in the plan's receive order,
 *  e.g. by a PETSc VecScatter): y = A_diag x + A_offd halo.  Not collective, no NCCL; the coupling
 *  part touches the subdomain's boundary rows only. */
int b200_dist_mat_apply_with_halo(b200_dist_mat *d, const double *d_x, const double *d_halo, double *d_y);
/** The Krylov drivers of b200_solve on the partitioned operator; M is the local (block-Jacobi)
 *  preconditioner of `diag`.  Collective.  info->iters etc. are identical on every rank. */
int b200_dist_solve(const char *solver, b200_dist_mat *A, b200_prec *M, const double *d_b,
                    double *d_x, double tol, int maxiter, int restart, b200_solve_info *info);

#ifdef __cplusplus
}
#endif
#endif
