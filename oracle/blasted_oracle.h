/** \file blasted_oracle.h
 * \brief Plain-C CPU restatement of the BLASTed asynchronous-preconditioner hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.  The product path
 * (blasted_b200/, libblasted_b200.so) never links, loads or calls anything under oracle/.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_vs_reference.py,
 * run in the build container) against the unmodified reference compiled from /root/reference into
 * oracle/_ref/libblasted_ref.so, and against the golden vectors committed under tests/golden/
 * (generated from that reference build by tests/golden/make_golden.py).
 *
 * All functions are sequential and deterministic: a sequential ascending-row sweep of the
 * asynchronous kernels is the reference's own definition of the exact ILU(0) factorisation and of
 * exact triangular substitution (tests/solverops/async_ilu_convergence.cpp:462-490,
 * tests/solverops/async_triangular_factors_convergence.cpp:289-367).
 *
 * Conventions: CSR/BSR with int32 indices, fp64 values; block jj occupies vals[jj*bs*bs ...];
 * inside a block entry (r,c) is at c*bs+r (column-major, rowmajor=0) or r*bs+c (rowmajor=1).
 * Citations are relative to /root/reference.
 */
#ifndef BLASTED_ORACLE_H
#define BLASTED_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BS 8

/* enum values follow include/async_initialization_decl.hpp:16-35 */
enum { ORC_INIT_F_ZERO = 0, ORC_INIT_F_ORIGINAL = 1, ORC_INIT_F_SGS = 2, ORC_INIT_F_NONE = 3 };
enum { ORC_INIT_A_ZERO = 0, ORC_INIT_A_JACOBI = 1, ORC_INIT_A_NONE = 2 };

/* ---- K9: SpMV (src/blas/matvecs.cpp:25-108) ---- */
void orc_spmv(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
              const double *vals, const double *x, double *y);
void orc_gemv3(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
               const double *vals, double a, const double *x, double b, const double *y, double *z);

/* ---- T4: ILU(0) position lists (src/ilu_pattern.cpp:32-163) ----
 * posptr has nnzb+1 entries and is always written. lowerp/upperp may be NULL (count-only pass).
 * Returns the total number of positions (posptr[nnzb]) as long long. */
long long orc_ilu_positions(int nbrows, const int *browptr, const int *bcolind, const int *diagind,
                            int *posptr, int *lowerp, int *upperp);

/* ---- T5: level schedule (src/levelschedule.cpp:12-71), contiguous greedy levels ----
 * levels must have room for nbrows+1 ints.  Returns nlevels+1 (entries written), -1 on a
 * structurally non-symmetric pattern ("Faulty dependency list!"). */
int orc_compute_levels(int nbrows, const int *browptr, const int *bcolind, int *levels);

/* True dependency (DAG) levels of the lower triangle: level[i] = 1 + max_{j<i, a_ij != 0} level[j]
 * (level 0 for rows without lower entries).  This is the device path's own wavefront schedule
 * (SURVEY.md section 7); it has no counterpart in the reference.  Returns the number of levels. */
int orc_dag_levels(int nbrows, const int *browptr, const int *bcolind, const int *diagind,
                   int *level_of_row);

/* ---- T6: symmetric scaling vector (src/rawsrmatrixutils.cpp:343-350) ---- */
void orc_scaling_vector(int bs, int nbrows, const double *vals, const int *diagind, double *scale);

/* ---- dense block inverse with partial pivoting (stands in for Eigen's .inverse()) ---- */
void orc_block_inverse(int bs, int rowmajor, const double *a, double *ainv);

/* ---- K4: factor initialisation (src/async_ilu_factor.cpp:47-58,110-151;
 *          src/async_blockilu_factor.cpp:63-94,207-254).  scale may be NULL. ---- */
void orc_ilu0_init(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *scale, int fact_init,
                   double *iluvals);

/* ---- K1/K2: nsweeps sequential ascending-row sweeps of the async ILU(0) row kernel
 *      (src/kernels/kernels_ilu0_factorize.hpp:19-53 scalar, :71-98 block).
 *      Diagonal blocks are left UN-inverted, as during the reference's sweeps. ---- */
void orc_ilu0_sweeps(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                     const double *vals, const int *diagind, const int *posptr, const int *lowerp,
                     const int *upperp, const double *scale, int nsweeps, double *iluvals);

/* One fully synchronous (Jacobi-type) sweep: every entry is computed from the PREVIOUS iterate.
 * This is the opposite extreme of chaotic iteration from the sequential sweep above and brackets
 * what a massively parallel device sweep does; used to test convergence-rate expectations. */
void orc_ilu0_sweep_synchronous(int bs, int rowmajor, int nbrows, const int *browptr,
                                const int *bcolind, const double *vals, const int *diagind,
                                const int *posptr, const int *lowerp, const int *upperp,
                                const double *scale, const double *ilu_old, double *ilu_new);

/* ---- K3: invert diagonal blocks in place (src/async_blockilu_factor.cpp:144-146);
 *      for bs==1 this is a no-op (the scalar apply divides, src/solverops_ilu0.cpp:312). ---- */
void orc_ilu0_invert_diag(int bs, int rowmajor, int nbrows, const int *diagind, double *iluvals);

/* ---- K10: nonlinear residual sum |(A - LU)_S| (src/async_ilu_factor.cpp:180-217;
 *      src/async_blockilu_factor.cpp:257-297); iluvals with UN-inverted diagonal blocks. ---- */
double orc_ilu0_nonlinear_res(int bs, int rowmajor, int nbrows, const int *browptr,
                              const int *bcolind, const double *vals, const int *diagind,
                              const int *posptr, const int *lowerp, const int *upperp,
                              const double *scale, const double *iluvals);

/* Entry-wise 1-norm of the (scaled) matrix, the normalisation used in BASELINE.json's
 * "||(A-LU)|_S|| / ||A||". */
double orc_matrix_abs_sum(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                          const double *vals, const double *scale);

/* ---- K5: ILU(0) application, sequential sweeps (src/solverops_ilu0.cpp:56-148 block with
 *      pre-inverted diagonal blocks, :240-321 scalar).  ytemp: nbrows*bs scratch.
 *      Returns 0, or 1 for apply_init == NONE ("Invalid init type!", :125-127, :298-300). ---- */
int orc_ilu0_apply(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                   const int *diagind, const double *iluvals, const double *scale, int napplysweeps,
                   int apply_init, const double *r, double *z, double *ytemp);

/* ---- K3: (block-)Jacobi setup dblocks = D^-1 (src/solverops_jacobi.cpp:31-48,141-147)
 *      and application z = D^-1 r (:50-63,164-171) ---- */
void orc_jacobi_setup(int bs, int rowmajor, int nbrows, const double *vals, const int *diagind,
                      double *dblocks);
void orc_jacobi_apply(int bs, int rowmajor, int nbrows, const double *dblocks, const double *r,
                      double *z);

/* ---- K6: SGS application, sequential sweeps (src/solverops_sgs.cpp:48-83 block, :148-177
 *      scalar; row kernels src/kernels/kernels_sgs.hpp:17-76).  apply_init NONE leaves y,z as
 *      they are (no throw in the SGS objects). ---- */
void orc_sgs_apply(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *dblocks, int napplysweeps,
                   int apply_init, const double *r, double *z, double *ytemp);

/* ---- K7: SGS relaxation, maxits x (forward pass, backward pass) in place
 *      (src/solverops_sgs.cpp:86-116,180-203; src/kernels/kernels_relaxation.hpp:17-54) ---- */
void orc_sgs_relax(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *dblocks, int maxits,
                   const double *b, double *x);

/* Forward-only chaotic relaxation sweeps (src/relaxation_chaotic.cpp:22-123), sequential order */
void orc_gs_relax(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                  const double *vals, const int *diagind, const double *dblocks, int nsweeps,
                  const double *b, double *x);

/* Jacobi relaxation without tolerance checks (src/solverops_jacobi.cpp:66-121,174-220, ctol=false) */
void orc_jacobi_relax(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
                      const double *vals, const int *diagind, const double *dblocks, int maxits,
                      const double *b, double *x, double *xtemp);

/* ---- diagonal dominance of the factors (src/matrix_properties.cpp:11-77):
 *      out = {lower avg, lower min, upper avg, upper min} ---- */
void orc_diagonal_dominance(int bs, int rowmajor, int nbrows, const int *browptr, const int *diagind,
                            const double *vals, double out[4]);

/* ---- front end ---- */

/* Coordinate triplets -> CSR/BSR: sort by row then column (src/coomatrix.cpp:222-260), then
 * convertToCSR (:262-297) / convertToBSR<bs,stor> (:299-403: block columns of a block row in order of
 * first appearance while scanning its scalar rows, blocks zero-filled, diagind = -1 where absent).
 * Call with bcolind == NULL to get the number of stored blocks; arrays sized by that on the 2nd call. */
long long orc_coo_convert(int nrows, long long nnz, const int *rowind, const int *colind,
                          const double *val, int bs, int rowmajor, int *browptr, int *bcolind,
                          int *diagind, double *vals);

/* Reordering::applyOrdering, matrix (src/reorderingscaling.cpp:77-205), in place; either ordering
 * may be NULL; diagind is not touched (nor is it in the reference) */
void orc_reorder_matrix(int bs, int nbrows, int *browptr, int *bcolind, double *vals,
                        const int *rord, const int *cord, int inverse);
/* Reordering::applyOrdering, vector (:211-266): forward v[i] <- v[ord[i]], inverse v[ord[i]] <- v[i] */
void orc_reorder_vector(int bs, int n, const int *ord, int inverse, double *vec);
/* ReorderingScaling::applyScaling, matrix (:282-337) and vector (:340-368) */
void orc_scale_matrix(int bs, int nbrows, const int *browptr, const int *bcolind, double *vals,
                      const double *rowscale, const double *colscale, int inverse);
void orc_scale_vector(int bs, int n, const double *scale, int inverse, double *vec);

#ifdef __cplusplus
}
#endif
#endif
