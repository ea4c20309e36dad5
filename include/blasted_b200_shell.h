/** \file blasted_b200_shell.h
 * \brief PETSc-free core of BLASTed's PCSHELL glue, on top of
 * blasted_b200.h: everything src/blasted_petsc.cpp does between PETSc's callbacks and the
 * preconditioner object, with the PETSc types taken out of the signatures.
 *
 * The reference's PCSHELL callbacks (include/blasted_petsc.h:88-167) each do three things: pull a
 * context and raw arrays out of PETSc objects (PCShellGetContext, Mat_SeqAIJ/BAIJ::i,j,a,diag,
 * VecGetArray), run BLASTed, and account the time.  The first part needs PETSc headers (absent in
 * this image); the other two are here, so that the PETSc-side file is a page of unwrapping calls
 * (INTEGRATION.md section 5 shows it).  Field names follow struct Blasted_node
 * (include/blasted_petsc.h:31-64) so that the mapping is one to one.
 *
 * Return convention as PetscErrorCode: 0 = success; the message is b200_last_error().
 */
#ifndef BLASTED_B200_SHELL_H
#define BLASTED_B200_SHELL_H

#include "blasted_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OPT_STRLEN 20            /* BLASTED_OPT_STRLEN, include/blasted_petsc.h:24 */
#define B200_SEQUENTIAL_SYMBOL (-1)   /* BLASTED_SEQUENTIAL_SYMBOL, include/solvertypes.h:29 */

/** The options setupDataFromOptions reads from the PETSc options database
 *  (src/blasted_petsc.cpp:137-208), already fetched by the caller. */
typedef struct b200_shell_options {
	char pc_type[B200_OPT_STRLEN];         /* -blasted_pc_type                                   */
	int  async_sweeps[2];                  /* -blasted_async_sweeps build,apply (-1: sequential) */
	int  use_symmetric_scaling;            /* -blasted_use_symmetric_scaling                     */
	char fact_init_type[B200_OPT_STRLEN];  /* -blasted_async_fact_init_type                      */
	char apply_init_type[B200_OPT_STRLEN]; /* -blasted_async_apply_init_type                     */
	int  thread_chunk_size;                /* -blasted_thread_chunk_size                         */
	int  compute_preconditioner_info;      /* -blasted_compute_preconditioner_info (default 0)   */
} b200_shell_options;

/** struct Blasted_node (include/blasted_petsc.h:31-64) with device handles in place of the host
 *  objects. */
typedef struct b200_shell_node {
	b200_prec *bprec;
	b200_mat *bmat;
	int bs;
	char prectypestr[B200_OPT_STRLEN];
	int prectype;
	int scale;
	int threadchunksize;
	int nbuildsweeps, napplysweeps;
	char factinittype[B200_OPT_STRLEN], applyinittype[B200_OPT_STRLEN];
	int compute_precinfo;
	void *infolist;                        /* PrecInfo records, one per setup (6 doubles each)   */
	int first_setup_done;                  /* MUST start as 0                                    */
	double cputime, walltime, factorcputime, factorwalltime, applycputime, applywalltime;
	struct b200_shell_node *next;
} b200_shell_node;

/** Blasted_data_list (include/blasted_petsc.h:72-86) */
typedef struct b200_shell_list {
	b200_shell_node *ctxlist;
	int size;
	double factorcputime, factorwalltime, applycputime, applywalltime;
} b200_shell_list;

/* newBlastedDataList / newBlastedDataContext / appendBlastedDataContext / computeTotalTimes /
 * destroyBlastedDataList (src/blasted_petsc.cpp:331-388, 723-735).  destroy also frees the device
 * objects still attached to the nodes and returns non-zero if the list was inconsistent. */
b200_shell_list b200_shell_list_new(void);
b200_shell_node b200_shell_node_new(void);
void b200_shell_list_append(b200_shell_list *list, b200_shell_node node);
void b200_shell_total_times(b200_shell_list *list);
int b200_shell_list_destroy(b200_shell_list *list);

/** setupDataFromOptions (src/blasted_petsc.cpp:137-208): validates the type string, keeps the sweep
 *  counts / initialisation strings only where the type uses them, resets the timers. */
int b200_shell_set_options(b200_shell_node *node, const b200_shell_options *opts);
/** The settings createNewPreconditioner builds from the node (src/blasted_petsc.cpp:252-275),
 *  including setSweeps_checkSeq (:94-134): a sweep count of -1 selects the sequential variant. */
int b200_shell_settings(const b200_shell_node *node, b200_settings *out);

/** compute_preconditioner_blasted (src/blasted_petsc.cpp:403-429): on the first call (node set up by
 *  b200_shell_set_options) creates the device matrix and preconditioner from the local block's raw
 *  arrays - Mat_SeqAIJ/BAIJ i, j, a, diag, column-major blocks (:278-298) -, on every call uploads
 *  the current values, runs compute(), appends the PrecInfo when requested and accumulates the
 *  factorisation times.  nbrows counts block rows.  Block sizes as the glue accepts them (:281). */
int b200_shell_setup(b200_shell_node *node, int bs, int nbrows, const int *ia, const int *ja,
                     const double *a, const int *diag);
/** apply_local_blasted (src/blasted_petsc.cpp:474-517); host pointers (VecGetArray) or device
 *  pointers (VecCUDAGetArray: zero copy). */
int b200_shell_apply(b200_shell_node *node, const double *r, double *z);
int b200_shell_apply_device(b200_shell_node *node, const double *d_r, double *d_z);
/** relax_local_blasted (src/blasted_petsc.cpp:519-576): setApplyParams({rtol, abstol, dtol, false,
 *  it}), x := 0 if guesszero, apply_relax; *reason = 4 (PCRICHARDSON_CONVERGED_ITS), *outits = it. */
int b200_shell_relax(b200_shell_node *node, const double *rhs, double *x, double rtol, double abstol,
                     double dtol, int it, int guesszero, int *outits, int *reason);
int b200_shell_relax_device(b200_shell_node *node, const double *d_rhs, double *d_x, double rtol,
                            double abstol, double dtol, int it, int guesszero, int *outits,
                            int *reason);
/** Whether setup_localpreconditioner_blasted registers the Richardson callback for this type
 *  (src/blasted_petsc.cpp:709-716: not for ilu0, cscbgs, none). */
int b200_shell_offers_relaxation(const b200_shell_node *node);
/** The same question from the -blasted_pc_type string alone, for the registration step, which runs
 *  before the options are read into the node (the reference tests the node's still uninitialised
 *  prectype there, :709-716).  Unknown strings answer 1; they fail at the first set-up. */
int b200_shell_type_offers_relaxation(const char *pc_type);
/** cleanup_blasted (src/blasted_petsc.cpp:391-401) */
int b200_shell_cleanup(b200_shell_node *node);
/** PrecInfo records gathered so far (PrecInfoList, include/preconditioner_diagnostics.hpp) */
int b200_shell_info_count(const b200_shell_node *node);
int b200_shell_info_get(const b200_shell_node *node, int i, double precinfo[6]);

#ifdef __cplusplus
}
#endif
#endif
