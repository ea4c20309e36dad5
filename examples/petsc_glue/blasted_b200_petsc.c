/** \file blasted_b200_petsc.c
 * \brief The PETSc-typed half of the PCSHELL glue: the C symbols of
 * the reference's include/blasted_petsc.h:88-167, implemented on libblasted_b200.so.
 *
 * Everything between PETSc's callbacks and the preconditioner object lives in the library
 * (include/blasted_b200_shell.h, csrc/shell.cu); this file is what is left: pulling contexts, raw
 * matrix arrays and vector arrays out of PETSc objects, the walk over the KSP/PC tree
 * (src/blasted_petsc.cpp:578-661), the options database -> b200_shell_options
 * (setupDataFromOptions, :137-208) and PETSc's error conventions.  An application switches from
 * the reference by linking this file + libblasted_b200.so instead of libblasted_petsc: names,
 * argument meaning and call order are the reference's (Blasted_data = b200_shell_node,
 * Blasted_data_list = b200_shell_list; the factory members are gone - the device library has one
 * factory).
 *
 * PETSc is not in the build image: the file is compiled and type-checked there against
 * examples/petsc_glue/petsc_stub/petscksp.h, a declaration-only header with PETSc's signatures
 * (tests/test_petsc_glue_cpu.py).  With PETSc, compile against <petscksp.h>,
 * <petsc/private/matimpl.h> and the Seq(B)AIJ implementation headers instead (B200_HAVE_PETSC).
 */
#ifdef B200_HAVE_PETSC
#include <petscksp.h>
#include <petsc/private/matimpl.h>
#include <../src/mat/impls/aij/seq/aij.h>
#include <../src/mat/impls/baij/seq/baij.h>
#else
#include "petscksp.h"
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "blasted_b200_shell.h"

typedef b200_shell_node Blasted_data;
typedef b200_shell_list Blasted_data_list;

PetscErrorCode setup_blasted_stack(KSP ksp, Blasted_data_list *const bctx);
PetscErrorCode setup_localpreconditioner_blasted(KSP ksp, Blasted_data *const bctx);
PetscErrorCode compute_preconditioner_blasted(PC pc);
PetscErrorCode apply_local_blasted(PC pc, Vec r, Vec z);
PetscErrorCode relax_local_blasted(PC pc, Vec rhs, Vec x, Vec w, PetscReal rtol, PetscReal abstol,
                                   PetscReal dtol, PetscInt it, PetscBool guesszero, PetscInt *outits,
                                   PCRichardsonConvergedReason *reason);
PetscErrorCode cleanup_blasted(PC pc);

/* ---- list functions: include/blasted_petsc.h:88-96, 122-131 ---- */

Blasted_data_list newBlastedDataList(void) { return b200_shell_list_new(); }
Blasted_data newBlastedDataContext(void) { return b200_shell_node_new(); }
void appendBlastedDataContext(Blasted_data_list *const bdl, const Blasted_data bd) { b200_shell_list_append(bdl, bd); }
void computeTotalTimes(Blasted_data_list *const bctv) { b200_shell_total_times(bctv); }
void destroyBlastedDataList(Blasted_data_list *const bdv)
{
	/* the reference throws std::logic_error on an inconsistent list (src/blasted_petsc.cpp:365);
	 * in C: report and abort, as its option errors do (:36-39) */
	if(b200_shell_list_destroy(bdv)) {
		fprintf(stderr, "BLASTed: %s\n", b200_last_error());
		abort();
	}
}

/* ---- options: get_*_petscoptions + setupDataFromOptions (src/blasted_petsc.cpp:25-92, 137-208) ---- */

static void need_string(const char *name, char *out)
{
	PetscBool set = PETSC_FALSE;
	PetscOptionsGetString(NULL, NULL, name, out, B200_OPT_STRLEN, &set);
	if(!set) { printf("BLASTed: String %s not set!\n", name); fflush(stdout); abort(); }
}

static PetscErrorCode options_to_node(PC pc, Blasted_data *ctx)
{
	b200_shell_options o;
	PetscBool set = PETSC_FALSE, flag = PETSC_FALSE;
	PetscInt nmax = 2, chunk = 0;
	PetscErrorCode ierr;
	char pcname[B200_OPT_STRLEN + 16];
	memset(&o, 0, sizeof o);
	need_string("-blasted_pc_type", o.pc_type);
	/* the remaining options are read where the type uses them (:149-185); reading them
	 * unconditionally and letting b200_shell_set_options ignore the unused ones is equivalent,
	 * except that a missing option must only be fatal where the reference makes it fatal */
	o.async_sweeps[0] = o.async_sweeps[1] = 1;
	PetscOptionsGetIntArray(NULL, NULL, "-blasted_async_sweeps", o.async_sweeps, &nmax, &flag);
	if(flag == PETSC_FALSE || nmax < 2) { o.async_sweeps[0] = o.async_sweeps[1] = 0; }   /* checked by set_options */
	PetscOptionsGetBool(NULL, NULL, "-blasted_use_symmetric_scaling", &set, &flag);
	o.use_symmetric_scaling = (flag && set) ? 1 : 0;
	strcpy(o.fact_init_type, "init_original");
	strcpy(o.apply_init_type, "init_jacobi");
	PetscOptionsGetString(NULL, NULL, "-blasted_async_fact_init_type", o.fact_init_type, B200_OPT_STRLEN, &flag);
	PetscOptionsGetString(NULL, NULL, "-blasted_async_apply_init_type", o.apply_init_type, B200_OPT_STRLEN, &flag);
	PetscOptionsGetInt(NULL, NULL, "-blasted_thread_chunk_size", &chunk, &flag);
	o.thread_chunk_size = chunk;
	set = PETSC_FALSE;
	PetscOptionsGetBool(NULL, NULL, "-blasted_compute_preconditioner_info", &set, &flag);
	o.compute_preconditioner_info = (flag && set) ? 1 : 0;
	if(b200_shell_set_options(ctx, &o)) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, b200_last_error());
	snprintf(pcname, sizeof pcname, "Blasted-%s", ctx->prectypestr);
	ierr = PCShellSetName(pc, pcname); CHKERRQ(ierr);
	return 0;
}

/* ---- PCSHELL callbacks ---- */

/** PCShellSetSetUp: compute_preconditioner_blasted, src/blasted_petsc.cpp:403-429 (with
 *  createNewPreconditioner :216-311 and updatePreconditioner :314-327 inside b200_shell_setup). */
PetscErrorCode compute_preconditioner_blasted(PC pc)
{
	Blasted_data *ctx = NULL;
	Mat A = NULL;
	PetscInt localrows = 0, localcols = 0, badrow = -1;
	PetscBool diagmissing = PETSC_FALSE;
	PetscErrorCode ierr;
	ierr = PCShellGetContext(pc, (void**)&ctx); CHKERRQ(ierr);
	if(!ctx->first_setup_done) { ierr = options_to_node(pc, ctx); CHKERRQ(ierr); }
	ierr = PCGetOperators(pc, NULL, &A); CHKERRQ(ierr);
	ierr = MatGetLocalSize(A, &localrows, &localcols); CHKERRQ(ierr);
	/* also makes PETSc compute the diagonal positions of BAIJ matrices (:241-246) */
	ierr = MatMissingDiagonal(A, &diagmissing, &badrow); CHKERRQ(ierr);
	if(diagmissing == PETSC_TRUE)
		SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_LIB, "! Zero diagonal in (block-)row %d!", badrow);
	if(ctx->bs == 1) {
		const Mat_SeqAIJ *const d = (const Mat_SeqAIJ*)A->data;
		if(b200_shell_setup(ctx, 1, localrows, d->i, d->j, d->a, d->diag))
			SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, b200_last_error());
	} else {
		const Mat_SeqBAIJ *const d = (const Mat_SeqBAIJ*)A->data;     /* column-major blocks (:256) */
		if(b200_shell_setup(ctx, ctx->bs, localrows/ctx->bs, d->i, d->j, d->a, d->diag))
			SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, b200_last_error());
	}
	return 0;
}

static int vec_on_device(Vec v)
{
	const char *type = NULL;
	if(VecGetType(v, &type) || !type) return 0;
	return strstr(type, "cuda") != NULL;                 /* VECSEQCUDA / VECMPICUDA / VECCUDA */
}

/** PCShellSetApply: apply_local_blasted, src/blasted_petsc.cpp:474-517.  Device vectors are used
 *  in place (zero copy), host vectors go through the library's staging copies. */
PetscErrorCode apply_local_blasted(PC pc, Vec r, Vec z)
{
	Blasted_data *ctx = NULL;
	const PetscScalar *ra = NULL;
	PetscScalar *za = NULL;
	PetscErrorCode ierr;
	int rc;
	ierr = PCShellGetContext(pc, (void**)&ctx); CHKERRQ(ierr);
	if(vec_on_device(r) && vec_on_device(z)) {
		ierr = VecCUDAGetArrayRead(r, &ra); CHKERRQ(ierr);
		ierr = VecCUDAGetArrayWrite(z, &za); CHKERRQ(ierr);
		rc = b200_shell_apply_device(ctx, ra, za);
		ierr = VecCUDARestoreArrayRead(r, &ra); CHKERRQ(ierr);
		ierr = VecCUDARestoreArrayWrite(z, &za); CHKERRQ(ierr);
	} else {
		ierr = VecGetArrayRead(r, &ra); CHKERRQ(ierr);
		ierr = VecGetArray(z, &za); CHKERRQ(ierr);
		rc = b200_shell_apply(ctx, ra, za);
		ierr = VecRestoreArrayRead(r, &ra); CHKERRQ(ierr);
		ierr = VecRestoreArray(z, &za); CHKERRQ(ierr);
	}
	if(rc) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, b200_last_error());
	return 0;
}

/** PCShellSetApplyRichardson: relax_local_blasted, src/blasted_petsc.cpp:519-576. */
PetscErrorCode relax_local_blasted(PC pc, Vec rhs, Vec x, Vec w, PetscReal rtol, PetscReal abstol,
                                   PetscReal dtol, PetscInt it, PetscBool guesszero, PetscInt *outits,
                                   PCRichardsonConvergedReason *reason)
{
	Blasted_data *ctx = NULL;
	const PetscScalar *ra = NULL;
	PetscScalar *xa = NULL;
	PetscErrorCode ierr;
	int rc, its = 0, why = 0;
	(void)w;
	ierr = PCShellGetContext(pc, (void**)&ctx); CHKERRQ(ierr);
	if(vec_on_device(rhs) && vec_on_device(x)) {
		ierr = VecCUDAGetArrayRead(rhs, &ra); CHKERRQ(ierr);
		ierr = VecCUDAGetArray(x, &xa); CHKERRQ(ierr);
		rc = b200_shell_relax_device(ctx, ra, xa, rtol, abstol, dtol, it, guesszero == PETSC_TRUE, &its, &why);
		ierr = VecCUDARestoreArrayRead(rhs, &ra); CHKERRQ(ierr);
		ierr = VecCUDARestoreArray(x, &xa); CHKERRQ(ierr);
	} else {
		ierr = VecGetArrayRead(rhs, &ra); CHKERRQ(ierr);
		ierr = VecGetArray(x, &xa); CHKERRQ(ierr);
		rc = b200_shell_relax(ctx, ra, xa, rtol, abstol, dtol, it, guesszero == PETSC_TRUE, &its, &why);
		ierr = VecRestoreArrayRead(rhs, &ra); CHKERRQ(ierr);
		ierr = VecRestoreArray(x, &xa); CHKERRQ(ierr);
	}
	if(rc) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, b200_last_error());
	*outits = its;
	*reason = (PCRichardsonConvergedReason)why;           /* PCRICHARDSON_CONVERGED_ITS (:573) */
	return 0;
}

/** PCShellSetDestroy: cleanup_blasted, src/blasted_petsc.cpp:391-401. */
PetscErrorCode cleanup_blasted(PC pc)
{
	Blasted_data *ctx = NULL;
	PetscErrorCode ierr = PCShellGetContext(pc, (void**)&ctx); CHKERRQ(ierr);
	if(b200_shell_cleanup(ctx)) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, b200_last_error());
	return 0;
}

/* ---- registration: setup_localpreconditioner_blasted, src/blasted_petsc.cpp:663-721 ---- */

PetscErrorCode setup_localpreconditioner_blasted(KSP ksp, Blasted_data *const bctx)
{
	Mat A = NULL;
	PC pc = NULL;
	PetscInt matbs = 1;
	MatType mtype = NULL;
	PetscBool isshell = PETSC_FALSE;
	PetscErrorCode ierr;
	int isblock, islocal;
	ierr = KSPGetOperators(ksp, NULL, &A); CHKERRQ(ierr);
	ierr = MatGetBlockSize(A, &matbs); CHKERRQ(ierr);
	ierr = MatGetType(A, &mtype); CHKERRQ(ierr);
	isblock = strstr(mtype, "baij") != NULL;
	islocal = strncmp(mtype, "seq", 3) == 0;
	ierr = KSPGetPC(ksp, &pc); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCSHELL, &isshell); CHKERRQ(ierr);
	if(!isshell) SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_ARG_WRONGSTATE, "Need SHELL preconditioner for BLASTed!\n");
	if(!islocal) SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "PC as PCSHELL is only supported for local solvers.");
	bctx->bs = isblock ? matbs : 1;
	bctx->first_setup_done = 0;
	ierr = PCShellSetContext(pc, (void*)bctx); CHKERRQ(ierr);
	ierr = PCShellSetSetUp(pc, &compute_preconditioner_blasted); CHKERRQ(ierr);
	ierr = PCShellSetApply(pc, &apply_local_blasted); CHKERRQ(ierr);
	ierr = PCShellSetDestroy(pc, &cleanup_blasted); CHKERRQ(ierr);
	/* The reference decides on the Richardson callback from bctx->prectype here (:709-716), before
	 * the options have been read into the node (uninitialised there); the type string is already in
	 * the options database, so it is asked instead: ilu0 / cscbgs / none leave PETSc's default. */
	{
		char pctype[B200_OPT_STRLEN] = "";
		PetscBool set = PETSC_FALSE;
		PetscOptionsGetString(NULL, NULL, "-blasted_pc_type", pctype, B200_OPT_STRLEN, &set);
		if(b200_shell_type_offers_relaxation(set ? pctype : NULL)) {
			ierr = PCShellSetApplyRichardson(pc, &relax_local_blasted); CHKERRQ(ierr);
		}
	}
	return 0;
}

/* ---- the walk over the solver stack: setup_blasted_stack, src/blasted_petsc.cpp:578-661 ---- */

PetscErrorCode setup_blasted_stack(KSP ksp, Blasted_data_list *const bctv)
{
	PC pc = NULL;
	PetscBool isbjacobi, isasm, isshell, ismg, isgamg, isksp;
	PetscErrorCode ierr;
	ierr = KSPGetPC(ksp, &pc); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCBJACOBI, &isbjacobi); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCASM, &isasm); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCSHELL, &isshell); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCMG, &ismg); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCGAMG, &isgamg); CHKERRQ(ierr);
	ierr = PetscObjectTypeCompare((PetscObject)pc, PCKSP, &isksp); CHKERRQ(ierr);
	if(isbjacobi || isasm) {
		PetscInt nlocalblocks = 0, firstlocalblock = 0;
		KSP *subksp = NULL;
		ierr = KSPSetUp(ksp); CHKERRQ(ierr);
		ierr = PCSetUp(pc); CHKERRQ(ierr);
		if(isbjacobi) { ierr = PCBJacobiGetSubKSP(pc, &nlocalblocks, &firstlocalblock, &subksp); CHKERRQ(ierr); }
		else { ierr = PCASMGetSubKSP(pc, &nlocalblocks, &firstlocalblock, &subksp); CHKERRQ(ierr); }
		if(nlocalblocks != 1)                       /* one subdomain = one GPU per rank (:604-606) */
			SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "Only one subdomain per rank is supported.");
		ierr = setup_blasted_stack(subksp[0], bctv); CHKERRQ(ierr);
	}
	else if(ismg || isgamg) {
		PetscInt nlevels = 0, ilvl;
		KSP smoother = NULL, coarse = NULL;
		ierr = KSPSetUp(ksp); CHKERRQ(ierr);
		ierr = PCSetUp(pc); CHKERRQ(ierr);
		ierr = PCMGGetLevels(pc, &nlevels); CHKERRQ(ierr);
		for(ilvl = 1; ilvl < nlevels; ilvl++) {
			ierr = PCMGGetSmoother(pc, ilvl, &smoother); CHKERRQ(ierr);
			ierr = setup_blasted_stack(smoother, bctv); CHKERRQ(ierr);
		}
		ierr = PCMGGetCoarseSolve(pc, &coarse); CHKERRQ(ierr);
		ierr = setup_blasted_stack(coarse, bctv); CHKERRQ(ierr);
	}
	else if(isksp) {
		KSP subksp = NULL;
		ierr = KSPSetUp(ksp); CHKERRQ(ierr);
		ierr = PCSetUp(pc); CHKERRQ(ierr);
		ierr = PCKSPGetKSP(pc, &subksp); CHKERRQ(ierr);
		ierr = setup_blasted_stack(subksp, bctv); CHKERRQ(ierr);
	}
	else if(isshell) {
		/* the new context goes to the head of the list (:646-655) */
		appendBlastedDataContext(bctv, newBlastedDataContext());
		ierr = setup_localpreconditioner_blasted(ksp, bctv->ctxlist); CHKERRQ(ierr);
	}
	return 0;
}
