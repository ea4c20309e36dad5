/** \file ref_driver.cpp
 * \brief C entry points over the UNMODIFIED BLASTed reference objects (oracle build only).
 *
 * TEST INFRASTRUCTURE.  This translation unit is compiled together with the reference sources
 * where they lie under /root/reference (see oracle/Makefile) into oracle/_ref/libblasted_ref.so.
 * It is used (1) to pin the plain-C restatement in oracle/blasted_oracle.c, (2) to generate the
 * golden vectors under tests/golden/, and (3) as the CPU baseline ("kind": "reference") in bench.py.
 * Nothing in the product path (blasted_b200/) links or loads it, and it links nothing of the product:
 * the entry points that drive the product's C++ adapters live in oracle/adapter_driver.cpp
 * (oracle/_ref/libb200_adapters.so).
 *
 * It drives the reference exactly the way its own callers do:
 *   - tests/testsolve.cpp:43-88   (matrix views, SRFactory::create_preconditioner, compute)
 *   - src/blasted_petsc.cpp:278-298 (wrapping raw CSR/BSR arrays in SRMatrixStorage<const,const>)
 *   - tests/solverops/async_ilu_convergence.cpp:24-30 (direct access to pattern + residual functions)
 */

#include <cstring>
#include <string>
#include <vector>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <omp.h>

#include "solverfactory.hpp"
#include "solverops_ilu0.hpp"
#include "solverops_jacobi.hpp"
#include "blockmatrices.hpp"
#include "ilu_pattern.hpp"
#include "levelschedule.hpp"
#include "async_ilu_factor.hpp"
#include "async_blockilu_factor.hpp"
#include "coomatrix.hpp"
#include "reorderingscaling.hpp"
#include "../tests/solvers.hpp"
#include "ref_driver_common.hpp"

using namespace blasted;
using namespace refdrv;

namespace {

thread_local std::string g_err;


// Access to protected factor storage, for parity on intermediate products
struct ExposeScalarILU : public AsyncILU0_SRPreconditioner<double,int> {
	using AsyncILU0_SRPreconditioner<double,int>::iluvals;
	using AsyncILU0_SRPreconditioner<double,int>::scale;
};
template <int bs, StorageOptions stor>
struct ExposeBlockILU : public AsyncBlockILU0_SRPreconditioner<double,int,bs,stor> {
	using AsyncBlockILU0_SRPreconditioner<double,int,bs,stor>::iluvals;
	using AsyncBlockILU0_SRPreconditioner<double,int,bs,stor>::scale;
};
struct ExposeJacobi : public JacobiSRPreconditioner<double,int> {
	using JacobiSRPreconditioner<double,int>::dblocks;
};
template <int bs, StorageOptions stor>
struct ExposeBJacobi : public BJacobiSRPreconditioner<double,int,bs,stor> {
	using BJacobiSRPreconditioner<double,int,bs,stor>::dblocks;
};


}

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
void ref_set_error(const char *msg) { g_err = msg ? msg : ""; }

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

/// Mirrors SRFactory::create_preconditioner (src/solverfactory.cpp:131) over raw arrays
void *ref_prec_create(const char *prectype, int bs, int rowmajor, int scale,
                      int nbuildsweeps, int napplysweeps, int fact_init, int apply_init,
                      int thread_chunk_size, int compute_precinfo,
                      int nbrows, const int *browptr, const int *bcolind, const double *vals,
                      const int *diagind)
{
	try {
		SRFactory<double,int> fact;
		AsyncSolverSettings s;
		s.prectype = fact.solverTypeFromString(prectype);
		s.bs = bs;
		s.blockstorage = rowmajor ? RowMajor : ColMajor;
		s.relax = false;
		s.thread_chunk_size = thread_chunk_size;
		s.scale = scale;
		s.nbuildsweeps = nbuildsweeps;
		s.napplysweeps = napplysweeps;
		s.fact_inittype = static_cast<FactInit>(fact_init);
		s.apply_inittype = static_cast<ApplyInit>(apply_init);
		s.compute_precinfo = compute_precinfo;
		RefPrec *h = new RefPrec;
		h->bs = bs;
		h->p = fact.create_preconditioner(wrap(nbrows, browptr, bcolind, vals, diagind, bs), s);
		return h;
	} catch(std::exception& e) {
		g_err = e.what();
		return nullptr;
	}
}

int ref_prec_compute(void *hh, double info[6])
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	try {
		CoutMute m;
		const PrecInfo pi = h->p->compute();
		if(info) for(int i = 0; i < 6; i++) info[i] = pi.f_info[i];
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

int ref_prec_apply(void *hh, const double *r, double *z)
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	try { h->p->apply(r, z); } catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

int ref_prec_apply_relax(void *hh, const double *b, double *x, int maxits)
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	try {
		// src/blasted_petsc.cpp:532 : {rtol, atol, dtol, ctol=false, maxits}
		h->p->setApplyParams(SolveParams<double>{1e-10, 1e-50, 1e10, false, maxits});
		h->p->apply_relax(b, x);
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

int ref_prec_dim(void *hh) { return static_cast<RefPrec*>(hh)->p->dim(); }
int ref_prec_relaxation_available(void *hh) {
	return static_cast<RefPrec*>(hh)->p->relaxationAvailable() ? 1 : 0;
}

/// Copies out the ILU factor values of an ILU0-type preconditioner (after compute)
int ref_prec_get_factor(void *hh, int rowmajor, long long n, double *out)
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	const double *src = nullptr;
	if(h->bs == 1) {
		auto *q = dynamic_cast<AsyncILU0_SRPreconditioner<double,int>*>(h->p);
		if(q) src = static_cast<ExposeScalarILU*>(q)->iluvals;
	} else if(h->bs == 4 && !rowmajor) {
		auto *q = dynamic_cast<AsyncBlockILU0_SRPreconditioner<double,int,4,ColMajor>*>(h->p);
		if(q) src = static_cast<ExposeBlockILU<4,ColMajor>*>(q)->iluvals;
	} else if(h->bs == 4 && rowmajor) {
		auto *q = dynamic_cast<AsyncBlockILU0_SRPreconditioner<double,int,4,RowMajor>*>(h->p);
		if(q) src = static_cast<ExposeBlockILU<4,RowMajor>*>(q)->iluvals;
	} else if(h->bs == 5) {
		auto *q = dynamic_cast<AsyncBlockILU0_SRPreconditioner<double,int,5,ColMajor>*>(h->p);
		if(q) src = static_cast<ExposeBlockILU<5,ColMajor>*>(q)->iluvals;
	}
	if(!src) { g_err = "not an ILU0 preconditioner"; return 1; }
	std::memcpy(out, src, n*sizeof(double));
	return 0;
}

/// Copies out the inverted diagonal (blocks) of a Jacobi-derived preconditioner (after compute)
int ref_prec_get_dblocks(void *hh, int rowmajor, long long n, double *out)
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	const double *src = nullptr;
	if(h->bs == 1) {
		auto *q = dynamic_cast<JacobiSRPreconditioner<double,int>*>(h->p);
		if(q) src = static_cast<ExposeJacobi*>(q)->dblocks;
	} else if(h->bs == 4 && !rowmajor) {
		auto *q = dynamic_cast<BJacobiSRPreconditioner<double,int,4,ColMajor>*>(h->p);
		if(q) src = static_cast<ExposeBJacobi<4,ColMajor>*>(q)->dblocks;
	} else if(h->bs == 4 && rowmajor) {
		auto *q = dynamic_cast<BJacobiSRPreconditioner<double,int,4,RowMajor>*>(h->p);
		if(q) src = static_cast<ExposeBJacobi<4,RowMajor>*>(q)->dblocks;
	} else if(h->bs == 5) {
		auto *q = dynamic_cast<BJacobiSRPreconditioner<double,int,5,ColMajor>*>(h->p);
		if(q) src = static_cast<ExposeBJacobi<5,ColMajor>*>(q)->dblocks;
	}
	if(!src) { g_err = "not a Jacobi-derived preconditioner"; return 1; }
	std::memcpy(out, src, n*sizeof(double));
	return 0;
}

void ref_prec_destroy(void *hh)
{
	RefPrec *h = static_cast<RefPrec*>(hh);
	if(h) { delete h->p; delete h; }
}

/// y = A x through CSRMatrixView / BSRMatrixView (src/blockmatrices.ipp:118-162)
int ref_spmv(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
             const double *vals, const int *diagind, const double *x, double *y)
{
	try {
		if(bs == 1) CSRMatrixView<double,int>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 3 && !rowmajor) BSRMatrixView<double,int,3,ColMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 3 && rowmajor) BSRMatrixView<double,int,3,RowMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 4 && !rowmajor) BSRMatrixView<double,int,4,ColMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 4 && rowmajor) BSRMatrixView<double,int,4,RowMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 5 && !rowmajor) BSRMatrixView<double,int,5,ColMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 7 && !rowmajor) BSRMatrixView<double,int,7,ColMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else if(bs == 7 && rowmajor) BSRMatrixView<double,int,7,RowMajor>(nbrows, browptr, bcolind, vals, diagind).apply(x, y);
		else { g_err = "unsupported block size"; return 1; }
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

/// z = a A x + b y
int ref_gemv3(int bs, int rowmajor, int nbrows, const int *browptr, const int *bcolind,
              const double *vals, const int *diagind, double a, const double *x, double b,
              const double *y, double *z)
{
	try {
		if(bs == 1) CSRMatrixView<double,int>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else if(bs == 3 && !rowmajor) BSRMatrixView<double,int,3,ColMajor>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else if(bs == 4 && !rowmajor) BSRMatrixView<double,int,4,ColMajor>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else if(bs == 4 && rowmajor) BSRMatrixView<double,int,4,RowMajor>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else if(bs == 5 && !rowmajor) BSRMatrixView<double,int,5,ColMajor>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else if(bs == 7 && !rowmajor) BSRMatrixView<double,int,7,ColMajor>(nbrows, browptr, bcolind, vals, diagind).gemv3(a,x,b,y,z);
		else { g_err = "unsupported block size"; return 1; }
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

/// ILU(0) position lists (src/ilu_pattern.cpp:32).  Two-step: create, query sizes, copy, destroy.
void *ref_ilu_positions_create(int nbrows, const int *browptr, const int *bcolind,
                               const int *diagind)
{
	CRawBSRMatrix<double,int> mat(browptr, bcolind, nullptr, diagind, browptr+1, nbrows,
	                              browptr[nbrows], browptr[nbrows]);
	ILUPositions<int> *pl = new ILUPositions<int>;
	*pl = compute_ILU_positions_CSR_CSR<double,int>(&mat);
	return pl;
}
long long ref_ilu_positions_size(void *p) { return (long long)static_cast<ILUPositions<int>*>(p)->lowerp.size(); }
void ref_ilu_positions_copy(void *p, int *posptr, int *lowerp, int *upperp)
{
	ILUPositions<int> *pl = static_cast<ILUPositions<int>*>(p);
	std::memcpy(posptr, pl->posptr.data(), pl->posptr.size()*sizeof(int));
	std::memcpy(lowerp, pl->lowerp.data(), pl->lowerp.size()*sizeof(int));
	std::memcpy(upperp, pl->upperp.data(), pl->upperp.size()*sizeof(int));
}
void ref_ilu_positions_destroy(void *p) { delete static_cast<ILUPositions<int>*>(p); }

/// Level schedule (src/levelschedule.cpp:12).  Returns number of entries written (nlevels+1), or -1.
int ref_compute_levels(int nbrows, const int *browptr, const int *bcolind, const int *diagind,
                       int *levels_out, int capacity)
{
	try {
		CRawBSRMatrix<double,int> mat(browptr, bcolind, nullptr, diagind, browptr+1, nbrows,
		                              browptr[nbrows], browptr[nbrows]);
		fflush(stdout);
		FILE *old = stdout; (void)old;
		const std::vector<int> lv = computeLevels<double,int>(&mat);
		if((int)lv.size() > capacity) { g_err = "capacity"; return -1; }
		std::memcpy(levels_out, lv.data(), lv.size()*sizeof(int));
		return (int)lv.size();
	} catch(std::exception& e) { g_err = e.what(); return -1; }
}

/// Nonlinear ILU residual (src/async_ilu_factor.cpp:180, src/async_blockilu_factor.cpp:257).
/// Only the instantiations the reference library itself provides: bs=1, and bs=4 column-major.
int ref_ilu_nonlinear_res(int bs, int nbrows, const int *browptr, const int *bcolind,
                          const double *vals, const int *diagind, const double *scale,
                          const double *iluvals, int thread_chunk_size, double *res)
{
	CRawBSRMatrix<double,int> mat(browptr, bcolind, vals, diagind, browptr+1, nbrows,
	                              browptr[nbrows], browptr[nbrows]);
	const ILUPositions<int> pl = compute_ILU_positions_CSR_CSR<double,int>(&mat);
	if(bs == 1) {
		*res = scale ? scalar_ilu0_nonlinear_res<double,int,true,true>(&mat, pl, thread_chunk_size, scale, scale, iluvals)
			: scalar_ilu0_nonlinear_res<double,int,false,false>(&mat, pl, thread_chunk_size, scale, scale, iluvals);
	} else if(bs == 4) {
		*res = scale ? block_ilu0_nonlinear_res<double,int,4,ColMajor,true>(&mat, pl, scale, iluvals, thread_chunk_size)
			: block_ilu0_nonlinear_res<double,int,4,ColMajor,false>(&mat, pl, scale, iluvals, thread_chunk_size);
	} else { g_err = "residual: only bs 1 and 4 instantiated in the reference"; return 1; }
	return 0;
}

/// Krylov drivers of tests/solvers.cpp over a preconditioner handle.
/// solver: "bicgstab" | "gcr" | "richardson".  Returns iterations in *iters.
int ref_solve(const char *solver, void *prechandle, int bs, int rowmajor,
              int nbrows, const int *browptr, const int *bcolind, const double *vals,
              const int *diagind, const double *b, double *x, double tol, int maxiter,
              int restart, int *iters, double *relres, double *walltime)
{
	RefPrec *h = static_cast<RefPrec*>(prechandle);
	try {
		CoutMute m;
		SRMatrixView<double,int> *A = nullptr;
		if(bs == 1) A = new CSRMatrixView<double,int>(nbrows, browptr, bcolind, vals, diagind);
		else if(bs == 4 && !rowmajor) A = new BSRMatrixView<double,int,4,ColMajor>(nbrows, browptr, bcolind, vals, diagind);
		else if(bs == 4 && rowmajor) A = new BSRMatrixView<double,int,4,RowMajor>(nbrows, browptr, bcolind, vals, diagind);
		else if(bs == 5) A = new BSRMatrixView<double,int,5,ColMajor>(nbrows, browptr, bcolind, vals, diagind);
		else { g_err = "unsupported block size"; return 1; }

		IterativeSolver *s = nullptr;
		const std::string sn(solver);
		if(sn == "bicgstab") s = new BiCGSTAB(*A, *h->p);
		else if(sn == "gcr") s = new GCR(*A, *h->p, restart);
		else if(sn == "richardson") s = new RichardsonSolver(*A, *h->p);
		else { delete A; g_err = "unknown solver"; return 1; }
		s->setParams(tol, maxiter);
		fflush(stdout);
		const SolveInfo info = s->solve(b, x);
		if(iters) *iters = info.iters;
		if(relres) *relres = info.resnorm/info.bnorm;
		if(walltime) *walltime = info.walltime;
		delete s;
		delete A;
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}


// ---- front end: COOMatrix::readMatrixMarket + getSRMatrixFromCOO (src/coomatrix.cpp:189-443)

struct RefSRMat {
	SRMatrixStorage<double,int> m; int bs;
	RefSRMat(SRMatrixStorage<double,int>&& mm, int b) : m(std::move(mm)), bs(b) { }
};

void *ref_read_mtx(const char *file, int bs, int rowmajor)
{
	try {
		CoutMute mute;
		COOMatrix<double,int> coo;
		coo.readMatrixMarket(file);
		const std::string so = rowmajor ? "rowmajor" : "colmajor";
		if(bs == 1) return new RefSRMat(getSRMatrixFromCOO<double,int,1>(coo, so), bs);
		if(bs == 3) return new RefSRMat(getSRMatrixFromCOO<double,int,3>(coo, so), bs);
		if(bs == 4) return new RefSRMat(getSRMatrixFromCOO<double,int,4>(coo, so), bs);
		if(bs == 7) return new RefSRMat(getSRMatrixFromCOO<double,int,7>(coo, so), bs);
		g_err = "getSRMatrixFromCOO: only bs 1,3,4,7 instantiated in the reference";
		return nullptr;
	} catch(std::exception& e) { g_err = e.what(); return nullptr; }
}
int ref_srmat_nbrows(void *hh) { return static_cast<RefSRMat*>(hh)->m.nbrows; }
int ref_srmat_nnzb(void *hh) { return static_cast<RefSRMat*>(hh)->m.nnzb; }
void ref_srmat_copy(void *hh, int *browptr, int *bcolind, int *diagind, double *vals)
{
	RefSRMat *h = static_cast<RefSRMat*>(hh);
	const int n = h->m.nbrows, nz = h->m.nnzb, bs2 = h->bs*h->bs;
	for(int i = 0; i <= n; i++) browptr[i] = h->m.browptr[i];
	for(int i = 0; i < n; i++) diagind[i] = h->m.diagind[i];
	for(int i = 0; i < nz; i++) bcolind[i] = h->m.bcolind[i];
	for(long long i = 0; i < (long long)nz*bs2; i++) vals[i] = h->m.vals[i];
}
void ref_srmat_destroy(void *hh) { delete static_cast<RefSRMat*>(hh); }

/// Reordering / ReorderingScaling (src/reorderingscaling.cpp) applied in place to a matrix (may be
/// null) and to a row-direction and a column-direction vector (may be null).  Scaling is applied
/// before the ordering, to the arrays as they are numbered on entry.
int ref_reorder_scale(int bs, int nbrows, int *browptr, int *bcolind, double *vals, int *diagind,
                      const int *rord, const int *cord, const double *rowscale, const double *colscale,
                      int inverse, double *rowvec, double *colvec)
{
	try {
		if(bs == 1) ref_reorder_scale<1>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else if(bs == 4) ref_reorder_scale<4>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else if(bs == 7) ref_reorder_scale<7>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else { g_err = "Reordering: only bs 1,4,7 instantiated in the reference"; return 1; }
	} catch(std::exception& e) { g_err = e.what(); return 1; }
	return 0;
}

}
