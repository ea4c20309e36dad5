/** \file b200_solverops.cpp
 * \brief Implementation of the reference-facing C++ adapters over the C ABI (see the header).
 */
#include <stdexcept>
#include <string>
#include "b200_solverops.hpp"
// SRMatrixView<double,int> has no explicit instantiation in the reference library (only its
// CSR/BSR subclasses, src/csrmatrix.cpp, src/bsrmatrix.cpp): pull in the template definitions
#include "blockmatrices.ipp"

namespace blasted_b200 {

static void check_runtime(const int rc)
{
	if(rc) throw std::runtime_error(b200_last_error());
}

B200Preconditioner::B200Preconditioner(SRMatrixStorage<const double,const int>&& matrix,
                                       const AsyncSolverSettings& s)
	: SRPreconditioner<double,int>(std::move(matrix)), bs{s.bs}, dmat{nullptr}, dprec{nullptr}
{
	const int storage = (s.blockstorage == blasted::RowMajor) ? B200_ROWMAJOR : B200_COLMAJOR;
	if(b200_mat_create_host(mat.nbrows, bs, storage, mat.browptr, mat.bcolind, mat.vals, mat.diagind,
	                        &dmat))
		throw std::invalid_argument(b200_last_error());
	b200_settings cs;
	cs.prectype = static_cast<int>(s.prectype);          // enum values identical (solvertypes.h:14-26)
	cs.bs = s.bs;
	cs.blockstorage = storage;
	cs.relax = s.relax;
	cs.thread_chunk_size = s.thread_chunk_size;
	cs.scale = s.scale;
	cs.nbuildsweeps = s.nbuildsweeps;
	cs.napplysweeps = s.napplysweeps;
	cs.fact_inittype = static_cast<int>(s.fact_inittype);  // async_initialization_decl.hpp:16-35
	cs.apply_inittype = static_cast<int>(s.apply_inittype);
	cs.compute_precinfo = s.compute_precinfo;
	cs.level_mode = B200_LEVELS_DAG;
	if(b200_prec_create(&cs, dmat, &dprec)) {
		const std::string msg = b200_last_error();
		b200_mat_destroy(dmat);
		throw std::invalid_argument(msg);
	}
}

B200Preconditioner::~B200Preconditioner()
{
	b200_prec_destroy(dprec);
	b200_mat_destroy(dmat);
}

bool B200Preconditioner::relaxationAvailable() const
{
	return b200_prec_relaxation_available(dprec) != 0;
}

PrecInfo B200Preconditioner::compute()
{
	// the caller (PETSc) has rewritten the wrapped values in place: upload and factorise in one
	// pipelined call (layout conversion and initial guess behind the copies)
	PrecInfo info;
	double f[6];
	check_runtime(b200_prec_compute_host(dprec, mat.vals, f));
	for(int i = 0; i < 6; i++) info.f_info[i] = f[i];
	return info;
}

void B200Preconditioner::apply(const double *const r, double *const __restrict z) const
{
	check_runtime(b200_prec_apply_host(dprec, r, z));
}

void B200Preconditioner::apply_relax(const double *const b, double *const __restrict x) const
{
	b200_prec_set_apply_params(dprec, solveparams.rtol, solveparams.atol, solveparams.dtol,
	                           solveparams.ctol, solveparams.maxits);
	check_runtime(b200_prec_apply_relax_host(dprec, b, x, solveparams.maxits));
}

blasted::SRPreconditioner<double,int>*
B200Factory::create_preconditioner(SRMatrixStorage<const double,const int>&& prec_matrix,
                                   const SolverSettings& settings) const
{
	const AsyncSolverSettings& opts = dynamic_cast<const AsyncSolverSettings&>(settings);
	return new B200Preconditioner(std::move(prec_matrix), opts);
}

BlastedSolverType B200Factory::solverTypeFromString(const std::string precstr) const
{
	// same strings as the reference (include/solverfactory.hpp:22-43)
	return blasted::SRFactory<double,int>().solverTypeFromString(precstr);
}

B200MatrixView::B200MatrixView(const int n_brows, const int *const brptrs, const int *const bcinds,
                               const double *const values, const int *const dinds,
                               const int block_size, const bool rowmajor)
	: blasted::SRMatrixView<double,int>(n_brows, brptrs, bcinds, values, dinds, block_size,
	                                    block_size == 1 ? blasted::VIEWCSR : blasted::VIEWBSR),
	  bs{block_size}, dmat{nullptr}
{
	if(b200_mat_create_host(n_brows, block_size, rowmajor ? B200_ROWMAJOR : B200_COLMAJOR, brptrs,
	                        bcinds, values, dinds, &dmat))
		throw std::invalid_argument(b200_last_error());
}

B200MatrixView::~B200MatrixView() { b200_mat_destroy(dmat); }

void B200MatrixView::apply(const double *const x, double *const __restrict y) const
{
	check_runtime(b200_mat_apply_host(dmat, x, y));
}

void B200MatrixView::gemv3(const double a, const double *const __restrict x, const double b,
                           const double *const y, double *const z) const
{
	check_runtime(b200_mat_gemv3_host(dmat, a, x, b, y, z));
}


// ---- reordering / scaling on the device behind the reference's ReorderingScaling interface

namespace {

void check_b200(const int rc)
{
	if(rc) throw std::runtime_error(b200_last_error());
}

/// host matrix -> device, transform, -> the same host arrays
template <typename F>
void on_device_matrix(blasted::RawBSRMatrix<double,int>& mat, const int bs, F&& transform)
{
	b200_mat *dm = nullptr;
	// blocks are moved / scaled as opaque bs*bs chunks: the layout flag does not matter here
	check_b200(b200_mat_create_host(mat.nbrows, bs, B200_COLMAJOR, mat.browptr, mat.bcolind, mat.vals,
	                                nullptr, &dm));
	try {
		transform(dm);
		check_b200(b200_mat_get_host(dm, mat.browptr, mat.bcolind, mat.diagind, mat.vals));
	} catch(...) { b200_mat_destroy(dm); throw; }
	b200_mat_destroy(dm);
}

}

template <int bs>
void B200ReorderingScaling<bs>::setScaling(const double *const rscale, const double *const cscale,
                                           const int length)
{
	if(rscale) rowscale.assign(rscale, rscale + length);
	if(cscale) colscale.assign(cscale, cscale + length);
}

template <int bs>
void B200ReorderingScaling<bs>::applyOrdering(blasted::RawBSRMatrix<double,int>& mat,
                                              const blasted::RSApplyMode mode) const
{
	if(rp.empty() && cp.empty()) return;
	on_device_matrix(mat, bs, [&](b200_mat *dm) {
		check_b200(b200_mat_reorder(dm, rp.empty() ? nullptr : rp.data(), cp.empty() ? nullptr : cp.data(),
		                            mode == blasted::INVERSE ? 1 : 0, 0));
	});
}

template <int bs>
void B200ReorderingScaling<bs>::applyOrdering(double *const vec, const blasted::RSApplyMode mode,
                                              const blasted::RSApplyDir dir) const
{
	const std::vector<int>& ord = (dir == blasted::ROW) ? rp : cp;
	if(ord.empty()) return;
	check_b200(b200_vec_reorder(vec, (long long)ord.size(), bs, ord.data(),
	                            mode == blasted::INVERSE ? 1 : 0, 0));
}

template <int bs>
void B200ReorderingScaling<bs>::applyScaling(blasted::RawBSRMatrix<double,int>& mat,
                                             const blasted::RSApplyMode mode) const
{
	if(rowscale.empty() && colscale.empty()) return;
	on_device_matrix(mat, bs, [&](b200_mat *dm) {
		check_b200(b200_mat_scale(dm, rowscale.empty() ? nullptr : rowscale.data(),
		                          colscale.empty() ? nullptr : colscale.data(),
		                          mode == blasted::INVERSE ? 1 : 0, 0));
	});
}

template <int bs>
void B200ReorderingScaling<bs>::applyScaling(double *const vec, const blasted::RSApplyMode mode,
                                             const blasted::RSApplyDir dir) const
{
	const std::vector<double>& sc = (dir == blasted::ROW) ? rowscale : colscale;
	if(sc.empty()) return;
	check_b200(b200_vec_scale(vec, (long long)sc.size(), bs, sc.data(),
	                          mode == blasted::INVERSE ? 1 : 0, 0));
}

// the block sizes the reference instantiates (src/reorderingscaling.cpp:268-270, 370-372)
template class B200ReorderingScaling<1>;
template class B200ReorderingScaling<4>;
template class B200ReorderingScaling<7>;

}
