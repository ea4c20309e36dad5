/** \file b200_solverops.hpp
 * \brief C++ adapters that plug the B200 C ABI (include/blasted_b200.h) in behind BLASTed's own
 * host interface: a maintainer of the reference adds these two files to the build, links
 * libblasted_b200.so, and passes a B200Factory where an SRFactory was used.
 *
 * The classes derive from the REFERENCE's own types (they need the reference's include directory):
 *   B200Preconditioner : blasted::SRPreconditioner<double,int>   include/solverops_base.hpp:68-78
 *   B200Factory        : blasted::FactoryBase<double,int>        include/solverfactory.hpp:71-85
 *   B200MatrixView     : blasted::SRMatrixView<double,int>       include/blockmatrices.hpp:27-60
 *   B200ReorderingScaling<bs> : blasted::ReorderingScaling<double,int,bs>  include/reorderingscaling.hpp:106-133
 * so the PETSc PCSHELL glue (src/blasted_petsc.cpp:216-327, 474-576), tests/testsolve.cpp:86-88 and the
 * Krylov drivers of tests/solvers.cpp use them unchanged.  Pointers handed to apply()/apply_relax()
 * are HOST pointers, as in the reference; copies to and from the device happen inside the call.
 * Error conventions follow the reference: std::invalid_argument from the factory
 * (src/solverfactory.cpp:122,203-206,220-227), std::runtime_error from the objects
 * (src/solverops_ilu0.cpp:126,215).
 */
#ifndef BLASTED_B200_SOLVEROPS_H
#define BLASTED_B200_SOLVEROPS_H

#include "solverfactory.hpp"
#include "solverops_base.hpp"
#include "blockmatrices.hpp"
#include "reorderingscaling.hpp"
#include "../../include/blasted_b200.h"

namespace blasted_b200 {

using blasted::SRMatrixStorage;
using blasted::SRPreconditioner;
using blasted::SolverSettings;
using blasted::AsyncSolverSettings;
using blasted::PrecInfo;

/// Device-resident stand-in for every SRPreconditioner subclass the SRFactory can create
class B200Preconditioner : public SRPreconditioner<double,int>
{
public:
	B200Preconditioner(SRMatrixStorage<const double,const int>&& matrix,
	                   const AsyncSolverSettings& settings);
	~B200Preconditioner();

	int dim() const { return mat.nbrows*bs; }
	bool relaxationAvailable() const;

	/// Uploads the matrix's current values (the caller may have rewritten them in place, pattern
	/// fixed - include/solverops_ilu0.hpp:53-56) and rebuilds the preconditioner on the device
	PrecInfo compute();
	void apply(const double *const r, double *const __restrict z) const;
	void apply_relax(const double *const b, double *const __restrict x) const;

	b200_prec *handle() const { return dprec; }

protected:
	using SRPreconditioner<double,int>::mat;
	using SRPreconditioner<double,int>::solveparams;
	int bs;
	b200_mat *dmat;
	b200_prec *dprec;
};

/// Drop-in for blasted::SRFactory<double,int>
class B200Factory : public blasted::FactoryBase<double,int>
{
public:
	SRPreconditioner<double,int>*
	create_preconditioner(SRMatrixStorage<const double,const int>&& prec_matrix,
	                      const SolverSettings& settings) const;
	BlastedSolverType solverTypeFromString(const std::string precstr) const;
};

/// Device-resident operator behind the reference's matrix-view interface (apply / gemv3 / dim)
class B200MatrixView : public blasted::SRMatrixView<double,int>
{
public:
	B200MatrixView(const int n_brows, const int *const brptrs, const int *const bcinds,
	               const double *const values, const int *const dinds, const int block_size,
	               const bool rowmajor);
	~B200MatrixView();
	void apply(const double *const x, double *const __restrict y) const;
	void gemv3(const double a, const double *const __restrict x, const double b,
	           const double *const y, double *const z) const;
	int dim() const { return mat.nbrows*bs; }
	b200_mat *handle() const { return dmat; }
protected:
	using blasted::SRMatrixView<double,int>::mat;
	int bs;
	b200_mat *dmat;
};

/// Reordering / scaling of host matrices and vectors carried out on the device: same virtuals,
/// same conventions as the reference classes (include/reorderingscaling.hpp:36-133,
/// src/reorderingscaling.cpp:77-368).  As in the reference the orderings and scalings themselves come
/// from outside (setOrdering, setScaling); compute() is a no-op.  One difference: the matrix'
/// diagind array, which the reference leaves stale after a permutation, is located again.
template <int bs>
class B200ReorderingScaling : public blasted::ReorderingScaling<double,int,bs>
{
public:
	void compute(const blasted::CRawBSRMatrix<double,int>&) override { }
	void setScaling(const double *const rscale, const double *const cscale, const int length);
	void applyOrdering(blasted::RawBSRMatrix<double,int>& mat, const blasted::RSApplyMode mode) const override;
	void applyOrdering(double *const vec, const blasted::RSApplyMode mode,
	                   const blasted::RSApplyDir dir) const override;
	void applyScaling(blasted::RawBSRMatrix<double,int>& mat, const blasted::RSApplyMode mode) const override;
	void applyScaling(double *const vec, const blasted::RSApplyMode mode,
	                  const blasted::RSApplyDir dir) const override;
protected:
	using blasted::ReorderingScaling<double,int,bs>::rp;
	using blasted::ReorderingScaling<double,int,bs>::cp;
	using blasted::ReorderingScaling<double,int,bs>::rowscale;
	using blasted::ReorderingScaling<double,int,bs>::colscale;
};

}
#endif
