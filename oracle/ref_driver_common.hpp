/** \file ref_driver_common.hpp
 * \brief Helpers shared by oracle/ref_driver.cpp (the unmodified reference behind a C ABI) and
 * oracle/adapter_driver.cpp (the product's C++ adapters behind the same handles).
 * TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORACLE_REF_DRIVER_COMMON_HPP
#define ORACLE_REF_DRIVER_COMMON_HPP

#include <string>
#include <sstream>
#include <iostream>

#include "solverfactory.hpp"
#include "reorderingscaling.hpp"

namespace refdrv {

using namespace blasted;

typedef SRMatrixStorage<const double, const int> CStorage;

inline CStorage wrap(const int nbrows, const int *browptr, const int *bcolind, const double *vals,
              const int *diagind, const int bs)
{
	// same wrapping as src/blasted_petsc.cpp:285-297
	return CStorage(browptr, bcolind, vals, diagind, browptr+1, nbrows, browptr[nbrows],
	                browptr[nbrows], bs);
}

struct RefPrec {
	SRPreconditioner<double,int> *p;
	int bs;
};


// silence the reference's chatter on stdout while a call is in flight
struct CoutMute {
	std::streambuf *old;
	std::ostringstream sink;
	CoutMute() { old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutMute() { std::cout.rdbuf(old); }
};

// front end: the reference's Reordering/ReorderingScaling are abstract (compute() comes from an
// external ordering package); this subclass only lets the driver set the vectors they apply
template <int bs>
struct SetOrderingScaling : public ReorderingScaling<double,int,bs> {
	void compute(const CRawBSRMatrix<double,int>&) override { }
	void setScaling(const double *rs, const double *cs, const int n) {
		if(rs) this->rowscale.assign(rs, rs + n);
		if(cs) this->colscale.assign(cs, cs + n);
	}
};

template <int bs, typename RS = SetOrderingScaling<bs>>
void ref_reorder_scale(int nbrows, int *browptr, int *bcolind, double *vals, int *diagind,
                       const int *rord, const int *cord, const double *rs, const double *cs,
                       int inverse, double *rowvec, double *colvec)
{
	RS r;
	r.setOrdering(rord, cord, nbrows);
	r.setScaling(rs, cs, nbrows);
	const RSApplyMode mode = inverse ? INVERSE : FORWARD;
	if(browptr) {
		RawBSRMatrix<double,int> mat(browptr, bcolind, vals, diagind, browptr+1, nbrows,
		                             browptr[nbrows], browptr[nbrows]);
		if(rs || cs) r.applyScaling(mat, mode);
		if(rord || cord) r.applyOrdering(mat, mode);
	}
	if(rowvec) {
		if(rs) r.applyScaling(rowvec, mode, ROW);
		if(rord) r.applyOrdering(rowvec, mode, ROW);
	}
	if(colvec) {
		if(cs) r.applyScaling(colvec, mode, COLUMN);
		if(cord) r.applyOrdering(colvec, mode, COLUMN);
	}
}


}  // namespace refdrv
#endif
