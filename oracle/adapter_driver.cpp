/** \file adapter_driver.cpp
 * \brief C entry points that put the PRODUCT's C++ adapters (blasted_b200/host: B200Factory,
 * B200Preconditioner, B200ReorderingScaling) behind the same handles as oracle/ref_driver.cpp, so
 * that the reference's own, unmodified Krylov drivers and front-end classes run on the device
 * implementation (tests/test_gpu_adapter.py, tests/test_gpu_frontend.py).
 *
 * TEST INFRASTRUCTURE.  Built into oracle/_ref/libb200_adapters.so, which links libblasted_ref.so
 * (the reference's base classes) and libblasted_b200.so.  Kept apart from libblasted_ref.so so that
 * the reference arm of bench.py loads nothing of the product.
 */
#include "ref_driver_common.hpp"
#include "../blasted_b200/host/b200_solverops.hpp"

using namespace blasted;
using namespace refdrv;

extern "C" void ref_set_error(const char *msg);      // libblasted_ref.so

extern "C" {

/// Same as ref_prec_create (ref_driver.cpp) but through the product's B200Factory: the returned
/// object is a device preconditioner behind the reference's SRPreconditioner interface, usable by
/// every entry point of ref_driver.cpp (compute/apply/solve with the reference's own Krylov code).
void *ref_prec_create_b200(const char *prectype, int bs, int rowmajor, int scale,
                           int nbuildsweeps, int napplysweeps, int fact_init, int apply_init,
                           int thread_chunk_size, int compute_precinfo,
                           int nbrows, const int *browptr, const int *bcolind, const double *vals,
                           const int *diagind)
{
	try {
		blasted_b200::B200Factory fact;
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
		const FactoryBase<double,int>& f = fact;      // through the abstract factory seam
		h->p = f.create_preconditioner(wrap(nbrows, browptr, bcolind, vals, diagind, bs), s);
		return h;
	} catch(std::exception& e) {
		ref_set_error(e.what());
		return nullptr;
	}
}

/// ref_reorder_scale (ref_driver.cpp) through the product's device implementation behind the
/// reference's own interface (B200ReorderingScaling<bs> : ReorderingScaling<double,int,bs>)
int ref_reorder_scale_b200(int bs, int nbrows, int *browptr, int *bcolind, double *vals, int *diagind,
                           const int *rord, const int *cord, const double *rowscale,
                           const double *colscale, int inverse, double *rowvec, double *colvec)
{
	using namespace blasted_b200;
	try {
		if(bs == 1) ref_reorder_scale<1,B200ReorderingScaling<1>>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else if(bs == 4) ref_reorder_scale<4,B200ReorderingScaling<4>>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else if(bs == 7) ref_reorder_scale<7,B200ReorderingScaling<7>>(nbrows, browptr, bcolind, vals, diagind, rord, cord, rowscale, colscale, inverse, rowvec, colvec);
		else { ref_set_error("Reordering: only bs 1,4,7 instantiated in the reference"); return 1; }
	} catch(std::exception& e) { ref_set_error(e.what()); return 1; }
	return 0;
}

}
