/** \file scalar_ilu.cu
 * \brief Scalar (bs = 1) fine-grained asynchronous ILU(0) factorisation on split storage (K1, K4, K10).
 *
 * Replaces async_ilu0_factorize_kernel / executeILU0Factorization
 * (src/kernels/kernels_ilu0_factorize.hpp:19-53, src/async_ilu_factor.cpp:154-177 of the reference),
 * the scalar initialisations (async_ilu_factor.cpp:47-58,110-151) and scalar_ilu0_nonlinear_res
 * (:180-217).
 *
 * Device storage of the factor: strict lower entries `lval`, strict upper entries `uval` (both in
 * row order) and the diagonal `udiag`, with work lists built once by pattern.cu.  One thread per
 * list item: every load of the lists, of A and of the result is unit-stride across the warp, all
 * lanes are active, and the only gathers are the partner entries of the products and u_jj - the
 * latter from the compact diagonal array, i.e. contiguous for neighbouring rows.
 * A sweep is a lower launch (l_ij = (a_ij - sum l_ik u_kj)/u_jj) followed by an upper launch
 * (u_ij = a_ij - sum l_ik u_kj); upper entries without products satisfy u_ij = a_ij identically and
 * are only touched when the initial guess did not already set them.
 * Chaotic-iteration rules as in the reference: one final store per entry, relaxed reads.
 *
 * Algorithmic bytes per sweep: lower 16+8+8 B per lower entry (+8 N for u_jj) + 24 B per product;
 * upper 16+8+8 B per work entry + 24 B per product.
 */
#include "common.cuh"

namespace b200 {

namespace {

enum { SM_SWEEP = 0, SM_RESIDUAL = 1, SM_INIT_ORIG = 2, SM_INIT_SGS = 3, SM_GATHER = 4 };

__device__ __forceinline__ double ld_iter(const double *p) { return __ldcg(p); }

__device__ __forceinline__ void block_sum_atomic(double v, double *out)
{
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
	__shared__ double wsum[8];
	if((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
	__syncthreads();
	if(threadIdx.x == 0) {
		double t = 0;
		for(int i = 0; i < (int)(blockDim.x >> 5); i++) t += wsum[i];
		if(t != 0) atomicAdd(out, t);
	}
}

/// Items per thread: the kernels are latency-bound (16-18 registers, one short dependent chain per
/// item), so every thread works on IPT items 256 apart and requests their loads together.
#ifndef B200_IPT
#define B200_IPT 2
#endif
constexpr int IPT = B200_IPT;

template <bool SCALE, int MODE>
__global__ void __launch_bounds__(256)
scalar_lower_kernel(const long long n, const int4 *__restrict__ lmeta,
                    const int *__restrict__ browind, const int *__restrict__ diagind,
                    const double *__restrict__ avals, const double *__restrict__ scale,
                    const int2 *__restrict__ spairs, double *lval, const double *uval,
                    const double *udiag, double *__restrict__ resout, int *__restrict__ changed,
                    const double *__restrict__ asplit, double *__restrict__ aout)
{
	// asplit: the (scaled) entries of A in list order, written through `aout` by the initialisation
	// or a gather pass; the sweeps read it instead of gathering A by entry index, where 8 useful
	// bytes cost a 32-byte sector (ncu: 1.23x / 1.44x the algorithmic traffic on a 7-point matrix)
	const long long t0 = (long long)blockIdx.x*(blockDim.x*IPT) + threadIdx.x;
	double res = 0;
	int4 m[IPT];                                         // {entry, col, ps, pe}
	bool ok[IPT];
	double sum[IPT], ujj[IPT], old[IPT];
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		const long long t = t0 + u*blockDim.x;
		ok[u] = t < n;
		m[u] = ok[u] ? __ldg(lmeta + t) : make_int4(0, 0, 0, 0);
	}
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		const long long t = t0 + u*blockDim.x;
		sum[u] = 0; ujj[u] = 1; old[u] = 0;
		if(!ok[u]) continue;
		if(MODE == SM_SWEEP && asplit) sum[u] = __ldg(asplit + t);
		else {
			sum[u] = __ldg(avals + m[u].x);
			if(SCALE) {
				sum[u] *= __ldg(scale + __ldg(browind + m[u].x));
				sum[u] *= __ldg(scale + m[u].y);
			}
		}
		if(MODE == SM_SWEEP || MODE == SM_RESIDUAL) {
			ujj[u] = ld_iter(udiag + m[u].y);
			if(MODE == SM_RESIDUAL || changed) old[u] = ld_iter(lval + t);
		}
	}
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		const long long t = t0 + u*blockDim.x;
		if(!ok[u]) continue;
		if((MODE == SM_INIT_ORIG || MODE == SM_INIT_SGS || MODE == SM_GATHER) && aout) aout[t] = sum[u];
		if(MODE == SM_GATHER) continue;
		if(MODE == SM_INIT_ORIG) lval[t] = sum[u];
		else if(MODE == SM_INIT_SGS) {
			// L' = L D^-1 on the (scaled) matrix, async_ilu_factor.cpp:110-133 (the reference
			// indexes `scale` out of bounds there; the intended a_cc s_c s_c is used)
			const double dg = __ldg(avals + __ldg(diagind + m[u].y));
			const double sc = SCALE ? __ldg(scale + m[u].y) : 1.0;
			lval[t] = sum[u] * (SCALE ? 1.0/(dg*sc*sc) : 1.0/dg);
		}
		else {
			double sm = sum[u];
			for(int k = m[u].z; k < m[u].w; k++) {
				const int2 pr = __ldg(spairs + k);
				sm = fma(-ld_iter(lval + pr.x), ld_iter(uval + pr.y), sm);
			}
			if(MODE == SM_RESIDUAL) res += fabs(sm - old[u]*ujj[u]);
			else {
				const double out = sm/ujj[u];
				if(changed && __double_as_longlong(old[u]) != __double_as_longlong(out)) *changed = 1;   // bitwise
				lval[t] = out;                            // single final store
			}
		}
	}
	if(MODE == SM_RESIDUAL) block_sum_atomic(res, resout);
}

template <bool SCALE, int MODE>
__global__ void __launch_bounds__(256)
scalar_upper_kernel(const long long n, const int4 *__restrict__ ulist,
                    const int *__restrict__ browind, const int *__restrict__ bcolind,
                    const double *__restrict__ avals, const double *__restrict__ scale,
                    const int2 *__restrict__ spairs, const double *lval, double *uval,
                    double *udiag, double *__restrict__ resout, int *__restrict__ changed,
                    const double *__restrict__ asplit, double *__restrict__ aout)
{
	const long long t0 = (long long)blockIdx.x*(blockDim.x*IPT) + threadIdx.x;
	double res = 0;
	int4 m[IPT];                                         // {entry, ps, pe, dest}
	bool ok[IPT];
	double sum[IPT];
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		const long long t = t0 + u*blockDim.x;
		ok[u] = t < n;
		m[u] = ok[u] ? __ldg(ulist + t) : make_int4(0, 0, 0, 0);
	}
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		sum[u] = 0;
		if(!ok[u]) continue;
		if(MODE == SM_SWEEP && asplit) sum[u] = __ldg(asplit + t0 + u*blockDim.x);
		else {
			sum[u] = __ldg(avals + m[u].x);
			if(SCALE) {
				sum[u] *= __ldg(scale + __ldg(browind + m[u].x));
				sum[u] *= __ldg(scale + __ldg(bcolind + m[u].x));
			}
		}
	}
#pragma unroll
	for(int u = 0; u < IPT; u++) {
		if(!ok[u]) continue;
		if(MODE == SM_GATHER) { aout[t0 + u*blockDim.x] = sum[u]; continue; }
		double *dst = (m[u].w < 0) ? udiag + (~m[u].w) : uval + m[u].w;
		if(MODE == SM_INIT_ORIG || MODE == SM_INIT_SGS) *dst = sum[u];
		else {
			// (requesting the pairs and factor entries of four products at once was measured: +7 % on
			// the 27-point lower launch, -7 % on its upper launch, nothing on 7-point - not kept)
			double sm = sum[u];
			for(int k = m[u].y; k < m[u].z; k++) {
				const int2 pr = __ldg(spairs + k);
				sm = fma(-ld_iter(lval + pr.x), ld_iter(uval + pr.y), sm);
			}
			if(MODE == SM_RESIDUAL) res += fabs(sm - ld_iter(dst));
			else {
				if(changed && __double_as_longlong(ld_iter(dst)) != __double_as_longlong(sm)) *changed = 1;
				*dst = sm;
			}
		}
	}
	if(MODE == SM_RESIDUAL) block_sum_atomic(res, resout);
}

__global__ void __launch_bounds__(256)
scalar_gather_kernel(const long long nlower, const long long nupper, const int4 *__restrict__ lmeta,
                     const int4 *__restrict__ uall, const double *__restrict__ lval,
                     const double *__restrict__ uval, const double *__restrict__ udiag,
                     double *__restrict__ out)
{
	const long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(t < nlower) out[lmeta[t].x] = lval[t];
	else if(t < nlower + nupper) {
		const int4 m = uall[t - nlower];
		out[m.x] = (m.w < 0) ? udiag[~m.w] : uval[m.w];
	}
}


// ------------------------------------------------------------------ exact factorisation, one launch

__device__ __forceinline__ int ld_poll_int(const int *p)
{
	int v;
	asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

/// Exact ILU(0) in ONE launch: one thread per row, rows taken in level-sorted order (CTA positions
/// from a ticket counter, as apply.cu::tri_syncfree_kernel), a row waits until every row named by
/// its lower part has raised its flag and then computes its entries in storage order - exactly the
/// sequential pass of the reference's kernel (tests/solverops/async_ilu_convergence.cpp:462-490),
/// with the per-entry arithmetic of the sweep kernels above, so the result is bit-identical to the
/// fixed point the sweeps iterate to.  While a warp waits only its first unfinished lane polls.
template <bool SCALE>
__global__ void __launch_bounds__(256)
scalar_exact_kernel(const int nrows, const int *__restrict__ rows, const int *__restrict__ lptr,
                    const int *__restrict__ uptr, const int *__restrict__ lcol,
                    const int4 *__restrict__ lmeta, const int4 *__restrict__ uall,
                    const int *__restrict__ browind, const int *__restrict__ bcolind,
                    const double *__restrict__ avals, const double *__restrict__ scale,
                    const int2 *__restrict__ spairs, double *lval, double *uval, double *udiag,
                    int *rowdone, int *__restrict__ ticket, int *__restrict__ err)
{
	__shared__ int s_cta;
	if(threadIdx.x == 0) s_cta = atomicAdd(ticket, 1);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const long long t = (long long)s_cta*blockDim.x + threadIdx.x;
	const bool valid = t < nrows;
	int row = 0, ls = 0, le = 0, us = 0, ue = 0;
	if(valid) {
		row = __ldg(rows + t);
		ls = __ldg(lptr + row); le = __ldg(lptr + row + 1);
		us = __ldg(uptr + row) + row; ue = __ldg(uptr + row + 1) + row + 1;
	}
	int kdep = ls;
	bool done = !valid;
	int spins = 0;
	while(true) {
		const unsigned waiting = __ballot_sync(0xffffffffu, !done);
		if(!waiting) break;
		const int leader = __ffs(waiting) - 1;
		if(lane == leader)
			while(kdep < le && ld_poll_int(rowdone + __ldg(lcol + kdep)) != 0) kdep++;
		const int go = __shfl_sync(0xffffffffu, (int)(kdep == le), leader);
		if(!go) {
			__nanosleep(64);
			if(++spins > (1 << 16) && ((spins & 1023) == 0)) {
				if(spins > (1 << 22) || *((volatile int*)err)) { *err = 1; kdep = le; }
			}
			continue;
		}
		if(!done) {
			while(kdep < le && ld_poll_int(rowdone + __ldg(lcol + kdep)) != 0) kdep++;
			if(kdep == le) {
				__threadfence();
				for(int i = ls; i < le; i++) {
					const int4 m = __ldg(lmeta + i);                 // {entry, col, ps, pe}
					double sum = __ldg(avals + m.x);
					if(SCALE) {
						sum *= __ldg(scale + __ldg(browind + m.x));
						sum *= __ldg(scale + m.y);
					}
					for(int k = m.z; k < m.w; k++) {
						const int2 pr = __ldg(spairs + k);
						sum = fma(-ld_iter(lval + pr.x), ld_iter(uval + pr.y), sum);
					}
					lval[i] = sum/ld_iter(udiag + m.y);
				}
				for(int i = us; i < ue; i++) {
					const int4 m = __ldg(uall + i);                  // {entry, ps, pe, dest}
					double sum = __ldg(avals + m.x);
					if(SCALE) {
						sum *= __ldg(scale + __ldg(browind + m.x));
						sum *= __ldg(scale + __ldg(bcolind + m.x));
					}
					for(int k = m.y; k < m.z; k++) {
						const int2 pr = __ldg(spairs + k);
						sum = fma(-ld_iter(lval + pr.x), ld_iter(uval + pr.y), sum);
					}
					if(m.w < 0) udiag[~m.w] = sum; else uval[m.w] = sum;
				}
				__threadfence();
				*((volatile int*)(rowdone + row)) = 1;
				done = true;
			}
		}
	}
}

}  // namespace

void scalar_ilu0_exact(const Mat& A, const IluPattern& pl, const int *level_rows, const double *scale,
                       ScalarFactor& F, int *rowdone, int *flags, cudaStream_t st)
{
	const int n = A.nbrows;
	if(n == 0) return;
	ProfScope ps(KC_FACTOR_LOWER, st);
	B200_CUDA(cudaMemsetAsync(rowdone, 0, (size_t)n*sizeof(int), st));
	B200_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));                // ticket
	const int grid = div_up(n, 256);
	if(scale)
		scalar_exact_kernel<true><<<grid, 256, 0, st>>>(n, level_rows, pl.lptr, pl.uptr, pl.lcol,
			pl.slmeta, pl.suall, A.browind, A.bcolind, A.vals, scale, pl.spairs, F.lval.p, F.uval.p,
			F.udiag.p, rowdone, flags, flags + 1);
	else
		scalar_exact_kernel<false><<<grid, 256, 0, st>>>(n, level_rows, pl.lptr, pl.uptr, pl.lcol,
			pl.slmeta, pl.suall, A.browind, A.bcolind, A.vals, scale, pl.spairs, F.lval.p, F.uval.p,
			F.udiag.p, rowdone, flags, flags + 1);
	B200_LAUNCHED();
}

namespace {

template <int MODE>
void run_lower(const Mat& A, const IluPattern& pl, const double *scale, const ScalarFactor& F,
               double *res, int *changed, cudaStream_t st, const double *asplit = nullptr,
               double *aout = nullptr)
{
	if(pl.nlower == 0) return;
	const int grid = div_up(pl.nlower, 256*IPT);
	if(scale)
		scalar_lower_kernel<true,MODE><<<grid,256,0,st>>>(pl.nlower, pl.slmeta, A.browind, A.diagind,
			A.vals, scale, pl.spairs, F.lval.p, F.uval.p, F.udiag.p, res, changed, asplit, aout);
	else
		scalar_lower_kernel<false,MODE><<<grid,256,0,st>>>(pl.nlower, pl.slmeta, A.browind, A.diagind,
			A.vals, scale, pl.spairs, F.lval.p, F.uval.p, F.udiag.p, res, changed, asplit, aout);
	B200_LAUNCHED();
}

template <int MODE>
void run_upper(const Mat& A, const IluPattern& pl, const double *scale, const ScalarFactor& F,
               bool all, double *res, int *changed, cudaStream_t st, const double *asplit = nullptr,
               double *aout = nullptr)
{
	const long long n = all ? pl.nupper : pl.nuwork;
	const int4 *list = all ? pl.suall.p : pl.suwork.p;
	if(n == 0) return;
	const int grid = div_up(n, 256*IPT);
	if(scale)
		scalar_upper_kernel<true,MODE><<<grid,256,0,st>>>(n, list, A.browind, A.bcolind, A.vals, scale,
			pl.spairs, F.lval.p, F.uval.p, F.udiag.p, res, changed, asplit, aout);
	else
		scalar_upper_kernel<false,MODE><<<grid,256,0,st>>>(n, list, A.browind, A.bcolind, A.vals, scale,
			pl.spairs, F.lval.p, F.uval.p, F.udiag.p, res, changed, asplit, aout);
	B200_LAUNCHED();
}

}  // namespace

void scalar_ilu0_init(const Mat& A, const IluPattern& pl, const double *scale, int fact_init,
                      ScalarFactor& F, cudaStream_t st)
{
	F.lval.alloc(std::max<long long>(pl.nlower, 1));
	F.uval.alloc(std::max<long long>(pl.nstrict, 1));
	F.udiag.alloc(std::max(A.nbrows, 1));
	// the (scaled) entries of A in list order for the sweeps: written by the lower initialisation
	// launch itself, by a gather over the upper work list, or - no initialisation - by gathers
	F.alow.alloc(std::max<long long>(pl.nlower, 1));
	F.aupw.alloc(std::max<long long>(pl.nuwork, 1));
	ProfScope ps(KC_FACTOR_INIT, st);
	if(fact_init == B200_INIT_F_NONE)
		run_lower<SM_GATHER>(A, pl, scale, F, nullptr, nullptr, st, nullptr, F.alow.p);
	// INIT_F_ZERO falls through into INIT_F_ORIGINAL in the scalar reference
	// (src/async_ilu_factor.cpp:48-54, missing break): replicated
	else if(fact_init == B200_INIT_F_SGS) {
		run_lower<SM_INIT_SGS>(A, pl, scale, F, nullptr, nullptr, st, nullptr, F.alow.p);
		run_upper<SM_INIT_SGS>(A, pl, scale, F, true, nullptr, nullptr, st);
	} else {
		run_lower<SM_INIT_ORIG>(A, pl, scale, F, nullptr, nullptr, st, nullptr, F.alow.p);
		run_upper<SM_INIT_ORIG>(A, pl, scale, F, true, nullptr, nullptr, st);
	}
	run_upper<SM_GATHER>(A, pl, scale, F, false, nullptr, nullptr, st, nullptr, F.aupw.p);
}

void scalar_ilu0_sweep(const Mat& A, const IluPattern& pl, const double *scale, ScalarFactor& F,
                       int *d_changed, bool all_upper, cudaStream_t st)
{
	{ ProfScope ps(KC_FACTOR_LOWER, st); run_lower<SM_SWEEP>(A, pl, scale, F, nullptr, d_changed, st, F.alow.p); }
	{
		ProfScope ps(KC_FACTOR_UPPER, st);
		// the split copy of A follows the work list; the (rare) pass over all upper entries reads A
		run_upper<SM_SWEEP>(A, pl, scale, F, all_upper, nullptr, d_changed, st,
		                    all_upper ? nullptr : F.aupw.p);
	}
}

double scalar_ilu0_residual(const Mat& A, const IluPattern& pl, const double *scale,
                            const ScalarFactor& F, double *d_scratch, cudaStream_t st)
{
	B200_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(double), st));
	run_lower<SM_RESIDUAL>(A, pl, scale, F, d_scratch, nullptr, st);
	run_upper<SM_RESIDUAL>(A, pl, scale, F, true, d_scratch, nullptr, st);
	double r = 0;
	B200_CUDA(cudaMemcpyAsync(&r, d_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return r;
}

void scalar_ilu0_gather(const Mat& A, const IluPattern& pl, const ScalarFactor& F, double *out,
                        cudaStream_t st)
{
	const long long n = pl.nlower + pl.nupper;
	if(n == 0) return;
	scalar_gather_kernel<<<div_up(n, 256), 256, 0, st>>>(pl.nlower, pl.nupper, pl.slmeta, pl.suall,
	                                                     F.lval, F.uval, F.udiag, out);
	B200_LAUNCHED();
}

}  // namespace b200
