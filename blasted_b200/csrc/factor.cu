/** \file factor.cu
 * \brief Fine-grained asynchronous ILU(0) factorisation on the device (K1, K2, K3, K4, K10).
 *
 * Replaces, from the reference:
 *   async_ilu0_factorize_kernel / executeILU0Factorization   src/kernels/kernels_ilu0_factorize.hpp:19-53,
 *                                                            src/async_ilu_factor.cpp:154-177
 *   async_block_ilu0_factorize / async_bilu0_sweeps          kernels_ilu0_factorize.hpp:71-98,
 *                                                            src/async_blockilu_factor.cpp:187-204
 *   initialisations                                          async_ilu_factor.cpp:47-58,110-151;
 *                                                            async_blockilu_factor.cpp:63-94,207-254
 *   diagonal block inversion                                 async_blockilu_factor.cpp:144-146,
 *                                                            src/solverops_jacobi.cpp:43-45,141-147
 *   nonlinear residual                                       async_ilu_factor.cpp:180-217,
 *                                                            async_blockilu_factor.cpp:257-297
 *   getScalingVector                                         src/rawsrmatrixutils.cpp:343-350
 *   diagonal_dominance                                       src/matrix_properties.cpp:11-77
 *
 * Chaotic-iteration semantics kept from the reference: every entry is recomputed from A and the
 * current iterate and written with ONE final store (never a partial sum,
 * kernels_ilu0_factorize.hpp:34-40); reads of other entries are unordered relaxed loads that may
 * observe old or new values.  One sweep = one pass over all stored entries; on the device a sweep
 * is issued as a lower-triangle launch followed by an upper-triangle launch, so U entries see the
 * L entries of the same sweep (the reference gets the same effect from processing a row's entries
 * in order).
 *
 * Mapping (blocks; the scalar factorisation is scalar_ilu.cu).  A group of bs lanes per stored
 * block ("entry-parallel", packed per-phase work lists from pattern.cu), lane r owns row r of the
 * block in registers; products stream the partner's rows through warp shuffles
 * (blockops.cuh::group_mul_sub); U_jj^-1 comes from the compact array `dinv`, which the upper launch
 * refreshes with a bs x bs Gauss-Jordan elimination with partial pivoting shared by the bs lanes of
 * the group (blockops.cuh::group_inverse).
 * All of it is HBM/L2-bound; algorithmic bytes per launch are listed in DESIGN.md section 3.
 */
#include "common.cuh"
#include "blockops.cuh"
#include "tma.cuh"
#include <cub/device/device_reduce.cuh>

namespace b200 {

// Loads of the iterate that other CTAs may be rewriting: go to L2 (ld.global.cg) so that values
// written earlier in the same launch are observed (L1 is not coherent across SMs).
__device__ __forceinline__ double ld_iter(const double *p) { return __ldcg(p); }

// ------------------------------------------------------------------ scaling vector

template <int BS>
__global__ void scaling_vector_kernel(const int nbrows, const double *__restrict__ vals,
                                      const int *__restrict__ diagind, double *__restrict__ scale)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= (long long)nbrows*BS) return;
	const int row = (int)(i / BS), j = (int)(i % BS);
	scale[i] = 1.0/sqrt(vals[(size_t)diagind[row]*BS*BS + j*BS + j]);
}

void launch_scaling_vector(const Mat& A, double *scale, cudaStream_t st)
{
	const long long n = (long long)A.nbrows*A.bs;
	if(n == 0) return;
	const int grid = div_up(n, 256);
	switch(A.bs) {
	case 1: scaling_vector_kernel<1><<<grid,256,0,st>>>(A.nbrows, A.vals, A.diagind, scale); break;
	case 4: scaling_vector_kernel<4><<<grid,256,0,st>>>(A.nbrows, A.vals, A.diagind, scale); break;
	case 5: scaling_vector_kernel<5><<<grid,256,0,st>>>(A.nbrows, A.vals, A.diagind, scale); break;
	default: throw Error("scaling: unsupported block size");
	}
	B200_LAUNCHED();
}

// (the scalar factorisation lives in scalar_ilu.cu)

// ------------------------------------------------------------------ block ILU(0): storage
//
// The block factor lives in SPLIT form (ScalarFactor, common.cuh): L part `lval` in row order (the
// order of the lower work list), diagonal blocks U_ii `udiag` (un-inverted; their inverses in the
// compact array `dinv`), strict upper part `uval` in row order.  Every launch then reads and writes
// contiguous arrays: the lower launch stores L with unit stride, the triangular sweeps stream
// their own part instead of skipping over the other one.
// OPTIONAL second copy `ut` of the strict upper part in COLUMN order (B200_UT=1, bs = 5): the
// products pair row i of L with column j of U, and for the diagonal entries of a structurally
// symmetric matrix the two runs are then index-aligned and contiguous.  Only `ut` is kept current
// during the sweeps in that mode; `uval` catches up once at the end (sync_upper).

enum { MODE_INIT_SGS = 2, MODE_INIT_ORIG = 3 };

/// Initial guess of point-block ILU(0) written straight into the split arrays.
/// A group of BS lanes per stored block; lane r holds row r of the block.
template <int BS, bool SCALE, int MODE>
__global__ void __launch_bounds__(256)
block_split_init_kernel(const long long entry0, const long long nnzb, const int *__restrict__ browptr,
                        const int *__restrict__ bcolind, const int *__restrict__ browind,
                        const int *__restrict__ diagind, const int *__restrict__ lptr,
                        const int *__restrict__ uptr, const int *__restrict__ utpos,
                        const double *__restrict__ avals, const double *__restrict__ scale,
                        double *__restrict__ lval, double *__restrict__ udiag,
                        double *__restrict__ uval, double *__restrict__ ut)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int g = lane / BS, r = lane - g*BS;
	const long long entry = entry0 + warp*GPW + g;           // entries [entry0, nnzb)
	if(g >= GPW || entry >= nnzb) return;
	const int row = __ldg(browind + entry), col = __ldg(bcolind + entry);
	const int dg = __ldg(diagind + row);
	double sum[BS];
	BlkIO<BS>::template load_row<false>(avals + (size_t)entry*BS2, r, sum);
	if(SCALE) {
		// scaleBlock: val(i,j) *= scale[brow*bs+i]*scale[bcol*bs+j]  (kernels_ilu0_factorize.hpp:61-69)
		const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
		for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)col*BS + c);
	}
	if(entry < dg) {
		double *op = lval + ((size_t)__ldg(lptr + row) + (entry - __ldg(browptr + row)))*BS2;
		if(MODE == MODE_INIT_SGS) {
			// L' = L D^-1 with D the (scaled) diagonal block of A: async_blockilu_factor.cpp:221-250
			double d[BS2], x[BS];
			BlkIO<BS>::template load_full<false>(avals + (size_t)__ldg(diagind + col)*BS2, d);
			if(SCALE) {
#pragma unroll
				for(int c = 0; c < BS; c++)
#pragma unroll
					for(int m = 0; m < BS; m++)
						d[c*BS+m] *= __ldg(scale + (size_t)col*BS + m)*__ldg(scale + (size_t)col*BS + c);
			}
			solve_right<BS>(d, sum, x);
			BlkIO<BS>::store_row(op, r, x);
		} else
			BlkIO<BS>::store_row(op, r, sum);
	}
	else if(entry == dg)
		BlkIO<BS>::store_row(udiag + (size_t)row*BS2, r, sum);
	else {
		const int ui = __ldg(uptr + row) + (int)(entry - dg - 1);
		BlkIO<BS>::store_row(uval + (size_t)ui*BS2, r, sum);
		if(utpos) BlkIO<BS>::store_row(ut + (size_t)__ldg(utpos + ui)*BS2, r, sum);
	}
}

/// Nonlinear residual sum |(A - LU)_S| over a factor in MATRIX order with un-inverted diagonal
/// blocks (assembled from the split form on request; diagnostics only).
/// A group of BS lanes per stored block; lane r holds row r of the block.
template <int BS, bool SCALE>
__global__ void __launch_bounds__(256)
block_ilu0_residual_kernel(const long long nnzb, const int *__restrict__ bcolind,
                  const int *__restrict__ browind, const int *__restrict__ diagind,
                  const double *__restrict__ avals, const int *__restrict__ posptr,
                  const int *__restrict__ lowerp, const int *__restrict__ upperp,
                  const double *__restrict__ scale, const double *__restrict__ ilu,
                  double *__restrict__ resout)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int g = lane / BS, r = lane - g*BS;
	const long long entry = warp*GPW + g;
	double res = 0;
	double sum[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) sum[c] = 0;
	const bool active = (g < GPW) && (entry < nnzb);
	if(active) {
		const int row = __ldg(browind + entry), col = __ldg(bcolind + entry);
		const bool lower = row > col;
		BlkIO<BS>::template load_row<false>(avals + (size_t)entry*BS2, r, sum);
		if(SCALE) {
			const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
			for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)col*BS + c);
		}
		const int ps = __ldg(posptr + entry), pe = __ldg(posptr + entry + 1);
		for(int k = ps; k < pe; k++) {
			double lr[BS], u[BS2];
			BlkIO<BS>::template load_row<false>(ilu + (size_t)__ldg(lowerp + k)*BS2, r, lr);
			BlkIO<BS>::template load_full<false>(ilu + (size_t)__ldg(upperp + k)*BS2, u);
#pragma unroll
			for(int c = 0; c < BS; c++)
#pragma unroll
				for(int m = 0; m < BS; m++)
					sum[c] = fma(-lr[m], u[c*BS+m], sum[c]);
		}
		double cur[BS];
		BlkIO<BS>::template load_row<false>(ilu + (size_t)entry*BS2, r, cur);
		if(lower) {
			double d[BS2];
			BlkIO<BS>::template load_full<false>(ilu + (size_t)__ldg(diagind + col)*BS2, d);
#pragma unroll
			for(int c = 0; c < BS; c++)
#pragma unroll
				for(int m = 0; m < BS; m++)
					sum[c] = fma(-cur[m], d[c*BS+m], sum[c]);
		} else {
#pragma unroll
			for(int c = 0; c < BS; c++) sum[c] -= cur[c];
		}
#pragma unroll
		for(int c = 0; c < BS; c++) res += fabs(sum[c]);
	}
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) res += __shfl_down_sync(0xffffffffu, res, off);
	__shared__ double wsum[8];
	const int w = threadIdx.x >> 5;
	if(lane == 0) wsum[w] = res;
	__syncthreads();
	if(threadIdx.x == 0) {
		double t = 0;
		for(int i = 0; i < (int)(blockDim.x >> 5); i++) t += wsum[i];
		atomicAdd(resout, t);
	}
}

/// out (matrix order) <- split factor; diagonal blocks from `dsrc` (U_ii or their inverses)
template <int BS>
__global__ void __launch_bounds__(256)
block_assemble_kernel(const long long nnzb, const int *__restrict__ browptr,
                      const int *__restrict__ browind, const int *__restrict__ diagind,
                      const int *__restrict__ lptr, const int *__restrict__ uptr,
                      const double *__restrict__ lval, const double *__restrict__ dsrc,
                      const double *__restrict__ uval, double *__restrict__ out)
{
	constexpr int BS2 = BS*BS;
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= nnzb*BS2) return;
	const long long entry = e / BS2;
	const int w = (int)(e - entry*BS2);
	const int row = browind[entry], dg = diagind[row];
	double v;
	if(entry < dg) v = lval[((size_t)lptr[row] + (entry - browptr[row]))*BS2 + w];
	else if(entry == dg) v = dsrc[(size_t)row*BS2 + w];
	else v = uval[((size_t)uptr[row] + (entry - dg - 1))*BS2 + w];
	out[e] = v;
}

/// uval[ut_order[d]] <- ut[d] for the strict upper entries of `list` (null: all nstrict blocks)
template <int BS>
__global__ void __launch_bounds__(256)
ut_to_uval_kernel(const long long nitems, const int4 *__restrict__ list,
                  const int *__restrict__ ut_order, const double *__restrict__ ut,
                  double *__restrict__ uval)
{
	constexpr int BS2 = BS*BS;
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= nitems*BS2) return;
	const long long t = e / BS2;
	const int w = (int)(e - t*BS2);
	int d = (int)t;
	if(list) { d = list[t].w; if(d < 0) return; }
	uval[(size_t)ut_order[d]*BS2 + w] = ut[(size_t)d*BS2 + w];
}


// ------------------------------------------------------------------ block ILU(0): sweep kernels
//
// The two launches of one sweep.  Work items come from the packed per-phase lists (IluPattern),
// a persistent grid walks them in ascending order with the metadata of the NEXT item prefetched
// while the current one is processed, so that only one memory latency (the block loads) is exposed
// per item instead of the index chain entry -> column -> diagonal position -> block.
//
// Lower entry (i,j), i>j:  L_ij = (A_ij - sum_k L_ik U_kj) U_jj^-1.  U_jj^-1 is read from the compact
// array `dinv`, which the upper launch of the previous sweep (or the initialisation) wrote from the
// very U_jj the reference would invert here (kernels_ilu0_factorize.hpp:91); during the lower launch
// nobody writes U_jj, so the value used is the same.  The result goes to lval[t], t the list index:
// unit-stride stores.
// Upper entry (i,j), i<=j: U_ij = A_ij - sum_k L_ik U_kj, written to ut (column order) or, for a
// diagonal entry, to udiag; a diagonal entry also refreshes dinv[i].
//
// (A single row-fused launch per sweep - a group walking its block row in column order, as the
// reference's row kernel does - was measured and dropped: 0.301 ms against 0.118 + 0.173 ms per
// sweep on C2, 1.87 against 1.74 ms on C3, with the same residual history; the dependent chain
// inside a row costs more than the saved re-read of the fresh L blocks.)

template <int BS>
__device__ __forceinline__ bool row_differs(const double *blk, const int r, const double (&v)[BS])
{
	double cur[BS];
	BlkIO<BS>::load_row_ordered(blk, r, cur);        // must precede the overwrite that follows
	bool ch = false;
#pragma unroll
	for(int c = 0; c < BS; c++) ch |= (__double_as_longlong(cur[c]) != __double_as_longlong(v[c]));   // bitwise: NaN == NaN
	return ch;
}

template <int BS, bool SCALE>
__global__ void __launch_bounds__(256, 3)
block_ilu0_lower_kernel(const long long nlower, const int4 *__restrict__ lmeta,
                        const int *__restrict__ browind, const double *__restrict__ avals,
                        const int2 *__restrict__ pairs, const double *__restrict__ scale,
                        const double *__restrict__ dinv, double *lval, const double *ut,
                        int *__restrict__ changed)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const long long stride = (((long long)gridDim.x*blockDim.x) >> 5)*GPW;
	const bool lanevalid = g < GPW;

	// Three-stage software pipeline over this group's items t, t+stride, ...:
	//   stage M: {entry, col, product range} of item i+2  (16 B, from the packed list)
	//   stage D: data of item i+1 - row r of A_ij, row r of U_jj^-1
	//   stage C: products (rare), L = S * U_jj^-1, single final store of item i
	// so the loads of the next item are in flight while the current one is computed and stored.
	// Every lane runs the same instruction stream (the group products are warp collectives);
	// lanes without an item work on zeros and store nothing.
	long long t = warp*GPW + g;
	const int4 none = make_int4(-1, -1, 0, 0);
	int4 meta1 = none, meta2 = none;
	if(lanevalid && t < nlower) meta1 = __ldg(lmeta + t);
	if(lanevalid && t + stride < nlower) meta2 = __ldg(lmeta + t + stride);

	double arow[BS], drow[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) { arow[c] = 0; drow[c] = 0; }
	if(meta1.x >= 0) {
		BlkIO<BS>::template load_row<false>(avals + (size_t)meta1.x*BS2, r, arow);
		BlkIO<BS>::template load_row<false>(dinv + (size_t)meta1.y*BS2, r, drow);
	}
	int4 meta = meta1;

	const long long niter = (nlower + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		// stage M for item it+2
		int4 meta3 = none;
		if(lanevalid && t + 2*stride < nlower) meta3 = __ldg(lmeta + t + 2*stride);
		// stage D for item it+1
		double arow_n[BS], drow_n[BS];
#pragma unroll
		for(int c = 0; c < BS; c++) { arow_n[c] = 0; drow_n[c] = 0; }
		if(meta2.x >= 0) {
			BlkIO<BS>::template load_row<false>(avals + (size_t)meta2.x*BS2, r, arow_n);
			BlkIO<BS>::template load_row<false>(dinv + (size_t)meta2.y*BS2, r, drow_n);
		}

		// stage C for item it
		const bool active = meta.x >= 0;
		const int entry = active ? meta.x : 0, col = active ? meta.y : 0;
		const int ps = meta.z, pe = meta.w;
		double sum[BS];
#pragma unroll
		for(int c = 0; c < BS; c++) sum[c] = arow[c];
		if(SCALE && active) {
			const int row = __ldg(browind + entry);
			const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
			for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)col*BS + c);
		}
		const int nk = __reduce_max_sync(0xffffffffu, pe - ps);
		for(int k = 0; k < nk; k++) {
			const bool has = ps + k < pe;
			int2 pr = make_int2(0, 0);
			if(has) pr = __ldg(pairs + ps + k);
			double lr[BS], ur[BS];
#pragma unroll
			for(int m = 0; m < BS; m++) { lr[m] = 0; ur[m] = 0; }
			if(has) {
				BlkIO<BS>::template load_row<true>(lval + (size_t)pr.x*BS2, r, lr);
				BlkIO<BS>::template load_row<true>(ut + (size_t)pr.y*BS2, r, ur);
			}
			group_mul_sub<BS>(sum, lr, ur, g*BS);
		}
		double out[BS];
		group_mul<BS>(out, sum, drow, g*BS);                                 // L = S * U_jj^-1
		if(active) {
			double *op = lval + (size_t)t*BS2;
			if(changed && row_differs<BS>(op, r, out)) *changed = 1;
			BlkIO<BS>::store_row(op, r, out);                            // single final store
		}

		// advance the pipeline
		meta = meta2; meta2 = meta3;
#pragma unroll
		for(int c = 0; c < BS; c++) { arow[c] = arow_n[c]; drow[c] = drow_n[c]; }
		t += stride;
	}
}

/// One lower entry: L_ij = (A_ij - sum_k L_ik U_kj) U_jj^-1, single final store to lval[t].
/// WARP-COLLECTIVE (lanes without an item pass meta.x < 0, work on zeros and store nothing).
/// FRESH: U_jj^-1 may have been written earlier in THIS launch (one-launch exact factorisation):
/// read it at L2 instead of through the non-coherent path.
template <int BS, bool SCALE, bool FRESH>
__device__ __forceinline__ void lower_item(const int4 meta, const long long t, const int g, const int r,
                                           const int *__restrict__ browind,
                                           const double *__restrict__ avals,
                                           const int2 *__restrict__ pairs,
                                           const double *__restrict__ scale, const double *dinv,
                                           double *lval, const double *ut, int *__restrict__ changed)
{
	constexpr int BS2 = BS*BS;
	const bool active = meta.x >= 0;
	const int entry = active ? meta.x : 0, col = active ? meta.y : 0;
	const int ps = active ? meta.z : 0, pe = active ? meta.w : 0;
	double sum[BS], drow[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) { sum[c] = 0; drow[c] = 0; }
	if(active) {
		BlkIO<BS>::template load_row<false>(avals + (size_t)entry*BS2, r, sum);
		BlkIO<BS>::template load_row<FRESH>(dinv + (size_t)col*BS2, r, drow);   // row r of U_jj^-1
	}
	if(SCALE && active) {
		const int row = __ldg(browind + entry);
		const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
		for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)col*BS + c);
	}
	const int nk = __reduce_max_sync(0xffffffffu, pe - ps);
	for(int k = 0; k < nk; k++) {
		const bool has = ps + k < pe;
		int2 pr = make_int2(0, 0);
		if(has) pr = __ldg(pairs + ps + k);
		double lr[BS], ur[BS];
#pragma unroll
		for(int m = 0; m < BS; m++) { lr[m] = 0; ur[m] = 0; }
		if(has) {
			BlkIO<BS>::template load_row<true>(lval + (size_t)pr.x*BS2, r, lr);
			BlkIO<BS>::template load_row<true>(ut + (size_t)pr.y*BS2, r, ur);
		}
		group_mul_sub<BS>(sum, lr, ur, g*BS);
	}
	double out[BS];
	group_mul<BS>(out, sum, drow, g*BS);                                 // L = S * U_jj^-1
	if(active) {
		double *op = lval + (size_t)t*BS2;
		if(changed && row_differs<BS>(op, r, out)) *changed = 1;
		BlkIO<BS>::store_row(op, r, out);                            // single final store
	}
}

/// One upper entry: U_ij = A_ij - sum_k L_ik U_kj, written to ut[dest] or, for a diagonal entry
/// (dest = ~row), to udiag[row] together with the refreshed inverse dinv[row].  WARP-COLLECTIVE.
template <int BS, bool SCALE>
__device__ __forceinline__ void upper_item(const int4 meta, const int g, const int r,
                                           const int *__restrict__ browind,
                                           const int *__restrict__ bcolind,
                                           const double *__restrict__ avals,
                                           const int2 *__restrict__ pairs,
                                           const double *__restrict__ scale, double *dinv,
                                           const double *lval, double *ut, double *udiag,
                                           int *__restrict__ changed)
{
	constexpr int BS2 = BS*BS;
	const bool active = meta.x >= 0;
	const bool isdiag = active && meta.w < 0;
	const int entry = active ? meta.x : 0;
	double sum[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) sum[c] = 0;
	if(active) {
		BlkIO<BS>::template load_row<false>(avals + (size_t)entry*BS2, r, sum);
		if(SCALE) {
			const int row = __ldg(browind + entry), col = __ldg(bcolind + entry);
			const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
			for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)col*BS + c);
		}
	}
	// products: warp-uniform trip count, partner blocks exchanged within the group.
	// (Prefetching the next item's first two product pairs one iteration ahead, so that the
	// block loads skip the meta -> pair -> block chain, was measured: 0.173 -> 0.1745 ms on C2
	// and slower for bs = 5.)
	const int ps = active ? meta.y : 0, pe = active ? meta.z : 0;
	const int nk = __reduce_max_sync(0xffffffffu, pe - ps);
	for(int k = 0; k < nk; k++) {
		const bool has = ps + k < pe;
		int2 pr = make_int2(0, 0);
		if(has) pr = __ldg(pairs + ps + k);
		double lr[BS], ur[BS];
#pragma unroll
		for(int m = 0; m < BS; m++) { lr[m] = 0; ur[m] = 0; }
		if(has) {
			BlkIO<BS>::template load_row<true>(lval + (size_t)pr.x*BS2, r, lr);
			BlkIO<BS>::template load_row<true>(ut + (size_t)pr.y*BS2, r, ur);
		}
		group_mul_sub<BS>(sum, lr, ur, g*BS);
	}
	if(active) {
		double *op = isdiag ? udiag + (size_t)(~meta.w)*BS2 : ut + (size_t)meta.w*BS2;
		if(changed && row_differs<BS>(op, r, sum)) *changed = 1;
		BlkIO<BS>::store_row(op, r, sum);
	}
	// a diagonal entry refreshes the compact inverse by the cooperative Gauss-Jordan of
	// blockops.cuh::group_inverse (row m of the new block lives in lane m)
	if(__any_sync(0xffffffffu, isdiag)) {
		double x[BS];
		int prow;
		group_inverse<BS>(sum, x, g*BS, r, prow);
		if(isdiag) BlkIO<BS>::store_row(dinv + (size_t)(~meta.w)*BS2, prow, x);
	}
}

// The lower launch without the register data stage of the pipeline: for bs = 5 the second set of
// row registers costs more in occupancy/spills than the overlap gains.  Three resident CTAs: at four
// (64 registers) the bs = 5 body spills 70 bytes per thread and the spill traffic goes through the
// very L1 data pipe that bounds the launch - C3 lower launch 0.79 ms at four CTAs, 0.64 ms at three.
// (Requesting the next item's blocks into L2 one iteration ahead with prefetch.global.L2 was
// measured as well: 4-11 % SLOWER; the launch is bound by L1 wavefronts, not by DRAM latency.)
#ifndef B200_L5_MINB
#define B200_L5_MINB 3
#endif
template <int BS, bool SCALE>
__global__ void __launch_bounds__(256, B200_L5_MINB)
block_ilu0_lower_simple_kernel(const long long nlower, const int4 *__restrict__ lmeta,
                        const int *__restrict__ browind, const double *__restrict__ avals,
                        const int2 *__restrict__ pairs, const double *__restrict__ scale,
                        const double *__restrict__ dinv, double *lval, const double *ut,
                        int *__restrict__ changed)
{
	constexpr int GPW = 32/BS;
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const long long stride = (((long long)gridDim.x*blockDim.x) >> 5)*GPW;
	const bool lanevalid = g < GPW;

	long long t = warp*GPW + g;
	const int4 none = make_int4(-1, -1, 0, 0);
	int4 meta = none, metan = none;
	if(lanevalid && t < nlower) meta = __ldg(lmeta + t);
	if(lanevalid && t + stride < nlower) metan = __ldg(lmeta + t + stride);
	const long long niter = (nlower + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		int4 metann = none;
		if(lanevalid && t + 2*stride < nlower) metann = __ldg(lmeta + t + 2*stride);   // indices two items ahead
		lower_item<BS,SCALE,false>(meta, t, g, r, browind, avals, pairs, scale, dinv, lval, ut, changed);
		meta = metan; metan = metann;
		t += stride;
	}
}

template <int BS, bool SCALE>
__global__ void __launch_bounds__(256, 4)
block_ilu0_upper_kernel(const long long nupper, const int4 *__restrict__ umeta,
                        const int *__restrict__ browind, const int *__restrict__ bcolind,
                        const double *__restrict__ avals, const int2 *__restrict__ pairs,
                        const double *__restrict__ scale, double *__restrict__ dinv,
                        const double *lval, double *ut, double *udiag, int *__restrict__ changed)
{
	constexpr int GPW = 32/BS;
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const long long stride = (((long long)gridDim.x*blockDim.x) >> 5)*GPW;
	const bool lanevalid = g < GPW;

	long long t = warp*GPW + g;
	const int4 none = make_int4(-1, 0, 0, 0);
	int4 meta = none, metan = none;                       // {entry, pos begin, pos end, dest}
	if(lanevalid && t < nupper) meta = __ldg(umeta + t);
	if(lanevalid && t + stride < nupper) metan = __ldg(umeta + t + stride);
	const long long niter = (nupper + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		int4 metann = none;
		if(lanevalid && t + 2*stride < nupper) metann = __ldg(umeta + t + 2*stride);
		upper_item<BS,SCALE>(meta, g, r, browind, bcolind, avals, pairs, scale, dinv, lval, ut, udiag, changed);
		meta = metan; metan = metann;
		t += stride;
	}
}

// ------------------------------------------------------------------ bs = 5 upper launch, staged
//
// The bs = 5 launches are bound by the L1 data pipe: a 5-lane group reads a 200-byte block as five
// 40-byte segments (5.6 wavefronts per load instruction) and the shuffles of the block products go
// through the same pipe.  Where the pattern allows it (IluPattern::diag_runs_ok: the diagonals are
// the only upper entries with products, at most three each, the L partners of row i a run
// L[l .. l+nk) - every face-neighbour block stencil) the global half is taken off that pipe: every
// warp stages the A block, the L run and the U partners of its six rows in shared memory with TMA
// bulk copies (cp.async.bulk, 208 / <= 624 bytes, completing on the warp's own mbarrier; measured
// in tools/ubench/gather5.cu: 208-byte copies at random positions arrive at 4.4 TB/s, runs of three
// blocks at 5.9, 40-byte segment loads at 2.2) and the groups read their rows from shared memory.
// One stage per warp and six resident CTAs of four warps: a warp's copies are covered by the other
// warps' arithmetic (with two stages per warp the shared memory halves the resident warps and the
// dependent chains of products and inverse are no longer covered: 0.85 against 0.69 ms on C3).
// No CTA-wide synchronisation; same arithmetic as upper_item.
// URUN: the U partners are an index-aligned run of the column-ordered copy (one copy instead of nk).

template <bool SCALE, bool URUN>
__global__ void __launch_bounds__(128)
block5_upper_staged_kernel(const int nrows, const int4 *__restrict__ dmeta,
                           const int4 *__restrict__ dmeta_u,
                           const double *__restrict__ avals, const double *__restrict__ scale,
                           double *__restrict__ dinv, const double *lval, const double *ut,
                           double *udiag)
{
	constexpr int BS = 5, GPW = 6, BS2 = 25;
	// per warp: six group areas {A block, U partners} and one L area for the runs of all six rows
	constexpr int A_BYTES = 208, RUN_BYTES = 624, GROUP_BYTES = A_BYTES + RUN_BYTES;         // 832
	constexpr int L_AREA = GPW*GROUP_BYTES;                                                 // 4992
	constexpr int STAGE_BYTES = L_AREA + GPW*RUN_BYTES;                                     // 8736
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bars[4];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int g = lane / BS, r = lane - g*BS;
	unsigned char *const wsm = smem_raw + (size_t)w*STAGE_BYTES;
	if(lane == 0) mbar_init(&bars[w], 1);
	__syncwarp();
	const long long warp = (long long)blockIdx.x*(blockDim.x >> 5) + w;
	const long long stride = (long long)gridDim.x*(blockDim.x >> 5)*GPW;
	const bool lanevalid = g < GPW;
	long long t = warp*GPW + g;
	const int4 none = make_int4(-1, 0, 0, 0);
	auto getmeta = [&](const long long tt) { return (lanevalid && tt < nrows) ? __ldg(dmeta + tt) : none; };
	auto getu = [&](const long long tt) { return (lanevalid && tt < nrows) ? __ldg(dmeta_u + tt) : none; };

	int4 m0 = getmeta(t), m1 = getmeta(t + stride);           // {A entry, first L index, products, -}
	int4 u0 = getu(t), u1 = getu(t + stride);                 // U indices of the products
	const long long niter = (nrows + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		const int4 m2 = getmeta(t + 2*stride), u2 = getu(t + 2*stride);
		const bool active = m0.x >= 0;
		const int row = active ? (int)t : 0, nk = active ? m0.z : 0;
		// lane 0 of every group copies the A block, lanes 2.. the U partners; the L runs of the
		// warp's six consecutive rows are adjacent in lval: ONE copy for all of them (lane 1 of the
		// warp) unless the span is longer than the staged area (rows whose run is not their whole
		// lower part), in which case lane 1 of every group copies its own run
		const int uidx = (r == 2) ? u0.x : (r == 3) ? u0.y : u0.z;
		const int lfirst = __reduce_min_sync(0xffffffffu, nk > 0 ? m0.y : 0x7fffffff);
		const int llast = __reduce_max_sync(0xffffffffu, nk > 0 ? m0.y + nk : 0);
		const bool lspan = (llast - lfirst) <= GPW*3;
		unsigned loff;                                        // byte offset of this group's L run in the L area
		{
			const size_t s0 = (size_t)(lval + (size_t)(lspan ? lfirst : m0.y)*BS2);
			loff = lspan ? (unsigned)((size_t)(nk > 0 ? m0.y - lfirst : 0)*BS2*8 + (s0 & 15))
			             : (unsigned)(g*RUN_BYTES + (s0 & 15));
		}
		{
			unsigned bytes = 0, dstoff = 0;
			size_t a = 0;
			bool to_l_area = false;
			if(lspan && lane == 1 && llast > lfirst) {
				a = (size_t)(lval + (size_t)lfirst*BS2); to_l_area = true;
				bytes = ((unsigned)((llast - lfirst)*BS2*8) + (unsigned)(a & 15) + 15u) & ~15u;
			}
			else if(active) {
				if(r == 0) { a = (size_t)(avals + (size_t)m0.x*BS2); bytes = A_BYTES; }
				else if(!lspan && r == 1 && nk > 0) {
					a = (size_t)(lval + (size_t)m0.y*BS2); to_l_area = true; dstoff = g*RUN_BYTES;
					bytes = ((unsigned)(nk*BS2*8) + (unsigned)(a & 15) + 15u) & ~15u;
				}
				else if(URUN && r == 2 && nk > 0) {
					a = (size_t)(ut + (size_t)u0.x*BS2); dstoff = A_BYTES;
					bytes = ((unsigned)(nk*BS2*8) + (unsigned)(a & 15) + 15u) & ~15u;
				}
				else if(!URUN && r >= 2 && r - 2 < nk) {
					a = (size_t)(ut + (size_t)uidx*BS2); dstoff = A_BYTES + (r - 2)*208;
					bytes = 208;
				}
			}
			const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
			if(lane == 0) mbar_expect_tx(&bars[w], total);
			if(bytes)
				bulk_g2s(to_l_area ? wsm + L_AREA + dstoff : wsm + (size_t)g*GROUP_BYTES + dstoff,
				         (const void*)(a & ~(size_t)15), bytes, &bars[w]);
		}
		mbar_wait(&bars[w], (unsigned)(it & 1));

		const unsigned char *gs = wsm + (size_t)g*GROUP_BYTES;
		double sum[BS];
#pragma unroll
		for(int c = 0; c < BS; c++) sum[c] = 0;
		if(active) {
			const double *sa = reinterpret_cast<const double*>(gs) + ((((size_t)(avals + (size_t)m0.x*BS2)) & 15) >> 3);
#pragma unroll
			for(int c = 0; c < BS; c++) sum[c] = sa[c*BS + r];
			if(SCALE) {
				const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
				for(int c = 0; c < BS; c++) sum[c] *= sr*__ldg(scale + (size_t)row*BS + c);
			}
		}
		const double *sl = reinterpret_cast<const double*>(wsm + L_AREA + loff);
		const int nkmax = __reduce_max_sync(0xffffffffu, nk);
		for(int k = 0; k < nkmax; k++) {
			double lr[BS], ur[BS];
#pragma unroll
			for(int c = 0; c < BS; c++) { lr[c] = 0; ur[c] = 0; }
			if(k < nk) {
				const double *su;
				if(URUN) su = reinterpret_cast<const double*>(gs + A_BYTES)
					+ ((((size_t)(ut + (size_t)u0.x*BS2)) & 15) >> 3) + k*BS2;
				else {
					const int uk = (k == 0) ? u0.x : (k == 1) ? u0.y : u0.z;
					su = reinterpret_cast<const double*>(gs + A_BYTES + k*208)
						+ ((((size_t)(ut + (size_t)uk*BS2)) & 15) >> 3);
				}
#pragma unroll
				for(int c = 0; c < BS; c++) { lr[c] = sl[k*BS2 + c*BS + r]; ur[c] = su[c*BS + r]; }
			}
			group_mul_sub<BS>(sum, lr, ur, g*BS);
		}
		if(active) BlkIO<BS>::store_row(udiag + (size_t)row*BS2, r, sum);
		{
			double x[BS];
			int prow;
			group_inverse<BS>(sum, x, g*BS, r, prow);
			if(active) BlkIO<BS>::store_row(dinv + (size_t)row*BS2, prow, x);
		}
		__syncwarp();                       // everybody has read the stage before it is refilled
		m0 = m1; m1 = m2; u0 = u1; u1 = u2; t += stride;
	}
}

/// The lower launch in the same form (IluPattern::lower_rows_ok: no lower entry has products, at
/// most three lower entries per row): a group per ROW; per row one copy of the run of A blocks of
/// its lower part and one 208-byte copy of U_jj^-1 per entry; L_ij = A_ij U_jj^-1 is stored to the
/// run lval[l .. l+nl).
template <bool SCALE>
__global__ void __launch_bounds__(128)
block5_lower_staged_kernel(const int nrows, const int4 *__restrict__ rmeta,
                           const int4 *__restrict__ rcols, const double *__restrict__ avals,
                           const double *__restrict__ scale, const double *__restrict__ dinv,
                           double *lval)
{
	constexpr int BS = 5, GPW = 6, BS2 = 25;
	// per warp: six runs of A blocks, then three areas (one per entry position e) of six U_jj^-1 blocks
	constexpr int RUN_BYTES = 624, D_AREA = GPW*208;                                        // 624, 1248
	constexpr int STAGE_BYTES = GPW*RUN_BYTES + 3*D_AREA;                                   // 7488
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bars[4];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int g = lane / BS, r = lane - g*BS;
	unsigned char *const wsm = smem_raw + (size_t)w*STAGE_BYTES;
	if(lane == 0) mbar_init(&bars[w], 1);
	__syncwarp();
	const long long warp = (long long)blockIdx.x*(blockDim.x >> 5) + w;
	const long long stride = (long long)gridDim.x*(blockDim.x >> 5)*GPW;
	const bool lanevalid = g < GPW;
	long long t = warp*GPW + g;
	const int4 none = make_int4(0, 0, 0, 0);
	auto getmeta = [&](const long long tt) { return (lanevalid && tt < nrows) ? __ldg(rmeta + tt) : none; };
	auto getcols = [&](const long long tt) { return (lanevalid && tt < nrows) ? __ldg(rcols + tt) : none; };

	int4 m0 = getmeta(t), m1 = getmeta(t + stride);           // {first A entry, first L index, entries, -}
	int4 c0 = getcols(t), c1 = getcols(t + stride);
	const long long niter = (nrows + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		const int4 m2 = getmeta(t + 2*stride), c2 = getcols(t + 2*stride);
		const int nl = m0.z;
		const int row = (int)t;
		// lane 0 of every group copies the run of A blocks, lanes 1..3 the inverses U_jj^-1 of its
		// entries - unless the e-th columns of the warp's six rows are consecutive (structured
		// grids: rows i..i+5 read columns j..j+5), then lane e+1 of the warp copies all six at once
		const int col = (r == 1) ? c0.x : (r == 2) ? c0.y : c0.z;
		bool merged[3];
		int colbase[3];
#pragma unroll
		for(int e = 0; e < 3; e++) {
			const int ce = (e == 0) ? c0.x : (e == 1) ? c0.y : c0.z;
			colbase[e] = __shfl_sync(0xffffffffu, ce, 0);
			merged[e] = __all_sync(0xffffffffu, !lanevalid || (t < nrows && nl > e && ce == colbase[e] + g));
		}
		{
			unsigned bytes = 0, dstoff = 0;
			size_t a = 0;
			// (selects, not merged[lane-1]: a lane-indexed array would live in local memory)
			const bool my_merged = (r == 1) ? merged[0] : (r == 2) ? merged[1] : merged[2];
			const int my_base = (r == 1) ? colbase[0] : (r == 2) ? colbase[1] : colbase[2];
			if(lane >= 1 && lane <= 3 && my_merged) {
				const int e = lane - 1;
				a = (size_t)(dinv + (size_t)my_base*BS2); dstoff = GPW*RUN_BYTES + e*D_AREA;
				bytes = ((unsigned)(GPW*BS2*8) + (unsigned)(a & 15) + 15u) & ~15u;
			}
			else if(nl > 0) {
				if(r == 0) {
					a = (size_t)(avals + (size_t)m0.x*BS2); dstoff = g*RUN_BYTES;
					bytes = ((unsigned)(nl*BS2*8) + (unsigned)(a & 15) + 15u) & ~15u;
				}
				else if(r >= 1 && r <= 3 && r - 1 < nl && !my_merged) {
					a = (size_t)(dinv + (size_t)col*BS2); dstoff = GPW*RUN_BYTES + (r - 1)*D_AREA + g*208;
					bytes = 208;
				}
			}
			const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
			if(lane == 0) mbar_expect_tx(&bars[w], total);
			if(bytes) bulk_g2s(wsm + dstoff, (const void*)(a & ~(size_t)15), bytes, &bars[w]);
		}
		mbar_wait(&bars[w], (unsigned)(it & 1));

		const double *sa = reinterpret_cast<const double*>(wsm + (size_t)g*RUN_BYTES)
			+ ((((size_t)(avals + (size_t)m0.x*BS2)) & 15) >> 3);
		const int nlmax = __reduce_max_sync(0xffffffffu, nl);
#pragma unroll
		for(int e = 0; e < 3; e++) {
			if(e >= nlmax) break;
			const bool has = e < nl;
			const int ce = (e == 0) ? c0.x : (e == 1) ? c0.y : c0.z;
			double s[BS], drow[BS];
#pragma unroll
			for(int c = 0; c < BS; c++) { s[c] = 0; drow[c] = 0; }
			if(has) {
				const unsigned char *da = wsm + GPW*RUN_BYTES + e*D_AREA;
				const double *sd = merged[e]
					? reinterpret_cast<const double*>(da + (((size_t)(dinv + (size_t)colbase[e]*BS2)) & 15)) + g*BS2
					: reinterpret_cast<const double*>(da + g*208 + (((size_t)(dinv + (size_t)ce*BS2)) & 15));
#pragma unroll
				for(int c = 0; c < BS; c++) { s[c] = sa[e*BS2 + c*BS + r]; drow[c] = sd[c*BS + r]; }
				if(SCALE) {
					const double sr = __ldg(scale + (size_t)row*BS + r);
#pragma unroll
					for(int c = 0; c < BS; c++) s[c] *= sr*__ldg(scale + (size_t)ce*BS + c);
				}
			}
			double out[BS];
			group_mul<BS>(out, s, drow, g*BS);                               // L = A_ij * U_jj^-1
			if(has) BlkIO<BS>::store_row(lval + (size_t)(m0.y + e)*BS2, r, out);   // single final store
		}
		__syncwarp();                       // everybody has read the stage before it is refilled
		m0 = m1; m1 = m2; c0 = c1; c1 = c2; t += stride;
	}
}

// ------------------------------------------------------------------ exact block ILU(0), one launch

__device__ __forceinline__ int ld_poll_flag(const int *p)
{
	int v;
	asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

/// Exact ("sequential") point-block ILU(0) in ONE launch instead of iterating sweeps to the fixed
/// point (up to one sweep per dependency level: 2047 on C2).  A group of BS lanes per block row;
/// rows are taken from a level-sorted list in which every level is padded to whole warps (slot
/// value -1), so the rows of a warp are mutually independent; CTAs take their position from a
/// ticket counter, so every row a warp can wait for belongs to a CTA that has already started.  A
/// warp waits until all rows named by its rows' lower parts have raised their flags, then every
/// group walks its own row in storage order - lower entries, diagonal, upper entries: the
/// reference's sequential pass (tests/solverops/async_ilu_convergence.cpp:462-490,
/// src/solverfactory.cpp:93-107) - with the very per-entry arithmetic of the sweep launches
/// (lower_item / upper_item), hence the same values as their fixed point.  The spin is bounded
/// (error flag), as in the one-launch substitutions of apply.cu.
template <int BS, bool SCALE>
__global__ void __launch_bounds__(256)
block_ilu0_exact_kernel(const int nslots, const int *__restrict__ slots, const int *__restrict__ lptr,
                        const int *__restrict__ uptr, const int *__restrict__ lcol,
                        const int4 *__restrict__ lmeta, const int4 *__restrict__ uall,
                        const int *__restrict__ browind, const int *__restrict__ bcolind,
                        const double *__restrict__ avals, const int2 *__restrict__ pairs,
                        const double *__restrict__ scale, double *dinv, double *lval, double *ut,
                        double *udiag, int *rowdone, int *__restrict__ ticket, int *__restrict__ err)
{
	constexpr int GPW = 32/BS;
	__shared__ int s_cta;
	if(threadIdx.x == 0) s_cta = atomicAdd(ticket, 1);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const long long warp = (long long)s_cta*(blockDim.x >> 5) + (threadIdx.x >> 5);
	const long long t = warp*GPW + g;
	int row = -1;
	if(g < GPW && t < nslots) row = __ldg(slots + t);
	int ls = 0, le = 0, us = 0, ue = 0;
	if(row >= 0) {
		ls = __ldg(lptr + row); le = __ldg(lptr + row + 1);
		us = __ldg(uptr + row) + row; ue = __ldg(uptr + row + 1) + row + 1;
	}
	// wait for the rows this row reads (every lane of a group polls the same flags: broadcast)
	int kdep = ls, spins = 0;
	while(true) {
		while(kdep < le && ld_poll_flag(rowdone + __ldg(lcol + kdep)) != 0) kdep++;
		if(__all_sync(0xffffffffu, kdep == le)) break;
		__nanosleep(64);
		if(++spins > (1 << 16) && ((spins & 1023) == 0)) {
			if(spins > (1 << 22) || *((volatile int*)err)) { *err = 1; kdep = le; }
		}
	}
	__threadfence();
	const int nl = __reduce_max_sync(0xffffffffu, le - ls);
	const int4 lnone = make_int4(-1, -1, 0, 0), unone = make_int4(-1, 0, 0, 0);
	for(int e = 0; e < nl; e++) {
		const int4 meta = (ls + e < le) ? __ldg(lmeta + ls + e) : lnone;
		lower_item<BS,SCALE,true>(meta, ls + e, g, r, browind, avals, pairs, scale, dinv, lval, ut, nullptr);
	}
	const int nu = __reduce_max_sync(0xffffffffu, ue - us);
	for(int e = 0; e < nu; e++) {
		const int4 meta = (us + e < ue) ? __ldg(uall + us + e) : unone;
		upper_item<BS,SCALE>(meta, g, r, browind, bcolind, avals, pairs, scale, dinv, lval, ut, udiag, nullptr);
	}
	__threadfence();
	__syncwarp();
	if(row >= 0 && r == 0) *((volatile int*)(rowdone + row)) = 1;
}

/// slots[base[l] + (i - level_ptr[l])] = level_rows[i] for row position i of level l
__global__ void __launch_bounds__(256)
exact_slots_kernel(const int n, const int nlevels, const int *__restrict__ level_ptr,
                   const int *__restrict__ slot_base, const int *__restrict__ level_rows,
                   int *__restrict__ slots)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	int lo = 0, hi = nlevels - 1;                 // level of position i: last l with level_ptr[l] <= i
	while(lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if(level_ptr[mid] <= i) lo = mid; else hi = mid - 1;
	}
	slots[slot_base[lo] + (i - level_ptr[lo])] = level_rows[i];
}


/// dst block at positions[i] <- src block i
template <int BS>
__global__ void scatter_blocks_kernel(const int nbrows, const double *__restrict__ src,
                                      const int *__restrict__ positions, double *__restrict__ dst)
{
	constexpr int BS2 = BS*BS;
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= (long long)nbrows*BS2) return;
	const long long i = e / BS2;
	const int w = (int)(e - i*BS2);
	dst[(size_t)positions[i]*BS2 + w] = src[e];
}

void launch_scatter_blocks(const Mat& A, const double *src, const int *positions, double *dst,
                           cudaStream_t st)
{
	const long long n = (long long)A.nbrows*A.bs*A.bs;
	if(n == 0) return;
	ProfScope ps(KC_DIAG_INVERT, st);
	const int grid = div_up(n, 256);
	switch(A.bs) {
	case 4: scatter_blocks_kernel<4><<<grid,256,0,st>>>(A.nbrows, src, positions, dst); break;
	case 5: scatter_blocks_kernel<5><<<grid,256,0,st>>>(A.nbrows, src, positions, dst); break;
	default: throw Error("scatter blocks: unsupported block size");
	}
	B200_LAUNCHED();
}

template <int BS>
__global__ void invert_blocks_kernel(const int nbrows, const double *src, const int *__restrict__ positions,
                                     double *dst, const bool dst_compact);

/// Persistent grid size: resident CTAs of `kernel` over all SMs (cached per kernel)
template <typename K>
static int persistent_grid(K kernel, long long nitems_per_cta_min, long long nitems)
{
	static int cached = 0;
	if(!cached) {
		int dev = 0, sms = 148, per = 4;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, 256, 0) != cudaSuccess || per < 1)
			per = 4;
		cached = sms*per;
	}
	const long long need = (nitems + nitems_per_cta_min - 1)/nitems_per_cta_min;
	return (int)std::max<long long>(1, std::min<long long>(cached, need));
}

template <int BS>
static void launch_block_sweep(const Mat& A, const IluPattern& pl, const double *scale, ScalarFactor& F,
                               double *dinv, int *changed, bool all_upper, cudaStream_t st)
{
	const long long nup = all_upper ? pl.nupper : pl.nuwork;
	const int4 *uplist = all_upper ? pl.suall.p : pl.suwork.p;
	constexpr int GPW = 32/BS;
	const long long per_cta = 8*GPW;
	static const bool force_simple = getenv("B200_LOWER_SIMPLE") != nullptr;   // A/B switch (development)
	const bool pipelined = (BS <= 4) && !force_simple;
	static const bool no_staged_l = getenv("B200_NO_STAGED") != nullptr || getenv("B200_NO_STAGED_LOWER") != nullptr;
	if(BS == 5 && pl.nlower > 0 && !changed && pl.lower_rows_ok && !no_staged_l) {
		ProfScope ps(KC_FACTOR_LOWER, st);
		constexpr int SMEM = 4*6*(624 + 3*208);
		static int grid = 0;
		if(!grid) {
			int dev = 0, sms = 148, per = 1;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			B200_CUDA(cudaFuncSetAttribute(block5_lower_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			B200_CUDA(cudaFuncSetAttribute(block5_lower_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, block5_lower_staged_kernel<false>, 128, SMEM) != cudaSuccess || per < 1) per = 1;
			grid = sms*per;
		}
		const int g = (int)std::max<long long>(1, std::min<long long>(grid, (A.nbrows + 23)/24));
		if(scale) block5_lower_staged_kernel<true><<<g, 128, SMEM, st>>>(A.nbrows, pl.lrow_meta, pl.lrow_cols, A.vals, scale, dinv, F.lval.p);
		else block5_lower_staged_kernel<false><<<g, 128, SMEM, st>>>(A.nbrows, pl.lrow_meta, pl.lrow_cols, A.vals, scale, dinv, F.lval.p);
		B200_LAUNCHED();
	}
	else if(pl.nlower > 0) {
		ProfScope ps(KC_FACTOR_LOWER, st);
#define B200_LOWER(K)                                                                             \
		{ auto k = K; k<<<persistent_grid(k, per_cta, pl.nlower), 256, 0, st>>>(pl.nlower, pl.slmeta, \
			A.browind, A.vals, pl.spairs, scale, dinv, F.lval.p, F.ut.p ? F.ut.p : F.uval.p, changed); }
		if(scale) {
			if(pipelined) B200_LOWER((block_ilu0_lower_kernel<BS,true>))
			else B200_LOWER((block_ilu0_lower_simple_kernel<BS,true>))
		} else {
			if(pipelined) B200_LOWER((block_ilu0_lower_kernel<BS,false>))
			else B200_LOWER((block_ilu0_lower_simple_kernel<BS,false>))
		}
#undef B200_LOWER
		B200_LAUNCHED();
	}
	static const bool no_staged = getenv("B200_NO_STAGED") != nullptr;            // A/B switch (development)
	if(BS == 5 && nup > 0 && !all_upper && !changed && pl.diag_runs_ok && !no_staged) {
		// the diagonals are the only upper entries that change, their L partners are runs: rows
		// staged in shared memory by TMA (block5_upper_staged_kernel)
		ProfScope ps(KC_FACTOR_UPPER, st);
		constexpr int SMEM = 4*6*(208 + 2*624);
		const bool urun = F.ut.p && pl.diag_u_runs;
		static int grid = 0;
		if(!grid) {
			int dev = 0, sms = 148, per = 1;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			B200_CUDA(cudaFuncSetAttribute(block5_upper_staged_kernel<true,true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			B200_CUDA(cudaFuncSetAttribute(block5_upper_staged_kernel<false,true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			B200_CUDA(cudaFuncSetAttribute(block5_upper_staged_kernel<true,false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			B200_CUDA(cudaFuncSetAttribute(block5_upper_staged_kernel<false,false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, block5_upper_staged_kernel<false,false>, 128, SMEM) != cudaSuccess || per < 1) per = 1;
			grid = sms*per;
		}
		const int g = (int)std::max<long long>(1, std::min<long long>(grid, (A.nbrows + 23)/24));
		const double *usrc = F.ut.p ? F.ut.p : F.uval.p;
#define B200_STAGED(SC, UR) block5_upper_staged_kernel<SC,UR><<<g, 128, SMEM, st>>>(A.nbrows, pl.dmeta, pl.dmeta_u, \
			A.vals, scale, dinv, F.lval.p, usrc, F.udiag.p)
		if(urun) { if(scale) B200_STAGED(true, true); else B200_STAGED(false, true); }
		else { if(scale) B200_STAGED(true, false); else B200_STAGED(false, false); }
#undef B200_STAGED
		B200_LAUNCHED();
	}
	else if(nup > 0) {
		ProfScope ps(KC_FACTOR_UPPER, st);
		// the diagonal entries refresh U_ii^-1 inside this launch by the cooperative Gauss-Jordan
		// of blockops.cuh::group_inverse (one row per lane: 64 registers, 4 CTAs/SM).  The
		// redundant per-lane elimination it replaced cost bs = 4 a resident CTA (80 registers,
		// 0.173 -> 0.165 ms on C2) and forced bs = 5 into a separate pass over the diagonal
		// blocks (0.441 -> 0.378 ms on 96^3).
#define B200_UPPER(K)                                                                             \
		{ auto k = K; k<<<persistent_grid(k, per_cta, nup), 256, 0, st>>>(nup, uplist, A.browind,     \
			A.bcolind, A.vals, pl.spairs, scale, dinv, F.lval.p, F.ut.p ? F.ut.p : F.uval.p, F.udiag.p, changed); }
		if(scale) B200_UPPER((block_ilu0_upper_kernel<BS,true>))
		else B200_UPPER((block_ilu0_upper_kernel<BS,false>))
#undef B200_UPPER
		B200_LAUNCHED();
	}
}

static size_t block_count(const Mat& A, long long nblocks) { return (size_t)std::max<long long>(nblocks, 1)*A.bs*A.bs; }

void block_factor_alloc(const Mat& A, const IluPattern& pl, ScalarFactor& F)
{
	F.lval.alloc(block_count(A, pl.nlower));
	F.uval.alloc(block_count(A, pl.nstrict));
	if(pl.utpos.p) F.ut.alloc(block_count(A, pl.nstrict));      // optional column-order copy
	F.udiag.alloc(block_count(A, A.nbrows));
}

template <int BS>
static void launch_block_init(const Mat& A, const IluPattern& pl, const double *scale, int fact_init,
                              ScalarFactor& F, cudaStream_t st, long long e0 = 0, long long e1 = -1)
{
	constexpr int GPW = 32/BS;
	if(e1 < 0) e1 = A.nnzb;
	if(e1 <= e0) return;
	const long long nwarps = (e1 - e0 + GPW - 1)/GPW;
	const int grid = div_up(nwarps*32, 256);
#define B200_INIT(SC, MODE) block_split_init_kernel<BS,SC,MODE><<<grid,256,0,st>>>(e0, e1, A.browptr, \
		A.bcolind, A.browind, A.diagind, pl.lptr, pl.uptr, pl.utpos, A.vals, scale, F.lval.p,         \
		F.udiag.p, F.uval.p, F.ut.p)
	if(fact_init == B200_INIT_F_SGS) { if(scale) B200_INIT(true, MODE_INIT_SGS); else B200_INIT(false, MODE_INIT_SGS); }
	else { if(scale) B200_INIT(true, MODE_INIT_ORIG); else B200_INIT(false, MODE_INIT_ORIG); }
#undef B200_INIT
	B200_LAUNCHED();
}

void launch_ilu0_init(const Mat& A, const IluPattern& pl, const double *scale, int fact_init,
                      ScalarFactor& F, cudaStream_t st)
{
	if(fact_init == B200_INIT_F_NONE || A.nnzb == 0) return;
	ProfScope ps(KC_FACTOR_INIT, st);
	if(fact_init == B200_INIT_F_ZERO) {
		// the block version really zeroes (async_blockilu_factor.cpp:65-69); the scalar one falls
		// through into INIT_F_ORIGINAL (async_ilu_factor.cpp:48-54) - both replicated
		B200_CUDA(cudaMemsetAsync(F.lval, 0, block_count(A, pl.nlower)*sizeof(double), st));
		B200_CUDA(cudaMemsetAsync(F.uval, 0, block_count(A, pl.nstrict)*sizeof(double), st));
		if(F.ut.p) B200_CUDA(cudaMemsetAsync(F.ut, 0, block_count(A, pl.nstrict)*sizeof(double), st));
		B200_CUDA(cudaMemsetAsync(F.udiag, 0, block_count(A, A.nbrows)*sizeof(double), st));
		return;
	}
	switch(A.bs) {
	case 4: launch_block_init<4>(A, pl, scale, fact_init, F, st); break;
	case 5: launch_block_init<5>(A, pl, scale, fact_init, F, st); break;
	default: throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
	}
}

void launch_ilu0_init_range(const Mat& A, const IluPattern& pl, ScalarFactor& F, long long e0,
                            long long e1, cudaStream_t st)
{
	// INIT_F_ORIGINAL without scaling reads nothing but the entry itself: it can run on the block
	// entries [e0, e1) alone (chunk-wise, behind the upload of the next chunk)
	ProfScope ps(KC_FACTOR_INIT, st);
	switch(A.bs) {
	case 4: launch_block_init<4>(A, pl, nullptr, B200_INIT_F_ORIGINAL, F, st, e0, e1); break;
	case 5: launch_block_init<5>(A, pl, nullptr, B200_INIT_F_ORIGINAL, F, st, e0, e1); break;
	default: throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
	}
}

void launch_ilu0_sweep(const Mat& A, const IluPattern& pl, const double *scale, ScalarFactor& F,
                       double *dinv, int *d_changed, bool all_upper, cudaStream_t st)
{
	if(A.nnzb == 0) return;
	if(A.bs == 4) { launch_block_sweep<4>(A, pl, scale, F, dinv, d_changed, all_upper, st); return; }
	if(A.bs == 5) { launch_block_sweep<5>(A, pl, scale, F, dinv, d_changed, all_upper, st); return; }
	throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
}

int block_exact_slots(const Mat& A, const Levels& lv, DevBuf<int>& slots, cudaStream_t st)
{
	// every level padded to whole warps of groups, so that the rows of a warp never depend on one
	// another (host prefix over the level sizes; the levels are a few thousand at most)
	const int GPW = 32/A.bs;
	std::vector<int> base(lv.nlevels + 1, 0);
	for(int l = 0; l < lv.nlevels; l++) {
		const int nl = lv.level_ptr[l+1] - lv.level_ptr[l];
		base[l+1] = base[l] + (nl + GPW - 1)/GPW*GPW;
	}
	const int nslots = base[lv.nlevels];
	slots.alloc(std::max(nslots, 1));
	B200_CUDA(cudaMemsetAsync(slots, 0xff, (size_t)std::max(nslots, 1)*sizeof(int), st));
	DevBuf<int> d_base;
	d_base.alloc(lv.nlevels + 1);
	B200_CUDA(cudaMemcpyAsync(d_base, base.data(), (lv.nlevels + 1)*sizeof(int), cudaMemcpyHostToDevice, st));
	exact_slots_kernel<<<div_up(A.nbrows, 256), 256, 0, st>>>(A.nbrows, lv.nlevels, lv.d_level_ptr, d_base,
	                                                       lv.level_rows, slots);
	B200_LAUNCHED();
	B200_CUDA(cudaStreamSynchronize(st));             // `base` and d_base go out of scope
	return nslots;
}

void launch_ilu0_exact(const Mat& A, const IluPattern& pl, const int *slots, int nslots,
                       const double *scale, ScalarFactor& F, double *dinv, int *rowdone, int *flags,
                       cudaStream_t st)
{
	if(A.nbrows == 0) return;
	ProfScope ps(KC_FACTOR_LOWER, st);
	B200_CUDA(cudaMemsetAsync(rowdone, 0, (size_t)A.nbrows*sizeof(int), st));
	B200_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));                // ticket
	const int GPW = 32/A.bs;
	const int grid = div_up((long long)nslots/GPW*32, 256);
	double *ut = F.ut.p ? F.ut.p : F.uval.p;
#define B200_EXACT(BSV, SC) block_ilu0_exact_kernel<BSV,SC><<<grid,256,0,st>>>(nslots, slots, pl.lptr, pl.uptr, \
		pl.lcol, pl.slmeta, pl.suall, A.browind, A.bcolind, A.vals, pl.spairs, scale, dinv, F.lval.p, ut,      \
		F.udiag.p, rowdone, flags, flags + 1)
	if(A.bs == 4) { if(scale) B200_EXACT(4, true); else B200_EXACT(4, false); }
	else if(A.bs == 5) { if(scale) B200_EXACT(5, true); else B200_EXACT(5, false); }
	else throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
#undef B200_EXACT
	B200_LAUNCHED();
}

void launch_sync_upper(const Mat& A, const IluPattern& pl, ScalarFactor& F, bool all, cudaStream_t st)
{
	// strict upper entries that the sweeps may have changed: row-order copy <- column-order copy
	if(!F.ut.p) return;                                               // one copy only: nothing to sync
	const long long n = all ? pl.nstrict : pl.nuwork - A.nbrows;      // work list = diagonals + entries with products
	if(n <= 0) return;
	const long long items = all ? pl.nstrict : pl.nuwork;
	const int4 *list = all ? nullptr : pl.suwork.p;
	ProfScope ps(KC_OTHER, st);
	const int grid = div_up(items*A.bs*A.bs, 256);
	switch(A.bs) {
	case 4: ut_to_uval_kernel<4><<<grid,256,0,st>>>(items, list, pl.ut_order, F.ut.p, F.uval.p); break;
	case 5: ut_to_uval_kernel<5><<<grid,256,0,st>>>(items, list, pl.ut_order, F.ut.p, F.uval.p); break;
	default: throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
	}
	B200_LAUNCHED();
}

void block_factor_assemble(const Mat& A, const IluPattern& pl, const ScalarFactor& F,
                           const double *diag_src, double *out, cudaStream_t st)
{
	if(A.nnzb == 0) return;
	const int grid = div_up(A.nnzb*A.bs*A.bs, 256);
	switch(A.bs) {
	case 4: block_assemble_kernel<4><<<grid,256,0,st>>>(A.nnzb, A.browptr, A.browind, A.diagind, pl.lptr,
	                                                    pl.uptr, F.lval.p, diag_src, F.uval.p, out); break;
	case 5: block_assemble_kernel<5><<<grid,256,0,st>>>(A.nnzb, A.browptr, A.browind, A.diagind, pl.lptr,
	                                                    pl.uptr, F.lval.p, diag_src, F.uval.p, out); break;
	default: throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
	}
	B200_LAUNCHED();
}

double ilu0_residual(const Mat& A, const IluPattern& pl, const double *scale, const ScalarFactor& F,
                     double *d_scratch, cudaStream_t st)
{
	if(A.nnzb == 0) return 0.0;
	// diagnostics only: assembled in matrix order (un-inverted diagonal blocks) for the reference's
	// own position lists
	DevBuf<double> tmp;
	tmp.alloc((size_t)A.nnzb*A.bs*A.bs);
	block_factor_assemble(A, pl, F, F.udiag.p, tmp, st);
	B200_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(double), st));
	const int GPW = 32/A.bs;
	const long long nwarps = (A.nnzb + GPW - 1)/GPW;
	const int grid = div_up(nwarps*32, 256);
#define B200_RES(BSV, SC) block_ilu0_residual_kernel<BSV,SC><<<grid,256,0,st>>>(A.nnzb, A.bcolind, A.browind, \
		A.diagind, A.vals, pl.posptr, pl.lowerp, pl.upperp, scale, tmp.p, d_scratch)
	if(A.bs == 4) { if(scale) B200_RES(4, true); else B200_RES(4, false); }
	else if(A.bs == 5) { if(scale) B200_RES(5, true); else B200_RES(5, false); }
	else throw Error("ILU0: unsupported block size " + std::to_string(A.bs));
#undef B200_RES
	B200_LAUNCHED();
	double r = 0;
	B200_CUDA(cudaMemcpyAsync(&r, d_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return r;
}

// ------------------------------------------------------------------ diagonal block inversion

/// dst block i <- inverse of src block at positions[i] (or i).  Works in place.
template <int BS>
__global__ void __launch_bounds__(256)
invert_blocks_kernel(const int nbrows, const double *src, const int *__restrict__ positions,
                     double *dst, const bool dst_compact)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int g = lane / BS, r = lane - g*BS;
	const long long rowl = warp*GPW + g;
	const bool active = (g < GPW) && (rowl < nbrows);
	size_t spos = 0;
	if(active) spos = positions ? (size_t)__ldg(positions + rowl) : (size_t)rowl;
	// every lane reads its own row of the block; the group inverts it cooperatively
	// (blockops.cuh::group_inverse) and each lane ends up with the row of the inverse whose pivot
	// its row supplied
	double rowv[BS], x[BS];
	int prow = 0;
#pragma unroll
	for(int c = 0; c < BS; c++) rowv[c] = (c == r) ? 1.0 : 0.0;     // idle groups invert an identity
	if(active) BlkIO<BS>::template load_row<true>(src + spos*BS2, r, rowv);
	group_inverse<BS>(rowv, x, g*BS, r, prow);
	__syncwarp();                          // all lanes have read the block before anyone overwrites it
	if(active)
		BlkIO<BS>::store_row(dst + (dst_compact ? (size_t)rowl : spos)*BS2, prow, x);
}

__global__ void invert_scalars_kernel(const int n, const double *__restrict__ src,
                                      const int *__restrict__ positions, double *__restrict__ dst)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) dst[i] = 1.0/src[positions ? positions[i] : i];
}

void launch_invert_diag_blocks(const Mat& A, const double *src_vals, const int *positions,
                               double *dst, bool dst_is_compact, cudaStream_t st)
{
	if(A.nbrows == 0) return;
	ProfScope ps(KC_DIAG_INVERT, st);
	if(A.bs == 1) {
		// only used for the Jacobi-type objects (scalar_jacobi_setup, solverops_jacobi.cpp:141-147)
		invert_scalars_kernel<<<div_up(A.nbrows,256),256,0,st>>>(A.nbrows, src_vals, positions, dst);
		B200_LAUNCHED();
		return;
	}
	const int GPW = 32/A.bs;
	const long long nwarps = ((long long)A.nbrows + GPW - 1)/GPW;
	const int grid = div_up(nwarps*32, 256);
	switch(A.bs) {
	case 4: invert_blocks_kernel<4><<<grid,256,0,st>>>(A.nbrows, src_vals, positions, dst, dst_is_compact); break;
	case 5: invert_blocks_kernel<5><<<grid,256,0,st>>>(A.nbrows, src_vals, positions, dst, dst_is_compact); break;
	default: throw Error("block inverse: unsupported block size");
	}
	B200_LAUNCHED();
}

// ------------------------------------------------------------------ diagonal dominance

template <int BS>
__global__ void diag_dom_kernel(const int nbrows, const int *__restrict__ browptr,
                                const int *__restrict__ diagind, const double *__restrict__ vals,
                                double *__restrict__ ddl, double *__restrict__ ddu)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= (long long)nbrows*BS) return;
	const int row = (int)(i / BS), r = (int)(i % BS);
	constexpr int BS2 = BS*BS;
	const int dp = diagind[row];
	double l = 0, u = 0;
	for(int c = 0; c < BS; c++)
		if(c != r) u += fabs(vals[(size_t)dp*BS2 + BlkIO<BS>::at(r, c)]);
	for(int jj = dp+1; jj < browptr[row+1]; jj++)
		for(int c = 0; c < BS; c++) u += fabs(vals[(size_t)jj*BS2 + BlkIO<BS>::at(r, c)]);
	for(int jj = browptr[row]; jj < dp; jj++)
		for(int c = 0; c < BS; c++) l += fabs(vals[(size_t)jj*BS2 + BlkIO<BS>::at(r, c)]);
	ddl[i] = 1.0 - l;
	ddu[i] = 1.0 - u/fabs(vals[(size_t)dp*BS2 + r*BS + r]);
}

void diag_dominance(const Mat& A, const double *vals, double out[4], double * /*d_scratch*/,
                    cudaStream_t st)
{
	const long long n = (long long)A.nbrows*A.bs;
	DevBuf<double> ddl, ddu, red;
	ddl.alloc(n); ddu.alloc(n); red.alloc(4);
	const int grid = div_up(n, 256);
	switch(A.bs) {
	case 1: diag_dom_kernel<1><<<grid,256,0,st>>>(A.nbrows, A.browptr, A.diagind, vals, ddl, ddu); break;
	case 4: diag_dom_kernel<4><<<grid,256,0,st>>>(A.nbrows, A.browptr, A.diagind, vals, ddl, ddu); break;
	case 5: diag_dom_kernel<5><<<grid,256,0,st>>>(A.nbrows, A.browptr, A.diagind, vals, ddl, ddu); break;
	default: throw Error("diag dominance: unsupported block size");
	}
	B200_LAUNCHED();
	size_t tb = 0, t2 = 0;
	cub::DeviceReduce::Sum(nullptr, tb, ddl.p, red.p, (int)n, st);
	cub::DeviceReduce::Min(nullptr, t2, ddl.p, red.p, (int)n, st);
	DevBuf<char> tmp;
	tmp.alloc(std::max(tb, t2));
	size_t t = tmp.n;
	B200_CUDA(cub::DeviceReduce::Sum(tmp.p, t, ddl.p, red.p + 0, (int)n, st)); t = tmp.n;
	B200_CUDA(cub::DeviceReduce::Min(tmp.p, t, ddl.p, red.p + 1, (int)n, st)); t = tmp.n;
	B200_CUDA(cub::DeviceReduce::Sum(tmp.p, t, ddu.p, red.p + 2, (int)n, st)); t = tmp.n;
	B200_CUDA(cub::DeviceReduce::Min(tmp.p, t, ddu.p, red.p + 3, (int)n, st));
	g_launches.fetch_add(4);
	double h[4];
	B200_CUDA(cudaMemcpyAsync(h, red.p, 4*sizeof(double), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	// {lower avg, lower min, upper avg, upper min}  (matrix_properties.cpp:76)
	out[0] = h[0]/(double)n; out[1] = h[1]; out[2] = h[2]/(double)n; out[3] = h[3];
}

}  // namespace b200
