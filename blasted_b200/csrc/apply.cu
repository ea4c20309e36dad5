/** \file apply.cu
 * \brief Asynchronous (block-)triangular solves, SGS sweeps, relaxation and Jacobi application
 * (K5, K6, K7, K8) as one family of row-sweep kernels.
 *
 * Replaces, from the reference:
 *   scalar_unit_lower_triangular / scalar_upper_triangular   src/kernels/kernels_ilu_apply.hpp:15-42
 *   block_unit_lower_triangular / block_upper_triangular     kernels_ilu_apply.hpp:54-94
 *   scalar_fgs / scalar_bgs / block_fgs / block_bgs          src/kernels/kernels_sgs.hpp:17-76
 *   scalar_relax / block_relax_kernel                        src/kernels/kernels_relaxation.hpp:17-54
 *   and the sweep loops around them                          src/solverops_ilu0.cpp:99-141,274-314;
 *                                                            src/solverops_sgs.cpp:62-115,156-202;
 *                                                            src/solverops_levels_{ilu0,sgs}.cpp
 *
 * Every kernel computes, for each (block-)row i of a range,
 *     x_i <- f( rhs_i , sum_{j in part(i)} V_ij x_j )
 * with ONE final store of x_i; x_j are relaxed L2 loads that may see old or new values (chaotic
 * iteration).  CTAs are mapped to rows ascending for lower/forward sweeps and descending for
 * upper/backward sweeps so that the hardware's in-order CTA dispatch gives the Gauss-Seidel-like
 * propagation of the reference's ascending / descending `omp for` loops.  The same kernels, launched
 * per level on an explicit row list, perform the exact level-scheduled substitutions.
 *
 * Mapping: scalar - LPR lanes per row + shuffle reduction; block - bs lanes per block-row, lane r
 * owns row r of every block (blockops.cuh; bs=4: one 256-bit load per block row), x_j is a
 * broadcast load, the bs outputs are one contiguous store.
 * HBM-bound: one L+U pair streams the factor once: (8 b^2 + 4) nnzb + 12 N + 48 b N bytes.
 */
#include "common.cuh"
#include "blockops.cuh"
#include "tma.cuh"

namespace b200 {

__device__ __forceinline__ double ld_iter(const double *p) { return __ldcg(p); }

struct TriDev {
	const int *part_ptr, *part_col;          ///< scalar split form (else nullptr)
	const double *part_diag;
	const int *browptr, *bcolind, *diagind;
	const double *vals, *dinv, *rhs, *rscale, *xsrc;
	double *x;
	const int *rows;
	int row_begin, row_end;
	int descending;
	int max_part_len;                        ///< longest row part of this sweep, 0 = unknown
};

template <int KIND>
__device__ __forceinline__ void part_range(const int s, const int d, const int e, int& js, int& je)
{
	if(KIND == TRI_ILU_LOWER || KIND == TRI_SGS_FWD) { js = s; je = d; }
	else if(KIND == TRI_ILU_UPPER || KIND == TRI_SGS_BWD) { js = d+1; je = e; }
	else { js = s; je = e; }
}

template <int LPR, int KIND>
__global__ void __launch_bounds__(256)
tri_scalar_kernel(const TriDev a)
{
	const long long tid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	const long long t = tid / LPR;
	const int lane = (int)(tid % LPR);
	const int nrows = a.row_end - a.row_begin;
	const bool valid = t < nrows;
	int row = 0, d = 0;
	double sum = 0;
	if(valid) {
		const int idx = a.descending ? a.row_end - 1 - (int)t : a.row_begin + (int)t;
		row = a.rows ? __ldg(a.rows + idx) : idx;
		int js, je;
		const int *cols = a.bcolind;
		if(a.part_ptr) {
			// split form: the part is its own CSR array, nothing to skip
			js = __ldg(a.part_ptr + row); je = __ldg(a.part_ptr + row + 1);
			cols = a.part_col;
			d = -1;
		} else {
			const int s = __ldg(a.browptr + row), e = __ldg(a.browptr + row + 1);
			d = __ldg(a.diagind + row);
			part_range<KIND>(s, d, e, js, je);
		}
		for(int j = js + lane; j < je; j += LPR) {
			if(KIND == TRI_RELAX && j == d) continue;
			sum = fma(__ldg(a.vals + j), ld_iter(a.xsrc + __ldg(cols + j)), sum);
		}
	}
#pragma unroll
	for(int off = LPR/2; off > 0; off >>= 1)
		sum += __shfl_down_sync(0xffffffffu, sum, off, LPR);
	if(valid && lane == 0) {
		double rhs = __ldg(a.rhs + row);
		if(a.rscale) rhs *= __ldg(a.rscale + row);
		double out;
		if(KIND == TRI_ILU_LOWER) out = rhs - sum;
		else if(KIND == TRI_ILU_UPPER)
			out = (1.0/(a.part_ptr ? __ldg(a.part_diag + row) : __ldg(a.vals + d))) * (rhs - sum);
		else if(KIND == TRI_SGS_FWD || KIND == TRI_RELAX) out = __ldg(a.dinv + row) * (rhs - sum);
		else out = rhs - __ldg(a.dinv + row)*sum;       // TRI_SGS_BWD
		a.x[row] = out;
	}
}

template <int BS, int KIND, bool VEC>
__global__ void __launch_bounds__(256, 6)
tri_block_kernel(const TriDev a)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int g = lane / BS, r = lane - g*BS;
	const long long t = warp*GPW + g;
	const int nrows = a.row_end - a.row_begin;
	const bool valid = (g < GPW) && (t < nrows);
	int row = 0, d = 0;
	double acc = 0, rhs = 0;
	double dr[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) dr[c] = 0;
	if(valid) {
		const int idx = a.descending ? a.row_end - 1 - (int)t : a.row_begin + (int)t;
		row = a.rows ? __ldg(a.rows + idx) : idx;
		int js, je;
		const int *cols = a.bcolind;
		if(a.part_ptr) {
			// split factor: the part is its own block CSR array, nothing to skip
			js = __ldg(a.part_ptr + row); je = __ldg(a.part_ptr + row + 1);
			cols = a.part_col;
			d = -1;
		} else {
			const int s = __ldg(a.browptr + row), e = __ldg(a.browptr + row + 1);
			d = __ldg(a.diagind + row);
			part_range<KIND>(s, d, e, js, je);
		}
		// everything that depends only on the row is requested before the block loop
		rhs = __ldg(a.rhs + (size_t)row*BS + r);
		if(a.rscale) rhs *= __ldg(a.rscale + (size_t)row*BS + r);
		if(KIND != TRI_ILU_LOWER)
			// compact inverted diagonal blocks: U_ii^-1 (ILU) or D_i^-1 of A (SGS, relaxation)
			BlkIO<BS>::template load_row<false>(a.dinv + (size_t)row*BS2, r, dr);
#pragma unroll 2
		for(int jj = js; jj < je; jj++) {
			if(KIND == TRI_RELAX && jj == d) continue;
			const int col = __ldg(cols + jj);
			double av[BS], xv[BS];
			BlkIO<BS>::template load_row<false>(a.vals + (size_t)jj*BS2, r, av);
			load_seg<BS,true,VEC>(a.xsrc + (size_t)col*BS, xv);
#pragma unroll
			for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
		}
	}
	double out;
	if(KIND == TRI_ILU_LOWER) out = rhs - acc;
	else {
		// multiply a bs-vector held one entry per lane by a bs x bs block: t_c via shuffles
		const double tv = (KIND == TRI_SGS_BWD) ? acc : rhs - acc;
		double prod = 0;
#pragma unroll
		for(int c = 0; c < BS; c++) {
			const double tc = __shfl_sync(0xffffffffu, tv, g*BS + c);
			prod = fma(dr[c], tc, prod);
		}
		out = (KIND == TRI_SGS_BWD) ? rhs - prod : prod;
	}
	if(valid) a.x[(size_t)row*BS + r] = out;
}

/// Persistent, software-pipelined form of tri_block_kernel for the plain asynchronous sweeps (whole
/// row range, no level list).  A row's loads form a chain  row indices -> column indices -> x_j;
/// with two or three blocks per part the chain, not the bandwidth, bounds the one-shot kernel.
/// Here every group walks rows t, t+G, t+2G, ... and keeps the indices of its row after next and the
/// column indices of its next row in registers, so that in the steady state every load of the
/// current row - block rows, vector segments, right-hand side, diagonal block - has a known
/// address and is requested at once.  Arithmetic and the single final store per row are those of
/// tri_block_kernel; rows are still visited in ascending (descending) waves.
template <int BS, int KIND, bool VEC>
__global__ void __launch_bounds__(256, 4)
tri_block_pipe_kernel(const TriDev a)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	constexpr int K = 3;                       // column indices prefetched per row part
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const int wpc = blockDim.x >> 5;
	const long long stride = (long long)gridDim.x*wpc*GPW;
	const long long wbase = ((long long)blockIdx.x*wpc + (threadIdx.x >> 5))*GPW;
	const int nrows = a.row_end - a.row_begin;

	const int *const cols_arr = a.part_ptr ? a.part_col : a.bcolind;
	auto load_meta = [&](const long long t, int& js, int& je, int& d) {
		js = 0; je = 0; d = -1;
		if(g < GPW && t < nrows) {
			const int row = a.descending ? a.row_end - 1 - (int)t : a.row_begin + (int)t;
			if(a.part_ptr) {
				// split factor: the part is its own block CSR array
				js = __ldg(a.part_ptr + row); je = __ldg(a.part_ptr + row + 1);
			} else {
				const int s = __ldg(a.browptr + row), e = __ldg(a.browptr + row + 1);
				d = __ldg(a.diagind + row);
				part_range<KIND>(s, d, e, js, je);
			}
		}
	};
	auto load_cols = [&](const int js, const int je, int (&c)[K]) {
#pragma unroll
		for(int q = 0; q < K; q++) c[q] = (js + q < je) ? __ldg(cols_arr + js + q) : 0;
	};

	int js1, je1, d1, c1[K], js2, je2, d2;
	load_meta(wbase + g, js1, je1, d1);
	load_cols(js1, je1, c1);
	load_meta(wbase + g + stride, js2, je2, d2);

	for(long long tw = wbase; tw < nrows; tw += stride) {
		const long long t = tw + g;
		const int js = js1, je = je1, d = d1;
		int cols[K];
#pragma unroll
		for(int q = 0; q < K; q++) cols[q] = c1[q];
		js1 = js2; je1 = je2; d1 = d2;
		load_cols(js1, je1, c1);
		load_meta(t + 2*stride, js2, je2, d2);

		const bool valid = (g < GPW) && (t < nrows);
		const int row = valid ? (a.descending ? a.row_end - 1 - (int)t : a.row_begin + (int)t) : 0;
		double acc = 0, rhs = 0;
		double dr[BS];
#pragma unroll
		for(int c = 0; c < BS; c++) dr[c] = 0;
		if(valid) {
			rhs = __ldg(a.rhs + (size_t)row*BS + r);
			if(a.rscale) rhs *= __ldg(a.rscale + (size_t)row*BS + r);
			if(KIND != TRI_ILU_LOWER)
				BlkIO<BS>::template load_row<false>(a.dinv + (size_t)row*BS2, r, dr);
#pragma unroll
			for(int q = 0; q < K; q++) {
				const int jj = js + q;
				if(jj < je && !(KIND == TRI_RELAX && jj == d)) {
					double av[BS], xv[BS];
					BlkIO<BS>::template load_row<false>(a.vals + (size_t)jj*BS2, r, av);
					load_seg<BS,true,VEC>(a.xsrc + (size_t)cols[q]*BS, xv);
#pragma unroll
					for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
				}
			}
			for(int jj = js + K; jj < je; jj++) {
				if(KIND == TRI_RELAX && jj == d) continue;
				const int col = __ldg(cols_arr + jj);
				double av[BS], xv[BS];
				BlkIO<BS>::template load_row<false>(a.vals + (size_t)jj*BS2, r, av);
				load_seg<BS,true,VEC>(a.xsrc + (size_t)col*BS, xv);
#pragma unroll
				for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
			}
		}
		double out;
		if(KIND == TRI_ILU_LOWER) out = rhs - acc;
		else {
			const double tv = (KIND == TRI_SGS_BWD) ? acc : rhs - acc;
			double prod = 0;
#pragma unroll
			for(int c = 0; c < BS; c++) {
				const double tc = __shfl_sync(0xffffffffu, tv, min(g*BS + c, 31));
				prod = fma(dr[c], tc, prod);
			}
			out = (KIND == TRI_SGS_BWD) ? rhs - prod : prod;
		}
		if(valid) a.x[(size_t)row*BS + r] = out;
	}
}

// ------------------------------------------------------------------ bs = 5 sweeps, staged
//
// tri_block_kernel<5> is latency-bound (ncu: L1 data pipe 58-64 %, long-scoreboard stalls): a group
// reads every 200-byte block as five 40-byte segments and every x segment as five broadcast loads.
// Where no row part has more than three blocks (every face-neighbour block stencil) the sweep is
// staged like the factor launches (factor.cu, "staged"): a warp owns six consecutive rows per
// iteration, lane 0 of every group copies the row's run of blocks and lane 1 the inverted diagonal
// block into shared memory with TMA bulk copies on the warp's mbarrier, and while they fly the lanes
// fetch the column indices, ONE x entry per lane and block (lane c needs x_c only: it multiplies
// column c of every block, read as 40 contiguous bytes from shared memory) and the right-hand
// side.  The five partial vectors of a group are summed once per row (4 shuffles per lane), the
// product with the inverted diagonal block is formed the same way.  One final store per row, x
// gathered with relaxed L2 loads, rows ascending (lower/forward) or descending (upper/backward):
// the chaotic-iteration rules of the generic kernel.  Summation order differs from the generic
// kernel at rounding level only.

/// y_r = sum over the group's lanes c of v_c[r]  (v = this lane's partial vector)
template <int BS>
__device__ __forceinline__ double group_reduce_columns(const double (&v)[BS], const int gbase, const int r)
{
	double y = pick<BS>(v, r);
#pragma unroll
	for(int t = 1; t < BS; t++) {
		const double x = pick<BS>(v, (r - t + BS) % BS);       // what lane (r-t) wants from this lane
		y += __shfl_sync(0xffffffffu, x, min(gbase + (r + t) % BS, 31));
	}
	return y;
}

template <int KIND>
__global__ void __launch_bounds__(128)
tri5_staged_kernel(const TriDev a)
{
	constexpr int BS = 5, GPW = 6, BS2 = 25, K = 3;
	// the six rows of a warp are consecutive, so their parts are ONE contiguous span of the (split
	// or matrix-ordered) value array when the parts are whole rows of a split array, and their
	// inverted diagonal blocks one span of the compact array: two copies per warp and iteration
	constexpr int SLOT_BYTES = K*BS2*8 + 24;                                               // 624: one row's run
	constexpr int RUN_BYTES = GPW*SLOT_BYTES, D_BYTES = GPW*BS2*8 + 16;                    // 3744, 1216
	constexpr int STAGE_BYTES = RUN_BYTES + D_BYTES;                                       // 4960
	constexpr bool NEED_D = (KIND != TRI_ILU_LOWER);
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bars[4];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int g = lane / BS, r = lane - g*BS;
	unsigned char *const wsm = smem_raw + (size_t)w*STAGE_BYTES;
	if(lane == 0) mbar_init(&bars[w], 1);
	__syncwarp();
	const long long warp = (long long)blockIdx.x*(blockDim.x >> 5) + w;
	const long long stride = (long long)gridDim.x*(blockDim.x >> 5)*GPW;
	const int nrows = a.row_end - a.row_begin;
	const int *const cols_arr = a.part_ptr ? a.part_col : a.bcolind;
	const bool spans = a.part_ptr != nullptr;            // parts of consecutive rows are adjacent
	auto rowof = [&](const long long tt) { return a.descending ? a.row_end - 1 - (int)tt : a.row_begin + (int)tt; };
	auto load_meta = [&](const long long tt, int& js, int& je) {
		js = 0; je = 0;
		if(g < GPW && tt < nrows) {
			const int row = rowof(tt);
			if(a.part_ptr) { js = __ldg(a.part_ptr + row); je = __ldg(a.part_ptr + row + 1); }
			else {
				const int s = __ldg(a.browptr + row), e = __ldg(a.browptr + row + 1);
				part_range<KIND>(s, __ldg(a.diagind + row), e, js, je);
			}
		}
	};
	long long t = warp*GPW + g;
	int js0, je0, js1, je1;
	load_meta(t, js0, je0);
	load_meta(t + stride, js1, je1);
	const long long niter = ((long long)nrows + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		int js2, je2;
		load_meta(t + 2*stride, js2, je2);
		const bool valid = (g < GPW) && (t < nrows);
		const int row = valid ? rowof(t) : 0;
		const int nb = je0 - js0;
		// spans of the warp's rows
		const int first = __reduce_min_sync(0xffffffffu, valid ? js0 : 0x7fffffff);
		const int last = __reduce_max_sync(0xffffffffu, valid ? je0 : 0);
		const int rmin = __reduce_min_sync(0xffffffffu, valid ? row : 0x7fffffff);
		const int rmax = __reduce_max_sync(0xffffffffu, valid ? row : -1);
		unsigned myoff = 0;                                  // byte offset of this group's run in the staged area
		{
			unsigned bytes = 0, dstoff = 0;
			size_t src = 0;
			if(spans) {
				const size_t s0 = (size_t)(a.vals + (size_t)first*BS2);
				myoff = (unsigned)((size_t)(js0 - first)*BS2*8 + (s0 & 15));
				if(lane == 0 && last > first) { src = s0; bytes = ((unsigned)((last - first)*BS2*8) + (unsigned)(s0 & 15) + 15u) & ~15u; }
			} else {
				// matrix-ordered values (SGS on A): one run per row, K blocks of room each
				const size_t s0 = (size_t)(a.vals + (size_t)js0*BS2);
				if(valid && r == 0 && nb > 0) {
					src = s0; dstoff = (unsigned)(g*SLOT_BYTES);
					bytes = ((unsigned)(nb*BS2*8) + (unsigned)(s0 & 15) + 15u) & ~15u;
				}
				myoff = (unsigned)(g*SLOT_BYTES) + (unsigned)(s0 & 15);
			}
			if(NEED_D && lane == 1 && rmax >= rmin) {
				src = (size_t)(a.dinv + (size_t)rmin*BS2); dstoff = RUN_BYTES;
				bytes = ((unsigned)((rmax - rmin + 1)*BS2*8) + (unsigned)(src & 15) + 15u) & ~15u;
			}
			const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
			if(lane == 0) mbar_expect_tx(&bars[w], total);
			if(bytes) bulk_g2s(wsm + dstoff, (const void*)(src & ~(size_t)15), bytes, &bars[w]);
		}
		// while the copies fly: column indices, this lane's x entries, the right-hand side
		double xk[K];
#pragma unroll
		for(int k = 0; k < K; k++) {
			xk[k] = 0;
			if(valid && k < nb) xk[k] = ld_iter(a.xsrc + (size_t)__ldg(cols_arr + js0 + k)*BS + r);
		}
		double rhs = 0;
		if(valid) {
			rhs = __ldg(a.rhs + (size_t)row*BS + r);
			if(a.rscale) rhs *= __ldg(a.rscale + (size_t)row*BS + r);
		}
		mbar_wait(&bars[w], (unsigned)(it & 1));

		const double *sv = reinterpret_cast<const double*>(wsm + myoff);
		double acc[BS];
#pragma unroll
		for(int i = 0; i < BS; i++) acc[i] = 0;
		if(valid) {
#pragma unroll
			for(int k = 0; k < K; k++)
				if(k < nb) {
#pragma unroll
					for(int i = 0; i < BS; i++) acc[i] = fma(sv[k*BS2 + r*BS + i], xk[k], acc[i]);   // column r of block k
				}
		}
		const double y = group_reduce_columns<BS>(acc, g*BS, r);          // (sum_j V_ij x_j)_r
		double out;
		if(KIND == TRI_ILU_LOWER) out = rhs - y;
		else {
			const double tv = (KIND == TRI_SGS_BWD) ? y : rhs - y;
			const size_t d0 = (size_t)(a.dinv + (size_t)(rmax >= rmin ? rmin : 0)*BS2);
			const double *sd = reinterpret_cast<const double*>(wsm + RUN_BYTES + (d0 & 15)) + (size_t)(valid ? row - rmin : 0)*BS2;
			double p[BS];
#pragma unroll
			for(int i = 0; i < BS; i++) p[i] = valid ? sd[r*BS + i]*tv : 0.0;   // column r of D^-1 times t_r
			const double prod = group_reduce_columns<BS>(p, g*BS, r);
			out = (KIND == TRI_SGS_BWD) ? rhs - prod : prod;
		}
		if(valid) a.x[(size_t)row*BS + r] = out;
		__syncwarp();                       // everybody has read the stage before it is refilled
		js0 = js1; je0 = je1; js1 = js2; je1 = je2;
		t += stride;
	}
}

template <int KIND>
static void launch_tri5_staged(const TriDev& d, cudaStream_t st)
{
	constexpr int SMEM = 4*(6*624 + 6*25*8 + 16);
	static int grid = 0;
	if(!grid) {
		int dev = 0, sms = 148, per = 1;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, tri5_staged_kernel<KIND>, 128, SMEM) != cudaSuccess || per < 1)
			per = 1;
		grid = sms*per;
	}
	const long long nrows = d.row_end - d.row_begin;
	const int g = (int)std::max<long long>(1, std::min<long long>(grid, (nrows + 23)/24));
	tri5_staged_kernel<KIND><<<g, 128, SMEM, st>>>(d);
}

/// Resident CTAs of a persistent kernel over all SMs (cached per kernel), capped by the work
template <typename Kern>
static int resident_grid(Kern kernel, long long rows_per_cta, long long nrows)
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
	const long long need = (nrows + rows_per_cta - 1)/rows_per_cta;
	return (int)std::max<long long>(1, std::min<long long>(cached, need));
}

template <int KIND>
static void launch_kind(const Mat& A, const TriDev& d, const double avg_part, cudaStream_t st)
{
	const long long nrows = d.row_end - d.row_begin;
	if(nrows <= 0) return;
	static const bool one_shot = getenv("B200_TRI1") != nullptr;      // A/B switches (development)
	static const bool no_staged = getenv("B200_NO_STAGED") != nullptr || getenv("B200_NO_STAGED_TRI") != nullptr;
	if(A.bs == 1) {
#define B200_TRI_CASE(L)                                                           \
		{                                                                          \
			const int grid = div_up(nrows*L, 256);                                 \
			tri_scalar_kernel<L,KIND><<<grid, 256, 0, st>>>(d);                    \
		}
		if(avg_part <= 2.5) B200_TRI_CASE(2)
		else if(avg_part <= 5) B200_TRI_CASE(4)
		else if(avg_part <= 10) B200_TRI_CASE(8)
		else if(avg_part <= 20) B200_TRI_CASE(16)
		else B200_TRI_CASE(32)
#undef B200_TRI_CASE
	}
	else if(A.bs == 4 && !d.rows && aligned32(d.xsrc) && !one_shot) {
		auto k = tri_block_pipe_kernel<4,KIND,true>;
		k<<<resident_grid(k, 64, nrows), 256, 0, st>>>(d);
	}
	// (bs = 5 measured slower pipelined - 64 registers cost more occupancy than the chain costs
	// time with 3-block parts - and stays on the one-shot kernel.)
	else if(A.bs == 4) {
		const long long nwarps = (nrows + 7)/8;
		if(aligned32(d.xsrc))
			tri_block_kernel<4,KIND,true><<<div_up(nwarps*32, 256), 256, 0, st>>>(d);
		else
			tri_block_kernel<4,KIND,false><<<div_up(nwarps*32, 256), 256, 0, st>>>(d);
	}
	else if(A.bs == 5 && KIND != TRI_RELAX && !d.rows && d.max_part_len >= 1 && d.max_part_len <= 3 && !no_staged) {
		launch_tri5_staged<KIND>(d, st);
	}
	else if(A.bs == 5) {
		const long long nwarps = (nrows + 5)/6;
		tri_block_kernel<5,KIND,false><<<div_up(nwarps*32, 256), 256, 0, st>>>(d);
	}
	else throw Error("triangular sweep: unsupported block size " + std::to_string(A.bs));
	B200_LAUNCHED();
}

void launch_tri_sweep(const Mat& A, TriKind kind, const TriArgs& a, cudaStream_t st)
{
	TriDev d;
	d.part_ptr = a.part_ptr; d.part_col = a.part_col; d.part_diag = a.part_diag;
	d.browptr = A.browptr; d.bcolind = A.bcolind; d.diagind = A.diagind;
	d.vals = a.vals; d.dinv = a.dinv; d.rhs = a.rhs; d.rscale = a.rscale;
	d.xsrc = a.xsrc ? a.xsrc : a.x; d.x = a.x; d.rows = a.rows;
	d.row_begin = a.row_begin; d.row_end = a.row_end; d.descending = a.descending ? 1 : 0;
	d.max_part_len = a.max_part_len;
	const double half = 0.5*(A.avg_row_len - 1.0);
	ProfScope ps((kind == TRI_ILU_LOWER || kind == TRI_SGS_FWD) ? KC_TRI_LOWER :
	             (kind == TRI_ILU_UPPER || kind == TRI_SGS_BWD) ? KC_TRI_UPPER : KC_OTHER, st);
	switch(kind) {
	case TRI_ILU_LOWER: launch_kind<TRI_ILU_LOWER>(A, d, half, st); break;
	case TRI_ILU_UPPER: launch_kind<TRI_ILU_UPPER>(A, d, half, st); break;
	case TRI_SGS_FWD: launch_kind<TRI_SGS_FWD>(A, d, half, st); break;
	case TRI_SGS_BWD: launch_kind<TRI_SGS_BWD>(A, d, half, st); break;
	case TRI_RELAX: launch_kind<TRI_RELAX>(A, d, A.avg_row_len, st); break;
	}
}

// ------------------------------------------------------------------ exact substitution, one launch

/// polled read of a value another CTA may be about to publish (never hoisted, read at L2)
__device__ __forceinline__ double ld_poll(const double *p)
{
	double v;
	asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ bool is_unset(const double v)
{
	return (unsigned long long)__double_as_longlong(v) == 0xffffffffffffffffull;
}

/// Exact (level-ordered) triangular substitution in ONE launch instead of one launch per level.
/// Rows are taken in the order of the level-sorted row list; the output vector is pre-filled with
/// an all-ones bit pattern and every value doubles as its own "ready" flag: a row consumes its
/// dependencies x_j in column order as soon as they stop being the sentinel (polled at L2), and
/// publishes x_i with ordinary 8-byte stores.  All dependencies of a row sit earlier in the list;
/// CTAs take their position from a ticket counter, so every CTA a row can wait for has already
/// started - no deadlock - and rows deep in finished levels never wait at all.  Nothing else
/// changes: same per-row arithmetic in the same column order as tri_scalar/tri_block_kernel, one
/// final store per value.  The spin is bounded: past the cap a warp raises *err and stops waiting
/// (a dependency pointing forward in the list - a pattern the level builder should have refused).
template <int BS, int KIND, bool VEC, int DK>    // DK: dependencies held in registers and polled together
__global__ void __launch_bounds__(256)
tri_syncfree_kernel(const TriDev a, int *__restrict__ ticket, int *__restrict__ err)
{
	constexpr int GPW = 32/BS;
	constexpr int BS2 = BS*BS;
	__shared__ int s_cta;
	if(threadIdx.x == 0) s_cta = atomicAdd(ticket, 1);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int g = lane / BS, r = lane - g*BS;
	const long long warp = (long long)s_cta*(blockDim.x >> 5) + (threadIdx.x >> 5);
	const long long t = warp*GPW + g;
	const int nrows = a.row_end - a.row_begin;
	const bool valid = (g < GPW) && (t < nrows);
	int row = 0, cbase = 0, je = 0;
	const int *cols = a.bcolind;
	double acc = 0, rhs = 0;
	double dr[BS];
#pragma unroll
	for(int c = 0; c < BS; c++) dr[c] = 0;
	if(valid) {
		const int idx = a.descending ? a.row_end - 1 - (int)t : a.row_begin + (int)t;
		row = a.rows ? __ldg(a.rows + idx) : idx;
		if(a.part_ptr) {
			cbase = __ldg(a.part_ptr + row); je = __ldg(a.part_ptr + row + 1);
			cols = a.part_col;
		} else {
			const int s = __ldg(a.browptr + row), e = __ldg(a.browptr + row + 1);
			const int d = __ldg(a.diagind + row);
			part_range<KIND>(s, d, e, cbase, je);
		}
	}
	// the current chunk of up to DK dependencies: column indices and matrix entries in registers,
	// loaded before any waiting so that the hand-over from one level to the next costs L2 round
	// trips only (polls of x), never a DRAM miss on the row's own data
	int pc[DK];
	double pv[DK][BS];
	int nb = 0, adv = 0;
	auto load_chunk = [&]() {
		nb = min(DK, je - cbase);
		adv = 0;
#pragma unroll
		for(int q = 0; q < DK; q++) {
			pc[q] = 0;
#pragma unroll
			for(int c = 0; c < BS; c++) pv[q][c] = 0;
			if(q < nb) {
				pc[q] = __ldg(cols + cbase + q);
				if(BS == 1) pv[q][0] = __ldg(a.vals + cbase + q);
				else BlkIO<BS>::template load_row<false>(a.vals + (size_t)(cbase + q)*BS2, r, pv[q]);
			}
		}
	};
	load_chunk();
	if(valid) {
		rhs = __ldg(a.rhs + (size_t)row*BS + r);
		if(a.rscale) rhs *= __ldg(a.rscale + (size_t)row*BS + r);
		if(BS > 1 && KIND != TRI_ILU_LOWER)
			BlkIO<BS>::template load_row<false>(a.dinv + (size_t)row*BS2, r, dr);
		else if(BS == 1 && KIND != TRI_ILU_LOWER) {
			if(KIND == TRI_ILU_UPPER)
				dr[0] = 1.0/(a.part_ptr ? __ldg(a.part_diag + row) : __ldg(a.vals + __ldg(a.diagind + row)));
			else dr[0] = __ldg(a.dinv + row);
		}
	}
	bool stored = !valid;
	int spins = 0;
	while(true) {
		// While a warp waits, only its first unfinished lane polls (its next dependency): the rows
		// of a warp are level-sorted, so that lane becomes ready first, and tens of thousands of
		// waiting rows polling all their dependencies saturate L2 and stretch every hand-over
		// (a sleep between polls was measured too: 0 / 32 / 100 / 300 ns make no difference or lose).
		const unsigned waiting = __ballot_sync(0xffffffffu, !stored && cbase < je);
		if(waiting) {
			const int leader = __ffs(waiting) - 1;
			int ready = 1;
			if(lane == leader) {
				int col = pc[0];
#pragma unroll
				for(int q = 1; q < DK; q++) if(q == adv) col = pc[q];
				ready = !is_unset(ld_poll(a.xsrc + (size_t)col*BS + ((BS > 1) ? r : 0)));
			}
			ready = __shfl_sync(0xffffffffu, ready, leader);
			const bool somebody_done = __any_sync(0xffffffffu, !stored && cbase >= je);
			if(!ready && !somebody_done) {
				if(++spins > (1 << 16) && ((spins & 1023) == 0)) {
					if(spins > (1 << 22) || *((volatile int*)err)) { *err = 1; cbase = je; }
				}
				continue;
			}
		}
		if(!stored && cbase < je) {
			// poll the rest of the chunk together and consume the ready prefix in column order
			double xv[DK][BS];
#pragma unroll
			for(int q = 0; q < DK; q++) {
#pragma unroll
				for(int c = 0; c < BS; c++) xv[q][c] = 0;
				if(q >= adv && q < nb) {
					if(BS == 4 && VEC)
						ld256_cg_ordered(a.xsrc + (size_t)pc[q]*BS, xv[q][0], xv[q][1], xv[q][2], xv[q][3]);
					else {
#pragma unroll
						for(int c = 0; c < BS; c++) xv[q][c] = ld_poll(a.xsrc + (size_t)pc[q]*BS + c);
					}
				}
			}
#pragma unroll
			for(int q = 0; q < DK; q++) {
				bool unset = false;
#pragma unroll
				for(int c = 0; c < BS; c++) unset |= is_unset(xv[q][c]);
				if(q == adv && q < nb && !unset) {
#pragma unroll
					for(int c = 0; c < BS; c++) acc = fma(pv[q][c], xv[q][c], acc);
					adv++;
				}
			}
			if(adv == nb) {
				cbase += nb;
				if(cbase < je) load_chunk();
			}
		}
		const bool fin = !stored && cbase >= je;
		// epilogue with the whole warp converged (the block forms exchange values by shuffles)
		double out;
		if(KIND == TRI_ILU_LOWER) out = rhs - acc;
		else if(BS == 1) out = (KIND == TRI_SGS_BWD) ? rhs - dr[0]*acc : dr[0]*(rhs - acc);
		else {
			const double tv = (KIND == TRI_SGS_BWD) ? acc : rhs - acc;
			double prod = 0;
#pragma unroll
			for(int c = 0; c < BS; c++) {
				const double tc = __shfl_sync(0xffffffffu, tv, min(g*BS + c, 31));
				prod = fma(dr[c], tc, prod);
			}
			out = (KIND == TRI_SGS_BWD) ? rhs - prod : prod;
		}
		if(fin) {
			if(is_unset(out)) out = __longlong_as_double(0x7ff8000000000000ll);
			a.x[(size_t)row*BS + r] = out;
			stored = true;
		}
		if(__all_sync(0xffffffffu, stored)) break;
	}
}

template <int KIND>
static void launch_syncfree_kind(const Mat& A, const TriDev& d, int *ticket, int *err, cudaStream_t st)
{
	const long long nrows = d.row_end - d.row_begin;
	if(nrows <= 0) return;
	B200_CUDA(cudaMemsetAsync(ticket, 0, sizeof(int), st));
	B200_CUDA(cudaMemsetAsync(d.x, 0xff, (size_t)A.dim()*sizeof(double), st));
	if(A.bs == 1) {
		// row parts of up to 8 / up to 16 entries in one chunk (7-point: 3, 27-point: 13)
		if(A.avg_row_len <= 17.0)
			tri_syncfree_kernel<1,KIND,false,8><<<div_up(nrows, 256), 256, 0, st>>>(d, ticket, err);
		else
			tri_syncfree_kernel<1,KIND,false,16><<<div_up(nrows, 256), 256, 0, st>>>(d, ticket, err);
	}
	else if(A.bs == 4) {
		const long long nwarps = (nrows + 7)/8;
		if(aligned32(d.xsrc))
			tri_syncfree_kernel<4,KIND,true,3><<<div_up(nwarps*32, 256), 256, 0, st>>>(d, ticket, err);
		else
			tri_syncfree_kernel<4,KIND,false,3><<<div_up(nwarps*32, 256), 256, 0, st>>>(d, ticket, err);
	}
	else if(A.bs == 5) {
		const long long nwarps = (nrows + 5)/6;
		tri_syncfree_kernel<5,KIND,false,3><<<div_up(nwarps*32, 256), 256, 0, st>>>(d, ticket, err);
	}
	else throw Error("triangular solve: unsupported block size " + std::to_string(A.bs));
	B200_LAUNCHED();
}

void launch_tri_syncfree(const Mat& A, TriKind kind, const TriArgs& a, int *ticket, int *err,
                         cudaStream_t st)
{
	if(a.x == a.rhs) throw Error("triangular solve: output aliases the right-hand side");
	TriDev d;
	d.part_ptr = a.part_ptr; d.part_col = a.part_col; d.part_diag = a.part_diag;
	d.browptr = A.browptr; d.bcolind = A.bcolind; d.diagind = A.diagind;
	d.vals = a.vals; d.dinv = a.dinv; d.rhs = a.rhs; d.rscale = a.rscale;
	d.xsrc = a.x; d.x = a.x; d.rows = a.rows;
	d.row_begin = a.row_begin; d.row_end = a.row_end; d.descending = a.descending ? 1 : 0;
	d.max_part_len = 0;
	ProfScope ps((kind == TRI_ILU_LOWER || kind == TRI_SGS_FWD) ? KC_TRI_LOWER : KC_TRI_UPPER, st);
	switch(kind) {
	case TRI_ILU_LOWER: launch_syncfree_kind<TRI_ILU_LOWER>(A, d, ticket, err, st); break;
	case TRI_ILU_UPPER: launch_syncfree_kind<TRI_ILU_UPPER>(A, d, ticket, err, st); break;
	case TRI_SGS_FWD: launch_syncfree_kind<TRI_SGS_FWD>(A, d, ticket, err, st); break;
	case TRI_SGS_BWD: launch_syncfree_kind<TRI_SGS_BWD>(A, d, ticket, err, st); break;
	default: throw Error("triangular solve: relaxation has no one-launch form");
	}
}

// ------------------------------------------------------------------ Jacobi application

template <int BS>
__global__ void __launch_bounds__(256)
jacobi_apply_kernel(const int nbrows, const double *__restrict__ dinv, const double *__restrict__ rr,
                    double *__restrict__ z)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= (long long)nbrows*BS) return;
	const long long row = i / BS;
	const int r = (int)(i - row*BS);
	double s = 0;
#pragma unroll
	for(int c = 0; c < BS; c++)
		s = fma(__ldg(dinv + row*BS*BS + BlkIO<BS>::at(r, c)), __ldg(rr + row*BS + c), s);
	z[i] = s;
}

void launch_jacobi_apply(const Mat& A, const double *dinv, const double *r, double *z, cudaStream_t st)
{
	const long long n = (long long)A.nbrows*A.bs;
	if(n == 0) return;
	const int grid = div_up(n, 256);
	switch(A.bs) {
	case 1: jacobi_apply_kernel<1><<<grid,256,0,st>>>(A.nbrows, dinv, r, z); break;
	case 4: jacobi_apply_kernel<4><<<grid,256,0,st>>>(A.nbrows, dinv, r, z); break;
	case 5: jacobi_apply_kernel<5><<<grid,256,0,st>>>(A.nbrows, dinv, r, z); break;
	default: throw Error("Jacobi: unsupported block size");
	}
	B200_LAUNCHED();
}

// ------------------------------------------------------------------ small vector kernels

__global__ void vec_scale_copy_kernel(const long long n, const double *__restrict__ scale,
                                      const double *in, double *out)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) out[i] = scale ? scale[i]*in[i] : in[i];
}

void launch_vec_scale_copy(long long n, const double *scale, const double *in, double *out,
                           cudaStream_t st)
{
	if(n == 0) return;
	vec_scale_copy_kernel<<<div_up(n,256),256,0,st>>>(n, scale, in, out);
	B200_LAUNCHED();
}

void launch_vec_fill(long long n, double v, double *out, cudaStream_t st)
{
	if(n == 0) return;
	if(v == 0.0) { B200_CUDA(cudaMemsetAsync(out, 0, n*sizeof(double), st)); return; }
	throw Error("vec_fill: only zero supported");
}

}  // namespace b200
