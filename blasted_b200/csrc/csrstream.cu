/** \file csrstream.cu
 * \brief Scalar (bs=1) row sweeps as a staged stream: SpMV, gemv3 and the asynchronous triangular
 * sweeps over the split L / U parts of the scalar ILU(0) factor.
 *
 * Replaces the same reference loops as spmv.cu / apply.cu (src/blas/matvecs.cpp:78-108,
 * src/kernels/kernels_ilu_apply.hpp:15-42, src/solverops_ilu0.cpp:274-314) for short scalar rows,
 * where a lanes-per-row mapping leaves too few bytes in flight per warp to cover HBM latency.
 *
 * One CTA owns a tile of R consecutive rows (R*max_row_len <= CAP entries).  The tile's CONTIGUOUS
 * spans of values and column indices are staged in shared memory by two 1D bulk asynchronous
 * copies (TMA, cp.async.bulk completing on an mbarrier): up to 48 KiB in flight per CTA without a
 * single register, so a few resident CTAs per SM cover the HBM latency.  Then one thread per row
 * (or LPR lanes per long row) reads its row from shared memory, gathers the vector entries - for
 * consecutive rows of a stencil matrix these gathers are consecutive addresses - and writes the
 * single final value.
 * Tiles are issued in ascending row order for lower/forward sweeps and descending order for
 * upper/backward sweeps (same propagation argument as apply.cu).
 *
 * HBM-bound: 12 B per stored entry + 4 B row pointer + vectors.
 */
#include "common.cuh"
#include "tma.cuh"

namespace b200 {

constexpr int STREAM_CAP = 3584;       // entries staged per CTA: 28 KiB of values + 14 KiB of indices

template <int KIND, int LPR, int RPT, int CAP, bool CHAIN>
__global__ void __launch_bounds__(256)
csr_stream_kernel(const StreamArgs a)
{
	// R rows per tile: short rows -> several rows per thread (RPT), long rows -> several lanes per
	// row (LPR); exactly one of LPR, RPT is > 1
	constexpr int R = 256*RPT/LPR;
	__shared__ __align__(16) double sval[CAP + 2];
	__shared__ __align__(16) int scol[CAP + 4];
	__shared__ int sptr[R + 1];
	__shared__ __align__(8) unsigned long long bar;

	const int ntiles = gridDim.x;
	const int tile = a.descending ? ntiles - 1 - (int)blockIdx.x : (int)blockIdx.x;
	const int r0 = a.row_begin + tile*R;
	const int nr = min(R, a.row_end - r0);
	const int tid = threadIdx.x;

	for(int i = tid; i <= nr; i += 256) sptr[i] = __ldg(a.ptr + r0 + i);
	if(tid == 0) mbar_init(&bar, 1);

	// right-hand sides of this thread's rows: issued now, consumed at the end
	double rhs[RPT];
	if(KIND != STREAM_SPMV) {
#pragma unroll
		for(int k = 0; k < RPT; k++) {
			const int lr = (LPR > 1) ? tid/LPR : tid + k*256;
			rhs[k] = 0;
			if(lr < nr) {
				const int row = r0 + lr;
				if(KIND == STREAM_GEMV3) rhs[k] = a.beta*a.yin[row];
				else {
					rhs[k] = __ldg(a.rhs + row);
					if(a.rscale) rhs[k] *= __ldg(a.rscale + row);
				}
			}
		}
	}
	__syncthreads();
	const int e0 = sptr[0], e1 = sptr[nr];

	// stage the tile's contiguous spans of values and column indices with two bulk async copies
	// (sources rounded down to 16-byte alignment; the arrays are padded at allocation)
	const int v0 = e0 & ~1, c0 = e0 & ~3;
	if(e1 > e0) {
		if(tid == 0) {
			const unsigned vbytes = (unsigned)(((e1 - v0)*8 + 15) & ~15);
			const unsigned cbytes = (unsigned)(((e1 - c0)*4 + 15) & ~15);
			mbar_expect_tx(&bar, vbytes + cbytes);
			bulk_g2s(sval, a.val + v0, vbytes, &bar);
			bulk_g2s(scol, a.col + c0, cbytes, &bar);
		}
		mbar_wait(&bar, 0);
	}
	const double *tv = sval + (e0 - v0) - e0;       // tv[e] = value of entry e
	const int *tc = scol + (e0 - c0) - e0;

	// what a row's new value is, given the sum over its stored entries
	auto finish = [&](const int k, const int row, const double sum) -> double {
		if(KIND == STREAM_SPMV) return sum;
		else if(KIND == STREAM_GEMV3) return a.alpha*sum + rhs[k];
		else if(KIND == STREAM_TRI_LOWER) return rhs[k] - sum;
		else if(KIND == STREAM_TRI_UPPER) return (1.0/__ldg(a.diag + row))*(rhs[k] - sum);
		else if(KIND == STREAM_SGS_FWD) return __ldg(a.diag + row)*(rhs[k] - sum);   // diag = 1/a_ii
		else return rhs[k] - __ldg(a.diag + row)*sum;                               // STREAM_SGS_BWD
	};

	// rows: gather, multiply, add, single final store per row.  A thread's row groups are taken in
	// sweep direction and each is stored before the next is gathered: how early a new value
	// becomes visible to the other CTAs decides how many sweeps a solve needs (7-point 512^3,
	// FGMRES(30), sweeps (5,3): 2631 iterations this way, 3435 with the stores after both groups).
	constexpr bool FORWARD = (KIND == STREAM_TRI_LOWER || KIND == STREAM_SGS_FWD);
	constexpr bool BACKWARD = (KIND == STREAM_TRI_UPPER || KIND == STREAM_SGS_BWD);
#pragma unroll
	for(int kk = 0; kk < RPT; kk++) {
		const int k = BACKWARD ? RPT - 1 - kk : kk;
		const int lr = (LPR > 1) ? tid/LPR : tid + k*256;
		const int lane = (LPR > 1) ? tid - lr*LPR : 0;
		const int w = tid & 31;
		double sum = 0, chain = 0;
		if(lr < nr) {
			const int s = sptr[lr], e = sptr[lr + 1];
			// CHAIN: the entry that couples the row to its neighbour in sweep direction is kept
			// apart when that neighbour belongs to the adjacent lane of this warp
			const int nb = !CHAIN ? -1 : FORWARD ? (w > 0 ? r0 + lr - 1 : -1)
			                                     : ((w < 31 && lr + 1 < nr) ? r0 + lr + 1 : -1);
#pragma unroll 4
			for(int i = s + lane; i < e; i += LPR) {
				const int c = tc[i];
				double xv = (KIND == STREAM_SPMV || KIND == STREAM_GEMV3) ? __ldg(a.x + c)
				                                                         : __ldcg(a.x + c);
				if(KIND != STREAM_SPMV && KIND != STREAM_GEMV3 && a.xscale) xv *= __ldg(a.xscale + c);
				double v = tv[i];
				if(CHAIN) {                      // branch-free: the chain entry leaves the sum
					const bool is = (c == nb);
					chain = is ? v : chain;
					v = is ? 0.0 : v;
				}
				sum = fma(v, xv, sum);
			}
		}
#pragma unroll
		for(int off = LPR/2; off > 0; off >>= 1)
			sum += __shfl_down_sync(0xffffffffu, sum, off, LPR);
		if(!CHAIN) {
			if(lr < nr && lane == 0) a.out[r0 + lr] = finish(k, r0 + lr, sum);
		}
		else {
			// The 32 consecutive rows of a warp are solved EXACTLY along that chain, as a thread
			// of the reference does for the rows of its chunk (kernels_ilu_apply.hpp:15-42 under
			// schedule(dynamic, chunk): rows in order, the predecessor's new value is used): the
			// new values obey v_i = A_i v_(i±1) + C_i, a first-order recurrence, evaluated by a
			// warp scan over the affine maps (5 steps).  Everything else stays a relaxed read.
			double C = 0, A = 0;
			if(lr < nr) {
				const int row = r0 + lr;
				if(KIND == STREAM_TRI_LOWER) { C = rhs[k] - sum; A = -chain; }
				else if(KIND == STREAM_TRI_UPPER) {
					const double dinv = 1.0/__ldg(a.diag + row);
					C = dinv*(rhs[k] - sum); A = -chain*dinv;
				}
				else {
					const double d = __ldg(a.diag + row);            // 1/a_ii
					C = (KIND == STREAM_SGS_FWD) ? d*(rhs[k] - sum) : rhs[k] - d*sum;
					A = -chain*d;
				}
			}
#pragma unroll
			for(int off = 1; off < 32; off <<= 1) {
				const double Ao = FORWARD ? __shfl_up_sync(0xffffffffu, A, off) : __shfl_down_sync(0xffffffffu, A, off);
				const double Co = FORWARD ? __shfl_up_sync(0xffffffffu, C, off) : __shfl_down_sync(0xffffffffu, C, off);
				if(FORWARD ? (w >= off) : (w + off < 32)) { C = fma(A, Co, C); A *= Ao; }
			}
			if(lr < nr) a.out[r0 + lr] = C;
		}
	}
}

template <int KIND>
static void launch_kind(const StreamArgs& a, int max_len, cudaStream_t st)
{
	const int nrows = a.row_end - a.row_begin;
	if(nrows <= 0) return;
	// tile shapes: rows per CTA R = 256*RPT/LPR with R*max_len <= CAP staged entries.  Small tiles
	// (more resident CTAs, staging and row phases of different CTAs overlap) for the short row
	// parts of the triangular sweeps; the largest tiles for full rows.
	static const bool small = getenv("B200_STREAM_LARGE") == nullptr;
	constexpr bool tri = (KIND != STREAM_SPMV && KIND != STREAM_GEMV3);
#define B200_STREAM_CASE(L, P, C)                                                  \
	{                                                                              \
		constexpr int R = 256*P/L;                                                 \
		bool done = false;                                                         \
		if constexpr(tri && L == 1) {                                              \
			if(a.chain) {                                                          \
				csr_stream_kernel<KIND,L,P,C,true><<<div_up(nrows, R), 256, 0, st>>>(a); \
				done = true;                                                       \
			}                                                                      \
		}                                                                          \
		if(!done) csr_stream_kernel<KIND,L,P,C,false><<<div_up(nrows, R), 256, 0, st>>>(a); \
	}
	if(tri && small && max_len <= 3) B200_STREAM_CASE(1, 2, 1536)
	else if(tri && small && max_len <= 7) B200_STREAM_CASE(1, 1, 1792)
	else if(max_len <= STREAM_CAP/1024) B200_STREAM_CASE(1, 4, STREAM_CAP)
	else if(max_len <= STREAM_CAP/512) B200_STREAM_CASE(1, 2, STREAM_CAP)
	else if(max_len <= STREAM_CAP/256) B200_STREAM_CASE(1, 1, STREAM_CAP)
	else if(max_len <= STREAM_CAP/128) B200_STREAM_CASE(2, 1, STREAM_CAP)
	else if(max_len <= STREAM_CAP/64) B200_STREAM_CASE(4, 1, STREAM_CAP)
	else B200_STREAM_CASE(8, 1, STREAM_CAP)
#undef B200_STREAM_CASE
	B200_LAUNCHED();
}

bool stream_supported(int max_row_len) { return max_row_len > 0 && max_row_len <= STREAM_CAP/32; }

void launch_csr_stream(StreamKind kind, const StreamArgs& a, int max_len, cudaStream_t st)
{
	if(!stream_supported(max_len)) throw Error("csr stream: rows too long for the staged kernel");
	switch(kind) {
	case STREAM_SPMV: launch_kind<STREAM_SPMV>(a, max_len, st); break;
	case STREAM_GEMV3: launch_kind<STREAM_GEMV3>(a, max_len, st); break;
	case STREAM_TRI_LOWER: launch_kind<STREAM_TRI_LOWER>(a, max_len, st); break;
	case STREAM_TRI_UPPER: launch_kind<STREAM_TRI_UPPER>(a, max_len, st); break;
	case STREAM_SGS_FWD: launch_kind<STREAM_SGS_FWD>(a, max_len, st); break;
	case STREAM_SGS_BWD: launch_kind<STREAM_SGS_BWD>(a, max_len, st); break;
	}
}

}  // namespace b200
