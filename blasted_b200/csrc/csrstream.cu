/** \file csrstream.cu
 * \brief Scalar (bs=1) row sweeps as a staged stream: SpMV, gemv3 and the asynchronous triangular
 * sweeps over the split L / U parts of the scalar ILU(0) factor.
 *
 * Replaces the same reference loops as spmv.cu / apply.cu (src/blas/matvecs.cpp:78-108,
 * src/kernels/kernels_ilu_apply.hpp:15-42, src/solverops_ilu0.cpp:274-314) for short scalar rows,
 * where a lanes-per-row mapping leaves too few bytes in flight per warp to cover HBM latency.
 *
 * One CTA owns a tile of R consecutive rows (R*max_row_len <= CAP entries).  Phase 1: all 256
 * threads walk the tile's CONTIGUOUS span of (value, column) pairs with unit stride - every load
 * is fully coalesced and independent of the others, so many are in flight per thread - multiply
 * by the gathered vector entry and park the products in shared memory.  Phase 2: LPR = 256/R
 * lanes per row add up that row's products from shared memory and write the single final value.
 * Tiles are issued in ascending row order for lower/forward sweeps and descending order for
 * upper/backward sweeps (same propagation argument as apply.cu).
 *
 * HBM-bound: 12 B per stored entry + 4 B row pointer + vectors.
 */
#include "common.cuh"

namespace b200 {

constexpr int STREAM_CAP = 4096;       // products staged per CTA (32 KiB of shared memory)

template <int KIND, int LPR>
__global__ void __launch_bounds__(256)
csr_stream_kernel(const StreamArgs a)
{
	constexpr int R = 256/LPR;
	__shared__ double prod[STREAM_CAP];
	__shared__ int sptr[R + 1];

	const int ntiles = gridDim.x;
	const int tile = a.descending ? ntiles - 1 - (int)blockIdx.x : (int)blockIdx.x;
	const int r0 = a.row_begin + tile*R;
	const int nr = min(R, a.row_end - r0);
	const int tid = threadIdx.x;

	if(tid < nr) sptr[tid] = __ldg(a.ptr + r0 + tid);
	if(tid == 0) sptr[nr] = __ldg(a.ptr + r0 + nr);
	__syncthreads();
	const int e0 = sptr[0];
	const int ne = sptr[nr] - e0;

	// phase 1: unit-stride walk over the tile's entries
	const int *__restrict__ col = a.col + e0;
	const double *__restrict__ val = a.val + e0;
#pragma unroll 4
	for(int i = tid; i < ne; i += 256) {
		const int c = __ldg(col + i);
		const double v = __ldg(val + i);
		const double xv = (KIND == STREAM_SPMV || KIND == STREAM_GEMV3) ? __ldg(a.x + c)
		                                                               : __ldcg(a.x + c);
		prod[i] = v*xv;
	}
	__syncthreads();

	// phase 2: LPR lanes per row
	const int lr = tid / LPR, lane = tid - lr*LPR;
	double sum = 0;
	if(lr < nr) {
		const int s = sptr[lr] - e0, e = sptr[lr + 1] - e0;
		for(int i = s + lane; i < e; i += LPR) sum += prod[i];
	}
#pragma unroll
	for(int off = LPR/2; off > 0; off >>= 1)
		sum += __shfl_down_sync(0xffffffffu, sum, off, LPR);
	if(lr < nr && lane == 0) {
		const int row = r0 + lr;
		if(KIND == STREAM_SPMV) a.out[row] = sum;
		else if(KIND == STREAM_GEMV3) a.out[row] = a.alpha*sum + a.beta*a.yin[row];
		else {
			double rhs = __ldg(a.rhs + row);
			if(a.rscale) rhs *= __ldg(a.rscale + row);
			if(KIND == STREAM_TRI_LOWER) a.out[row] = rhs - sum;
			else a.out[row] = (1.0/__ldg(a.diag + row))*(rhs - sum);     // STREAM_TRI_UPPER
		}
	}
}

template <int KIND>
static void launch_kind(const StreamArgs& a, int max_len, cudaStream_t st)
{
	const int nrows = a.row_end - a.row_begin;
	if(nrows <= 0) return;
#define B200_STREAM_CASE(L)                                                        \
	{                                                                              \
		constexpr int R = 256/L;                                                   \
		csr_stream_kernel<KIND,L><<<div_up(nrows, R), 256, 0, st>>>(a);            \
	}
	if(max_len <= STREAM_CAP/256) B200_STREAM_CASE(1)
	else if(max_len <= STREAM_CAP/128) B200_STREAM_CASE(2)
	else if(max_len <= STREAM_CAP/64) B200_STREAM_CASE(4)
	else B200_STREAM_CASE(8)
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
	}
}

}  // namespace b200
