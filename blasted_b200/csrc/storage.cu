/** \file storage.cu
 * \brief Device storage helpers: block transposition at the boundary, entry->row map, diagonal
 * location.  Stands in for include/device_container.hpp, srmatrixdefs.hpp, rawsrmatrixutils.cpp of
 * the reference (host, 64-byte aligned vectors) with arrays resident in HBM.
 */
#include "common.cuh"

namespace b200 {

/// out block (r,c) <- in block (c,r): converts row-major <-> column-major blocks
__global__ void transpose_blocks_kernel(int bs, long long n, const double *__restrict__ in,
                                        double *__restrict__ out)
{
	const int bs2 = bs*bs;
	const long long total = n*bs2;
	for(long long e = blockIdx.x*(long long)blockDim.x + threadIdx.x; e < total;
	    e += (long long)gridDim.x*blockDim.x)
	{
		const long long blk = e / bs2;
		const int w = (int)(e - blk*bs2);
		const int r = w % bs, c = w / bs;
		out[e] = in[blk*bs2 + r*bs + c];
	}
}

void transpose_blocks(int bs, long long nblocks, const double *in, double *out, cudaStream_t st)
{
	if(nblocks == 0) return;
	const long long total = nblocks*bs*bs;
	const int grid = (int)std::min<long long>(div_up(total, 256), 148*16);
	transpose_blocks_kernel<<<grid, 256, 0, st>>>(bs, nblocks, in, out);
	B200_LAUNCHED();
}

__global__ void browind_kernel(int nbrows, const int *__restrict__ browptr, int *__restrict__ browind)
{
	// one warp per row keeps the writes coalesced for long rows and cheap for short ones
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int nwarps = (gridDim.x*blockDim.x) >> 5;
	for(int row = warp; row < nbrows; row += nwarps) {
		const int s = browptr[row], e = browptr[row+1];
		for(int j = s + lane; j < e; j += 32) browind[j] = row;
	}
}

void build_browind(const Mat& A, cudaStream_t st)
{
	if(A.nbrows == 0) return;
	const int grid = std::min(div_up((long long)A.nbrows*32, 256), 148*32);
	browind_kernel<<<grid, 256, 0, st>>>(A.nbrows, A.browptr, A.browind);
	B200_LAUNCHED();
}

__global__ void find_diag_kernel(int nbrows, const int *__restrict__ browptr,
                                 const int *__restrict__ bcolind, int *__restrict__ diagind,
                                 int *__restrict__ nmissing)
{
	for(int row = blockIdx.x*blockDim.x + threadIdx.x; row < nbrows; row += gridDim.x*blockDim.x) {
		int lo = browptr[row], hi = browptr[row+1] - 1, pos = -1;
		while(lo <= hi) {                       // columns are sorted (levelschedule.hpp:15)
			const int mid = (lo + hi) >> 1;
			const int c = bcolind[mid];
			if(c == row) { pos = mid; break; }
			if(c < row) lo = mid + 1; else hi = mid - 1;
		}
		diagind[row] = pos;
		if(pos < 0) atomicAdd(nmissing, 1);
	}
}

int find_diagonals(Mat& A, cudaStream_t st)
{
	if(A.nbrows == 0) return 0;
	DevBuf<int> d_missing;
	d_missing.alloc(1);
	B200_CUDA(cudaMemsetAsync(d_missing, 0, sizeof(int), st));
	const int grid = std::min(div_up(A.nbrows, 256), 148*16);
	find_diag_kernel<<<grid, 256, 0, st>>>(A.nbrows, A.browptr, A.bcolind, A.diagind, d_missing);
	B200_LAUNCHED();
	int missing = 0;
	B200_CUDA(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return missing;
}

__global__ void max_row_len_kernel(int nbrows, const int *__restrict__ browptr, int *__restrict__ out)
{
	int m = 0;
	for(int row = blockIdx.x*blockDim.x + threadIdx.x; row < nbrows; row += gridDim.x*blockDim.x)
		m = max(m, browptr[row+1] - browptr[row]);
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, off));
	if((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

int max_row_length(const Mat& A, cudaStream_t st)
{
	if(A.nbrows == 0) return 0;
	DevBuf<int> d;
	d.alloc(1);
	B200_CUDA(cudaMemsetAsync(d, 0, sizeof(int), st));
	max_row_len_kernel<<<std::min(div_up(A.nbrows, 256), 148*8), 256, 0, st>>>(A.nbrows, A.browptr, d);
	B200_LAUNCHED();
	int m = 0;
	B200_CUDA(cudaMemcpyAsync(&m, d, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return m;
}

}  // namespace b200
