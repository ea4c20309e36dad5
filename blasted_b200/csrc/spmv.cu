/** \file spmv.cu
 * \brief CSR / BSR sparse matrix-vector products (K9): y = A x and z = a A x + b y.
 *
 * Replaces BLAS_CSR::matrix_apply / gemv3 and BLAS_BSR::matrix_apply / gemv3
 * (src/blas/matvecs.cpp:25-108 of the reference; `omp parallel for` over rows).
 *
 * HBM-bound: algorithmic bytes  CSR 12 nnz + 4(N+1) + 16 N ; BSR (8 b^2 + 4) nnzb + 4(N+1) + 16 b N
 * (SURVEY.md section 8(d)).  Mapping:
 *  - CSR: a group of LPR lanes (2..32, chosen from the mean row length) per row; consecutive groups
 *    take consecutive rows, so a warp's loads of vals/bcolind cover one contiguous span; partial
 *    sums are combined with warp shuffles.
 *  - BSR: a group of bs lanes per block-row, lane r owns row r of every block (blockops.cuh: for
 *    bs=4 one 256-bit load per block row, so a group touches each 128-byte line exactly once); the
 *    x segment is a broadcast load, no shuffles are needed, the bs results are written contiguously.
 */
#include "common.cuh"
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include "blockops.cuh"

namespace b200 {

template <int LPR, bool G3>
__global__ void __launch_bounds__(256)
csr_spmv_kernel(const int nrows, const int *__restrict__ rowptr, const int *__restrict__ colind,
                const double *__restrict__ vals, const double *__restrict__ x,
                const double a, const double b, const double *yin, double *z)
{
	const long long tid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	const int row = (int)(tid / LPR);
	const int lane = (int)(tid % LPR);
	double sum = 0;
	if(row < nrows) {
		const int s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
		for(int j = s + lane; j < e; j += LPR)
			sum += __ldg(vals + j) * __ldg(x + __ldg(colind + j));
	}
#pragma unroll
	for(int off = LPR/2; off > 0; off >>= 1)
		sum += __shfl_down_sync(0xffffffffu, sum, off, LPR);
	if(lane == 0 && row < nrows) {
		if(G3) z[row] = a*sum + b*yin[row];
		else z[row] = sum;
	}
}

template <int BS, bool G3, bool VEC>
__global__ void __launch_bounds__(256)
bsr_spmv_kernel(const int nbrows, const int *__restrict__ browptr, const int *__restrict__ bcolind,
                const double *__restrict__ vals, const double *__restrict__ x,
                const double a, const double b, const double *yin, double *z)
{
	constexpr int GPW = 32/BS;                     // block-rows per warp
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const int g = lane / BS, r = lane - g*BS;
	const long long rowl = warp*GPW + g;
	if(g >= GPW || rowl >= nbrows) return;
	const int row = (int)rowl;

	const int s = __ldg(browptr + row), e = __ldg(browptr + row + 1);
	double acc = 0;
#pragma unroll 2
	for(int jj = s; jj < e; jj++) {
		const int col = __ldg(bcolind + jj);
		double av[BS], xv[BS];
		BlkIO<BS>::template load_row<false>(vals + (size_t)jj*(BS*BS), r, av);
		load_seg<BS,false,VEC>(x + (size_t)col*BS, xv);
#pragma unroll
		for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
	}
	const size_t o = (size_t)row*BS + r;
	if(G3) z[o] = a*acc + b*yin[o];
	else z[o] = acc;
}

/// Persistent, software-pipelined form for bs = 4 (the same idea as apply.cu::tri_block_pipe_kernel):
/// with five blocks per row the chain  row pointers -> column indices -> x segments  bounds the
/// one-shot kernel, so every group keeps the row pointers of its row after next and the first PK
/// column indices of its next row in registers and requests all loads of the current row at once.
/// Accumulation stays in column order (deterministic, identical to the one-shot kernel).
template <bool G3>
__global__ void __launch_bounds__(256, 4)
bsr4_spmv_pipe_kernel(const int nbrows, const int *__restrict__ browptr,
                      const int *__restrict__ bcolind, const double *__restrict__ vals,
                      const double *__restrict__ x, const double a, const double b, const double *yin,
                      double *z)
{
	constexpr int BS = 4, GPW = 8, BS2 = 16, PK = 6;
	const int lane = threadIdx.x & 31;
	const int g = lane >> 2, r = lane & 3;
	const int wpc = blockDim.x >> 5;
	const long long stride = (long long)gridDim.x*wpc*GPW;
	const long long wbase = ((long long)blockIdx.x*wpc + (threadIdx.x >> 5))*GPW;

	auto load_meta = [&](const long long t, int& js, int& je) {
		js = 0; je = 0;
		if(t < nbrows) { js = __ldg(browptr + t); je = __ldg(browptr + t + 1); }
	};
	auto load_cols = [&](const int js, const int je, int (&c)[PK]) {
#pragma unroll
		for(int q = 0; q < PK; q++) c[q] = (js + q < je) ? __ldg(bcolind + js + q) : 0;
	};
	int js1, je1, c1[PK], js2, je2;
	load_meta(wbase + g, js1, je1);
	load_cols(js1, je1, c1);
	load_meta(wbase + g + stride, js2, je2);

	for(long long tw = wbase; tw < nbrows; tw += stride) {
		const long long t = tw + g;
		const int js = js1, je = je1;
		int cols[PK];
#pragma unroll
		for(int q = 0; q < PK; q++) cols[q] = c1[q];
		js1 = js2; je1 = je2;
		load_cols(js1, je1, c1);
		load_meta(t + 2*stride, js2, je2);
		if(t >= nbrows) continue;
		double acc = 0;
#pragma unroll
		for(int q = 0; q < PK; q++) {
			const int jj = js + q;
			if(jj < je) {
				double av[BS], xv[BS];
				BlkIO<BS>::template load_row<false>(vals + (size_t)jj*BS2, r, av);
				load_seg<BS,false,true>(x + (size_t)cols[q]*BS, xv);
#pragma unroll
				for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
			}
		}
		for(int jj = js + PK; jj < je; jj++) {
			const int col = __ldg(bcolind + jj);
			double av[BS], xv[BS];
			BlkIO<BS>::template load_row<false>(vals + (size_t)jj*BS2, r, av);
			load_seg<BS,false,true>(x + (size_t)col*BS, xv);
#pragma unroll
			for(int c = 0; c < BS; c++) acc = fma(av[c], xv[c], acc);
		}
		const size_t o = (size_t)t*BS + r;
		if(G3) z[o] = a*acc + b*yin[o];
		else z[o] = acc;
	}
}

template <bool G3>
static void launch_csr(const Mat& A, double a, const double *x, double b, const double *y, double *z,
                       cudaStream_t st)
{
	const int n = A.nbrows;
	const double avg = A.avg_row_len;
	if(stream_supported(A.max_row_len)) {
		// short rows: staged unit-stride stream (csrstream.cu)
		StreamArgs sa;
		sa.ptr = A.browptr; sa.col = A.bcolind; sa.val = A.vals; sa.x = x; sa.out = z; sa.yin = y;
		sa.alpha = a; sa.beta = b; sa.row_end = n;
		launch_csr_stream(G3 ? STREAM_GEMV3 : STREAM_SPMV, sa, A.max_row_len, st);
		return;
	}
#define B200_CSR_CASE(L)                                                                           \
	{                                                                                              \
		const int grid = div_up((long long)n*L, 256);                                              \
		csr_spmv_kernel<L,G3><<<grid, 256, 0, st>>>(n, A.browptr, A.bcolind, A.vals, x, a, b, y, z); \
	}
	if(avg <= 2.5) B200_CSR_CASE(2)
	else if(avg <= 5) B200_CSR_CASE(4)
	else if(avg <= 12) B200_CSR_CASE(8)
	else if(avg <= 24) B200_CSR_CASE(16)
	else B200_CSR_CASE(32)
#undef B200_CSR_CASE
	B200_LAUNCHED();
}

template <int BS, bool G3>
static void launch_bsr(const Mat& A, double a, const double *x, double b, const double *y, double *z,
                       cudaStream_t st)
{
	constexpr int GPW = 32/BS;
	const long long nwarps = ((long long)A.nbrows + GPW - 1)/GPW;
	const int grid = div_up(nwarps*32, 256);
	static const bool one_shot = getenv("B200_SPMV1") != nullptr;        // A/B switch (development)
	if(BS == 4 && aligned32(x) && !one_shot && A.nbrows >= 4096) {
		static int resident = 0;
		if(!resident) {
			int dev = 0, sms = 148, per = 4;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, bsr4_spmv_pipe_kernel<G3>, 256, 0) != cudaSuccess || per < 1)
				per = 4;
			resident = sms*per;
		}
		const int pg = (int)std::min<long long>(resident, (A.nbrows + 63)/64);
		bsr4_spmv_pipe_kernel<G3><<<pg, 256, 0, st>>>(A.nbrows, A.browptr, A.bcolind, A.vals, x, a, b, y, z);
	}
	else if(BS == 4 && aligned32(x))
		bsr_spmv_kernel<BS,G3,true><<<grid, 256, 0, st>>>(A.nbrows, A.browptr, A.bcolind, A.vals, x, a, b, y, z);
	else
		bsr_spmv_kernel<BS,G3,false><<<grid, 256, 0, st>>>(A.nbrows, A.browptr, A.bcolind, A.vals, x, a, b, y, z);
	B200_LAUNCHED();
}

template <bool G3>
static void dispatch(const Mat& A, double a, const double *x, double b, const double *y, double *z,
                     cudaStream_t st)
{
	if(A.nbrows == 0) return;
	ProfScope ps(KC_SPMV, st);
	switch(A.bs) {
	case 1: launch_csr<G3>(A, a, x, b, y, z, st); break;
	case 3: launch_bsr<3,G3>(A, a, x, b, y, z, st); break;
	case 4: launch_bsr<4,G3>(A, a, x, b, y, z, st); break;
	case 5: launch_bsr<5,G3>(A, a, x, b, y, z, st); break;
	case 7: launch_bsr<7,G3>(A, a, x, b, y, z, st); break;
	default: throw Error("SpMV: block size " + std::to_string(A.bs) + " not supported");
	}
}

void launch_spmv(const Mat& A, const double *x, double *y, cudaStream_t st)
{
	dispatch<false>(A, 1.0, x, 0.0, nullptr, y, st);
}

void launch_gemv3(const Mat& A, double a, const double *x, double b, const double *y, double *z,
                  cudaStream_t st)
{
	dispatch<true>(A, a, x, b, y, z, st);
}

// z(rows) += a (A x)(rows) over a LIST of block rows: the coupling part of a partitioned operator
// has entries in the subdomain's boundary rows only (2 of n planes of a z-slab), so the product
// touches those rows instead of streaming the row pointers and z of the whole subdomain.

template <int BS>
__global__ void __launch_bounds__(256)
rows_gemv_add_kernel(const int nlist, const int *__restrict__ rows, const int *__restrict__ browptr,
                     const int *__restrict__ bcolind, const double *__restrict__ vals,
                     const double *__restrict__ x, const double a, double *z)
{
	const long long tid = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	const long long idx = tid / BS;
	const int r = (int)(tid - idx*BS);
	if(idx >= nlist) return;
	const int row = __ldg(rows + idx);
	const int s = __ldg(browptr + row), e = __ldg(browptr + row + 1);
	double acc = 0;
	for(int jj = s; jj < e; jj++) {
		const double *xc = x + (size_t)__ldg(bcolind + jj)*BS;
		if(BS == 1) acc = fma(__ldg(vals + jj), __ldg(xc), acc);
		else {
			double av[BS];
			BlkIO<BS>::template load_row<false>(vals + (size_t)jj*(BS*BS), r, av);
#pragma unroll
			for(int c = 0; c < BS; c++) acc = fma(av[c], __ldg(xc + c), acc);
		}
	}
	z[(size_t)row*BS + r] += a*acc;
}

void launch_gemv_add_rows(const Mat& A, int nlist, const int *d_rows, double a, const double *x,
                          double *z, cudaStream_t st)
{
	if(nlist == 0) return;
	ProfScope ps(KC_SPMV, st);
#define B200_ROWS_CASE(B)                                                                          \
	case B: rows_gemv_add_kernel<B><<<div_up((long long)nlist*B, 256), 256, 0, st>>>(              \
	            nlist, d_rows, A.browptr, A.bcolind, A.vals, x, a, z); break;
	switch(A.bs) {
	B200_ROWS_CASE(1) B200_ROWS_CASE(3) B200_ROWS_CASE(4) B200_ROWS_CASE(5) B200_ROWS_CASE(7)
	default: throw Error("SpMV: block size " + std::to_string(A.bs) + " not supported");
	}
#undef B200_ROWS_CASE
	B200_LAUNCHED();
}

__global__ void __launch_bounds__(256)
nonempty_flags_kernel(const int n, const int *__restrict__ browptr, char *__restrict__ flag)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) flag[i] = browptr[i+1] > browptr[i];
}

int nonempty_rows(const Mat& A, DevBuf<int>& rows, cudaStream_t st)
{
	const int n = A.nbrows;
	if(n == 0 || A.nnzb == 0) return 0;
	DevBuf<char> flag, tmp;
	DevBuf<int> d_cnt, all;
	flag.alloc(n); d_cnt.alloc(1); all.alloc(n);
	nonempty_flags_kernel<<<div_up(n, 256), 256, 0, st>>>(n, A.browptr, flag);
	B200_LAUNCHED();
	size_t tb = 0;
	thrust::counting_iterator<int> ids(0);
	cub::DeviceSelect::Flagged(nullptr, tb, ids, flag.p, all.p, d_cnt.p, n, st);
	tmp.alloc(tb);
	B200_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, ids, flag.p, all.p, d_cnt.p, n, st));
	int cnt = 0;
	B200_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	rows.alloc(std::max(cnt, 1));
	B200_CUDA(cudaMemcpyAsync(rows, all, (size_t)cnt*sizeof(int), cudaMemcpyDeviceToDevice, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return cnt;
}

}  // namespace b200
