/** \file frontend.cu
 * \brief The step before the hot path, on the device: coordinate triplets -> CSR/BSR, and
 * permutation / scaling of a matrix that is already resident.
 *
 * Replaces, from the reference:
 *   COOMatrix::readMatrixMarket's sorting (row, then column)          src/coomatrix.cpp:222-260
 *   COOMatrix::convertToCSR / convertToBSR<bs,stor>                   src/coomatrix.cpp:262-403
 *   Reordering::applyOrdering (matrix and vector, forward / inverse)  src/reorderingscaling.cpp:77-266
 *   ReorderingScaling::applyScaling (matrix and vector)               src/reorderingscaling.cpp:282-368
 *
 * The reference does all of this with serial loops, std::sort per row and linear std::find searches
 * per entry.  Here both conversions and both permutations are ONE stable radix sort of 64-bit keys
 * (block-row << 32 | block-column) followed by gathers:
 *   triplets -> keys -> sort -> head flags -> scan -> {bcolind, browptr, scatter of values};
 *   reorder: key = (new row of old row << 32 | new column of old column) over the stored entries.
 * Integer outputs are identical to the reference's for matrices whose scalar rows share their block
 * row's block pattern (every fixture and every assembled block matrix); for ragged block rows the
 * reference appends block columns in order of first appearance (src/coomatrix.cpp:330-348), which
 * breaks its own sorted-columns assumption (levelschedule.hpp:15) - here they are always sorted.
 */
#include "common.cuh"
#include "blockops.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace b200 {

namespace {

typedef unsigned long long u64;

__global__ void coo_keys_kernel(const long long nnz, const int bs, const int *__restrict__ row,
                                const int *__restrict__ col, u64 *__restrict__ key,
                                int *__restrict__ idx, int *__restrict__ bad, const int nrows)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= nnz) return;
	const int r = row[i], c = col[i];
	if(r < 0 || c < 0 || r >= nrows || c >= nrows) { atomicAdd(bad, 1); key[i] = ~0ull; idx[i] = (int)i; return; }
	key[i] = ((u64)(unsigned)(r/bs) << 32) | (unsigned)(c/bs);
	idx[i] = (int)i;
}

/// head[i] = 1 where sorted entry i starts a new stored (block) entry; bs == 1 keeps duplicates
/// as separate entries, as convertToCSR does
__global__ void heads_kernel(const long long nnz, const int bs, const u64 *__restrict__ key,
                             int *__restrict__ head)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i > nnz) return;
	if(i == nnz) { head[i] = 0; return; }
	head[i] = (bs == 1 || i == 0 || key[i] != key[i-1]) ? 1 : 0;
}

/// per sorted entry: write its block's column index (heads only) and fill browptr for every block
/// row that begins at or before this block and after the previous block's row
__global__ void block_pattern_kernel(const long long nnz, const int nbrows, const long long nnzb,
                                     const u64 *__restrict__ key, const int *__restrict__ head,
                                     const int *__restrict__ blockid, int *__restrict__ bcolind,
                                     int *__restrict__ browptr)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= nnz) {
		if(i == nnz) {
			// rows after the last stored one
			const int last = nnz ? (int)(key[nnz-1] >> 32) : -1;
			for(int r = last + 1; r <= nbrows; r++) browptr[r] = (int)nnzb;
		}
		return;
	}
	if(!head[i]) return;
	const int b = blockid[i];
	const int brow = (int)(key[i] >> 32);
	bcolind[b] = (int)(key[i] & 0xffffffffu);
	const int prev = i ? (int)(key[i-1] >> 32) : -1;
	for(int r = prev + 1; r <= brow; r++) browptr[r] = b;
}

template <int BS>
__global__ void scatter_values_kernel(const long long nnz, const int *__restrict__ idx,
                                      const int *__restrict__ head,
                                      const int *__restrict__ blockid, const int *__restrict__ row,
                                      const int *__restrict__ col, const double *__restrict__ val,
                                      double *__restrict__ vals)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= nnz) return;
	const int src = idx[i];
	const int r = row[src] % BS, c = col[src] % BS;
	// blockid is the exclusive scan of the head flags: the entry's own block for a head, one past
	// it for the entries that follow in the same block
	const int b = blockid[i] - (head[i] ? 0 : 1);
	vals[(size_t)b*BS*BS + BlkIO<BS>::at(r, c)] = val[src];
}

__global__ void reorder_keys_kernel(const long long nnzb, const int *__restrict__ browind,
                                    const int *__restrict__ bcolind, const int *__restrict__ rowmap,
                                    const int *__restrict__ colmap, u64 *__restrict__ key,
                                    int *__restrict__ idx)
{
	const long long j = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(j >= nnzb) return;
	const int r = browind[j], c = bcolind[j];
	const int nr = rowmap ? rowmap[r] : r, nc = colmap ? colmap[c] : c;
	key[j] = ((u64)(unsigned)nr << 32) | (unsigned)nc;
	idx[j] = (int)j;
}

__global__ void new_row_len_kernel(const int nbrows, const int *__restrict__ browptr,
                                   const int *__restrict__ rowmap, int *__restrict__ len)
{
	const int r = blockIdx.x*blockDim.x + threadIdx.x;
	if(r > nbrows) return;
	if(r == nbrows) { len[nbrows] = 0; return; }
	len[rowmap ? rowmap[r] : r] = browptr[r+1] - browptr[r];
}

__global__ void gather_blocks_kernel(const long long nnzb, const int bs2, const int *__restrict__ idx,
                                     const u64 *__restrict__ key, const double *__restrict__ in,
                                     double *__restrict__ out, int *__restrict__ bcolind)
{
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= nnzb*bs2) return;
	const long long j = e / bs2;
	const int w = (int)(e - j*bs2);
	out[e] = in[(size_t)idx[j]*bs2 + w];
	if(w == 0) bcolind[j] = (int)(key[j] & 0xffffffffu);
}

__global__ void invert_perm_kernel(const int n, const int *__restrict__ p, int *__restrict__ inv,
                                   int *__restrict__ bad)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	const int t = p[i];
	if(t < 0 || t >= n) { atomicAdd(bad, 1); return; }
	inv[t] = i;
}

/// MODE 0: v *= s ; MODE 1: v /= s   (true divisions, as the reference: results are bit-identical)
template <int MODE>
__global__ void scale_matrix_kernel(const long long nnzb, const int bs2, const int *__restrict__ browind,
                                    const int *__restrict__ bcolind, const double *__restrict__ rowscale,
                                    const double *__restrict__ colscale, double *__restrict__ vals)
{
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= nnzb*bs2) return;
	const long long j = e / bs2;
	double v = vals[e];
	if(rowscale) { const double s = rowscale[browind[j]]; v = MODE ? v / s : v * s; }
	if(colscale) { const double s = colscale[bcolind[j]]; v = MODE ? v / s : v * s; }
	vals[e] = v;
}

template <int MODE>
__global__ void scale_vector_kernel(const long long n, const int bs, const double *__restrict__ scale,
                                    double *__restrict__ vec)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n*bs) return;
	const double s = scale[i / bs];
	vec[i] = MODE ? vec[i] / s : vec[i] * s;
}

/// forward: out[i] = in[ord[i]]; inverse: out[ord[i]] = in[i]   (block entries of bs doubles)
template <int MODE>
__global__ void permute_vector_kernel(const long long n, const int bs, const int *__restrict__ ord,
                                      const double *__restrict__ in, double *__restrict__ out)
{
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= n*bs) return;
	const long long i = e / bs;
	const int k = (int)(e - i*bs);
	const long long o = ord[i];
	if(MODE == 0) out[e] = in[o*bs + k];
	else out[o*bs + k] = in[e];
}

void sort_pairs(const long long n, DevBuf<u64>& key, DevBuf<u64>& key_out, DevBuf<int>& idx,
                DevBuf<int>& idx_out, const int end_bit, cudaStream_t st)
{
	if(n == 0) return;
	if(n > 0x7fffffffLL) throw Error("front end: more than 2^31-1 entries");
	size_t tb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key_out.p, idx.p, idx_out.p, (int)n, 0, end_bit, st);
	DevBuf<char> tmp;
	tmp.alloc(std::max<size_t>(tb, 16));
	B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key_out.p, idx.p, idx_out.p, (int)n,
	                                          0, end_bit, st));
	B200_LAUNCHED();
	B200_CUDA(cudaStreamSynchronize(st));        // tmp is released on return
}

void exclusive_scan(const long long n, const int *in, int *out, cudaStream_t st)
{
	size_t tb = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)n, st);
	DevBuf<char> tmp;
	tmp.alloc(std::max<size_t>(tb, 16));
	B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, (int)n, st));
	B200_LAUNCHED();
	B200_CUDA(cudaStreamSynchronize(st));
}

int read_flag(const int *d, cudaStream_t st)
{
	int h = 0;
	B200_CUDA(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	return h;
}

}  // namespace

// ------------------------------------------------------------------ coordinate -> CSR / BSR

void coo_to_mat(Mat& A, const int nrows, const long long nnz, const int *d_row, const int *d_col,
                const double *d_val, cudaStream_t st)
{
	const int bs = A.bs;
	if(nrows < 0 || nnz < 0) throw Error("coordinate matrix: negative size");
	if(nrows % bs != 0) throw Error("coordinate matrix: dimension is not a multiple of the block size");
	A.nbrows = nrows/bs;
	A.browptr.alloc((size_t)A.nbrows + 1);

	DevBuf<u64> key, key_s;
	DevBuf<int> idx, idx_s, head, blockid, bad;
	const size_t cap = (size_t)std::max<long long>(nnz, 1);
	key.alloc(cap); key_s.alloc(cap); idx.alloc(cap); idx_s.alloc(cap);
	head.alloc(cap + 1); blockid.alloc(cap + 1); bad.alloc(1);
	B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
	if(nnz) {
		coo_keys_kernel<<<div_up(nnz, 256), 256, 0, st>>>(nnz, bs, d_row, d_col, key, idx, bad, nrows);
		B200_LAUNCHED();
	}
	if(read_flag(bad, st)) throw Error("coordinate matrix: index out of range");
	sort_pairs(nnz, key, key_s, idx, idx_s, 64, st);
	heads_kernel<<<div_up(nnz + 1, 256), 256, 0, st>>>(nnz, bs, key_s, head);
	B200_LAUNCHED();
	exclusive_scan(nnz + 1, head, blockid, st);
	const long long nnzb = read_flag(blockid.p + nnz, st);
	A.nnzb = nnzb;
	A.bcolind.alloc((size_t)std::max<long long>(nnzb, 1));
	A.vals.alloc(std::max<size_t>((size_t)nnzb*bs*bs, 1));
	block_pattern_kernel<<<div_up(nnz + 1, 256), 256, 0, st>>>(nnz, A.nbrows, nnzb, key_s, head, blockid,
	                                                          A.bcolind, A.browptr);
	B200_LAUNCHED();
	if(nnzb) {
		B200_CUDA(cudaMemsetAsync(A.vals, 0, (size_t)nnzb*bs*bs*sizeof(double), st));
		const int grid = div_up(nnz, 256);
		switch(bs) {
#define B200_SCATTER(B) case B: scatter_values_kernel<B><<<grid, 256, 0, st>>>(nnz, idx_s, head, blockid, d_row, d_col, d_val, A.vals); break;
		B200_SCATTER(1) B200_SCATTER(3) B200_SCATTER(4) B200_SCATTER(5) B200_SCATTER(7)
#undef B200_SCATTER
		default: throw Error("coordinate matrix: unsupported block size " + std::to_string(bs));
		}
		B200_LAUNCHED();
	}
	finish_matrix(A, nullptr, st);
	B200_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------ reordering

void mat_reorder(Mat& A, const int *d_rord, const int *d_cord, const bool inverse, cudaStream_t st)
{
	if(!d_rord && !d_cord) return;
	const int n = A.nbrows;
	const long long nnzb = A.nnzb;
	if(n == 0) return;
	// forward: new row i is old row rord[i], i.e. old row r moves to irp[r]; columns are renamed
	// with the inverse column permutation (reorderingscaling.cpp:83-140).  Inverse: old row i
	// moves to rord[i], columns are renamed with cord itself (:141-205).
	DevBuf<int> irp, icp, bad;
	bad.alloc(1);
	B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
	const int *rowmap = d_rord, *colmap = d_cord;
	if(!inverse) {
		if(d_rord) { irp.alloc(n); invert_perm_kernel<<<div_up(n,256),256,0,st>>>(n, d_rord, irp, bad); rowmap = irp; B200_LAUNCHED(); }
		if(d_cord) { icp.alloc(n); invert_perm_kernel<<<div_up(n,256),256,0,st>>>(n, d_cord, icp, bad); colmap = icp; B200_LAUNCHED(); }
	} else {
		// validate the ranges through the same kernel
		DevBuf<int> scratch;
		scratch.alloc(n);
		if(d_rord) { invert_perm_kernel<<<div_up(n,256),256,0,st>>>(n, d_rord, scratch, bad); B200_LAUNCHED(); }
		if(d_cord) { invert_perm_kernel<<<div_up(n,256),256,0,st>>>(n, d_cord, scratch, bad); B200_LAUNCHED(); }
		if(read_flag(bad, st)) throw Error("reordering: permutation entry out of range");
	}
	if(read_flag(bad, st)) throw Error("reordering: permutation entry out of range");

	DevBuf<int> len, newptr;
	len.alloc((size_t)n + 1); newptr.alloc((size_t)n + 1);
	new_row_len_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, A.browptr, rowmap, len);
	B200_LAUNCHED();
	exclusive_scan(n + 1, len, newptr, st);

	if(nnzb) {
		DevBuf<u64> key, key_s;
		DevBuf<int> idx, idx_s;
		key.alloc(nnzb); key_s.alloc(nnzb); idx.alloc(nnzb); idx_s.alloc(nnzb);
		reorder_keys_kernel<<<div_up(nnzb, 256), 256, 0, st>>>(nnzb, A.browind, A.bcolind, rowmap, colmap, key, idx);
		B200_LAUNCHED();
		sort_pairs(nnzb, key, key_s, idx, idx_s, 64, st);
		DevBuf<double> nv;
		const int bs2 = A.bs*A.bs;
		nv.alloc((size_t)nnzb*bs2);
		gather_blocks_kernel<<<div_up(nnzb*bs2, 256), 256, 0, st>>>(nnzb, bs2, idx_s, key_s, A.vals, nv, A.bcolind);
		B200_LAUNCHED();
		B200_CUDA(cudaMemcpyAsync(A.vals, nv.p, (size_t)nnzb*bs2*sizeof(double), cudaMemcpyDeviceToDevice, st));
		B200_CUDA(cudaStreamSynchronize(st));
	}
	B200_CUDA(cudaMemcpyAsync(A.browptr, newptr.p, ((size_t)n + 1)*sizeof(int), cudaMemcpyDeviceToDevice, st));
	finish_matrix(A, nullptr, st);
	B200_CUDA(cudaStreamSynchronize(st));
}

void vec_reorder(const long long n, const int bs, const int *d_ord, const bool inverse, double *d_vec,
                 cudaStream_t st)
{
	if(n == 0 || !d_ord) return;
	DevBuf<double> tv;
	tv.alloc((size_t)n*bs);
	B200_CUDA(cudaMemcpyAsync(tv.p, d_vec, (size_t)n*bs*sizeof(double), cudaMemcpyDeviceToDevice, st));
	const int grid = div_up(n*bs, 256);
	if(!inverse) permute_vector_kernel<0><<<grid, 256, 0, st>>>(n, bs, d_ord, tv, d_vec);
	else permute_vector_kernel<1><<<grid, 256, 0, st>>>(n, bs, d_ord, tv, d_vec);
	B200_LAUNCHED();
	B200_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------ scaling

void mat_scale(Mat& A, const double *d_rowscale, const double *d_colscale, const bool inverse,
               cudaStream_t st)
{
	if((!d_rowscale && !d_colscale) || A.nnzb == 0) return;
	const int bs2 = A.bs*A.bs;
	const int grid = div_up(A.nnzb*bs2, 256);
	if(!inverse) scale_matrix_kernel<0><<<grid, 256, 0, st>>>(A.nnzb, bs2, A.browind, A.bcolind, d_rowscale, d_colscale, A.vals);
	else scale_matrix_kernel<1><<<grid, 256, 0, st>>>(A.nnzb, bs2, A.browind, A.bcolind, d_rowscale, d_colscale, A.vals);
	B200_LAUNCHED();
}

void vec_scale(const long long n, const int bs, const double *d_scale, const bool inverse, double *d_vec,
               cudaStream_t st)
{
	if(n == 0 || !d_scale) return;
	const int grid = div_up(n*bs, 256);
	if(!inverse) scale_vector_kernel<0><<<grid, 256, 0, st>>>(n, bs, d_scale, d_vec);
	else scale_vector_kernel<1><<<grid, 256, 0, st>>>(n, bs, d_scale, d_vec);
	B200_LAUNCHED();
}

}  // namespace b200
