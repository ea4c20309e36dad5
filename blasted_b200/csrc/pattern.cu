/** \file pattern.cu
 * \brief Setup built on the device (K12): ILU(0) position lists and level schedules.
 *
 * ILU positions replace compute_ILU_positions_CSR_CSR (src/ilu_pattern.cpp:32-163 of the reference:
 * serial, linear inner_search).  Here: one thread per stored entry counts its (L-position,
 * U-position) pairs using binary search in the sorted upper part of the partner row, an exclusive
 * scan (CUB) gives posptr, and a second pass fills lowerp/upperp in the same ascending-k order, so
 * the three arrays are bit-identical to the reference's.
 *
 * Levels: (a) CONTIGUOUS reproduces computeLevels (src/levelschedule.cpp:12-71) exactly.  With
 * sorted columns and a structurally symmetric pattern its std::list bookkeeping reduces to: the
 * level starting at row s extends over consecutive rows r whose largest lower column d(r) < s.
 * (b) DAG computes true dependency wavefronts level[i] = 1 + max_{j<i, a_ij != 0} level[j] by
 * chaotic relaxation to the (unique) fixed point, then orders rows by level with a stable radix
 * sort (CUB).
 */
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda/functional>

namespace b200 {

// ------------------------------------------------------------------ ILU positions

__device__ __forceinline__ int search_sorted(const int *__restrict__ a, int lo, int hi, const int key)
{
	// [lo, hi) sorted ascending; returns position of key or -1
	hi -= 1;
	while(lo <= hi) {
		const int mid = (lo + hi) >> 1;
		const int c = __ldg(a + mid);
		if(c == key) return mid;
		if(c < key) lo = mid + 1; else hi = mid - 1;
	}
	return -1;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
ilu_positions_kernel(const long long nnzb, const int *__restrict__ browptr,
                     const int *__restrict__ bcolind, const int *__restrict__ diagind,
                     const int *__restrict__ browind, const int *__restrict__ posptr,
                     int *__restrict__ counts, int *__restrict__ lowerp, int *__restrict__ upperp,
                     int2 *__restrict__ pairs)
{
	const long long j = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(j >= nnzb) return;
	const int row = __ldg(browind + j);
	const int col = __ldg(bcolind + j);
	// l_ij (row > col): k-columns below col;  u_ij: k-columns below row   (ilu_pattern.cpp:46-48, :62-63)
	const int lim = row > col ? col : row;
	const int rs = __ldg(browptr + row), re = __ldg(browptr + row + 1);
	int cnt = 0;
	const int base = FILL ? __ldg(posptr + j) : 0;
	for(int k = rs; k < re; k++) {
		const int kc = __ldg(bcolind + k);
		if(kc >= lim) break;
		const int ipos = search_sorted(bcolind, __ldg(diagind + kc), __ldg(browptr + kc + 1), col);
		if(ipos >= 0) {
			if(FILL) {
				lowerp[base + cnt] = k; upperp[base + cnt] = ipos;
				if(pairs) pairs[base + cnt] = make_int2(k, ipos);
			}
			cnt++;
		}
	}
	if(!FILL) counts[j] = cnt;
}

/// 64-bit total of the per-entry counts (the reference accumulates in int, ilu_pattern.cpp:87)
__global__ void __launch_bounds__(256)
sum_counts_kernel(const long long n, const int *__restrict__ counts, unsigned long long *__restrict__ total)
{
	unsigned long long s = 0;
	for(long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n;
	    i += (long long)gridDim.x*blockDim.x)
		s += (unsigned long long)counts[i];
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
	if((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void lower_count_kernel(const int n, const int *__restrict__ browptr,
                                   const int *__restrict__ diagind, int *__restrict__ nl_row)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) nl_row[i] = diagind[i] - browptr[i];
	else if(i == n) nl_row[i] = 0;
}

__global__ void upper_work_flags_kernel(const long long n, const int4 *__restrict__ umeta,
                                        char *__restrict__ flags, const bool scalar_form)
{
	const long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(t >= n) return;
	const int4 m = umeta[t];
	const bool isdiag = scalar_form ? (m.w < 0) : (m.w >= 0);
	flags[t] = (m.z > m.y || isdiag) ? 1 : 0;         // has products, or is a diagonal entry
}

// ---- scalar split form (see IluPattern) ----

__global__ void __launch_bounds__(256)
scalar_lists_kernel(const long long nnz, const int *__restrict__ rowptr,
                    const int *__restrict__ colind, const int *__restrict__ diagind,
                    const int *__restrict__ rowind, const int *__restrict__ posptr,
                    const int *__restrict__ loff, int4 *__restrict__ lmeta, int4 *__restrict__ uall,
                    int *__restrict__ lcol, int *__restrict__ ucol)
{
	const long long j = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(j >= nnz) return;
	const int row = rowind[j], col = colind[j];
	const int rs = rowptr[row], dg = diagind[row], lo = loff[row];
	const int ps = posptr[j], pe = posptr[j+1];
	if(j < dg) {
		const int t = lo + (int)(j - rs);
		lmeta[t] = make_int4((int)j, col, ps, pe);
		lcol[t] = col;
	} else {
		const int tu = (rs - lo) + (int)(j - dg);          // index among upper entries incl. diagonals
		if(j == dg) uall[tu] = make_int4((int)j, ps, pe, ~row);
		else {
			const int ui = tu - (row + 1);                  // index among strict upper entries
			uall[tu] = make_int4((int)j, ps, pe, ui);
			ucol[ui] = col;
		}
	}
}

__global__ void part_max_len_kernel(const int n, const int *__restrict__ rowptr,
                                    const int *__restrict__ diagind, int *__restrict__ out)
{
	int ml = 0, mu = 0;
	for(int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		ml = max(ml, diagind[i] - rowptr[i]);
		mu = max(mu, rowptr[i+1] - diagind[i] - 1);
	}
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) {
		ml = max(ml, __shfl_down_sync(0xffffffffu, ml, off));
		mu = max(mu, __shfl_down_sync(0xffffffffu, mu, off));
	}
	if((threadIdx.x & 31) == 0) { atomicMax(out, ml); atomicMax(out + 1, mu); }
}

__global__ void scalar_uptr_kernel(const int n, const int *__restrict__ rowptr,
                                   const int *__restrict__ loff, int *__restrict__ uptr)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i <= n) uptr[i] = rowptr[i] - loff[i] - i;         // strict upper entries before row i
}

__global__ void __launch_bounds__(256)
scalar_pairs_kernel(const long long npos, const int *__restrict__ lowerp,
                    const int *__restrict__ upperp, const int *__restrict__ rowptr,
                    const int *__restrict__ diagind, const int *__restrict__ rowind,
                    const int *__restrict__ loff, const int *__restrict__ utpos,
                    int2 *__restrict__ spairs)
{
	const long long k = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(k >= npos) return;
	const int p = lowerp[k], q = upperp[k];
	const int rp = rowind[p], rq = rowind[q];
	const int li = loff[rp] + (p - rowptr[rp]);
	const int ui = (rowptr[rq] - loff[rq] - rq) + (q - diagind[rq] - 1);
	spairs[k] = make_int2(li, utpos ? utpos[ui] : ui);     // blocks: U partner in column-major storage
}

/// utpos[ut_order[d]] = d
__global__ void invert_order_kernel(const long long n, const int *__restrict__ order, int *__restrict__ pos)
{
	const long long d = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(d < n) pos[order[d]] = (int)d;
}

/// Per row, for the staged bs = 5 upper launch: dmeta = {entry of A, first L index, products, 0} of
/// the diagonal entry, dmeta_u = the U indices of its (at most 3) products.  bad[0] is raised unless
/// every diagonal entry has at most nkmax products whose L partners are the run l, l+1, ...; bad[1]
/// unless the U partners are a run as well (column-ordered U copy).
__global__ void __launch_bounds__(256)
diag_runs_kernel(const int nbrows, const int *__restrict__ uptr, const int4 *__restrict__ uall,
                 const int2 *__restrict__ spairs, const int nkmax, int4 *__restrict__ dmeta,
                 int4 *__restrict__ dmeta_u, int *__restrict__ bad)
{
	const int row = blockIdx.x*blockDim.x + threadIdx.x;
	if(row >= nbrows) return;
	const int4 m = uall[uptr[row] + row];            // the diagonal is the first upper entry of its row
	const int nk = m.z - m.y;
	bool ok = (m.w == ~row) && nk <= nkmax, urun = true;
	int u[3] = {0, 0, 0};
	int lfirst = 0;
	if(ok && nk > 0) {
		const int2 first = spairs[m.y];
		lfirst = first.x; u[0] = first.y;
		for(int k = 1; k < nk; k++) {
			const int2 p = spairs[m.y + k];
			ok &= (p.x == first.x + k);
			urun &= (p.y == first.y + k);
			if(k < 3) u[k] = p.y;
		}
	}
	if(!ok) bad[0] = 1;
	if(!urun) bad[1] = 1;
	dmeta[row] = make_int4(m.x, lfirst, nk, 0);
	dmeta_u[row] = make_int4(u[0], u[1], u[2], 0);
}

/// Per row, for the staged bs = 5 lower launch: {first A entry of the row, first L index, lower
/// entries, 0} and the columns of the (at most 3) lower entries
__global__ void __launch_bounds__(256)
lower_rows_kernel(const int nbrows, const int *__restrict__ browptr, const int *__restrict__ lptr,
                  const int *__restrict__ lcol, int4 *__restrict__ meta, int4 *__restrict__ cols)
{
	const int row = blockIdx.x*blockDim.x + threadIdx.x;
	if(row >= nbrows) return;
	const int lb = lptr[row], nl = lptr[row+1] - lb;
	int c[3] = {0, 0, 0};
	for(int e = 0; e < nl && e < 3; e++) c[e] = lcol[lb + e];
	meta[row] = make_int4(browptr[row], lb, nl, 0);
	cols[row] = make_int4(c[0], c[1], c[2], 0);
}

/// blocks: destinations of the strict upper entries move to the column-major copy
__global__ void upper_dest_transposed_kernel(const long long n, int4 *__restrict__ uall,
                                             const int *__restrict__ utpos)
{
	const long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(t >= n) return;
	int4 m = uall[t];
	if(m.w >= 0) { m.w = utpos[m.w]; uall[t] = m; }
}

__global__ void iota_kernel(int n, int *out);
__global__ void fill_int_kernel(int n, int v, int *out);

void build_ilu_pattern(const Mat& A, IluPattern& pl, cudaStream_t st)
{
	const long long nnzb = A.nnzb;
	pl.posptr.alloc(nnzb + 1);
	pl.npos = 0;
	if(nnzb == 0) {
		B200_CUDA(cudaMemsetAsync(pl.posptr, 0, sizeof(int), st));
		pl.built = true;
		return;
	}
	DevBuf<int> counts;
	counts.alloc(nnzb + 1);
	B200_CUDA(cudaMemsetAsync(counts.p + nnzb, 0, sizeof(int), st));
	const int grid = div_up(nnzb, 256);
	ilu_positions_kernel<false><<<grid, 256, 0, st>>>(nnzb, A.browptr, A.bcolind, A.diagind, A.browind,
	                                                  nullptr, counts, nullptr, nullptr, nullptr);
	B200_LAUNCHED();

	// total in 64 bits first
	DevBuf<unsigned long long> d_total;
	d_total.alloc(1);
	B200_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
	sum_counts_kernel<<<std::min(grid, 148*8), 256, 0, st>>>(nnzb, counts, d_total);
	B200_LAUNCHED();
	unsigned long long utotal = 0;
	B200_CUDA(cudaMemcpyAsync(&utotal, d_total.p, sizeof(utotal), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	const long long total = (long long)utotal;
	if(total >= 2147483647LL)
		throw Error("ILU position list exceeds int32 indexing (npos = " + std::to_string(total) + ")");
	size_t tb = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, pl.posptr.p, (int)(nnzb + 1), st);
	DevBuf<char> tmp;
	tmp.alloc(tb);
	B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, pl.posptr.p, (int)(nnzb + 1), st));
	g_launches.fetch_add(1);

	pl.npos = total;
	pl.lowerp.alloc(std::max<long long>(total, 1));
	pl.upperp.alloc(std::max<long long>(total, 1));
	if(total > 0) {
		ilu_positions_kernel<true><<<grid, 256, 0, st>>>(nnzb, A.browptr, A.bcolind, A.diagind,
		                                                 A.browind, pl.posptr, nullptr, pl.lowerp,
		                                                 pl.upperp, nullptr);
		B200_LAUNCHED();
	}

	// per-phase work lists: lower entries and upper entries (row-major order kept)
	{
		const int n = A.nbrows;
		DevBuf<int> nl_row, loff;
		nl_row.alloc((size_t)n + 1);
		loff.alloc((size_t)n + 1);
		lower_count_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, A.browptr, A.diagind, nl_row);
		B200_LAUNCHED();
		size_t tb2 = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb2, nl_row.p, loff.p, n + 1, st);
		DevBuf<char> tmp2;
		tmp2.alloc(tb2);
		B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, tb2, nl_row.p, loff.p, n + 1, st));
		g_launches.fetch_add(1);
		int nl = 0;
		B200_CUDA(cudaMemcpyAsync(&nl, loff.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		pl.nlower = nl;
		pl.nupper = nnzb - nl;
		pl.nstrict = pl.nupper - n;
		DevBuf<int4> *all_list = &pl.suall, *work_list = &pl.suwork;
		{
			// split form (scalars AND blocks): L part, strict U part (CSR each) and lists indexed into
			// the split value arrays
			pl.slmeta.alloc(std::max<long long>(pl.nlower, 1));
			pl.suall.alloc(std::max<long long>(pl.nupper, 1));
			pl.lcol.alloc(std::max<long long>(pl.nlower, 1));
			pl.ucol.alloc(std::max<long long>(pl.nstrict, 1));
			pl.lptr.alloc((size_t)n + 1);
			pl.uptr.alloc((size_t)n + 1);
			B200_CUDA(cudaMemcpyAsync(pl.lptr, loff, ((size_t)n + 1)*sizeof(int), cudaMemcpyDeviceToDevice, st));
			scalar_uptr_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, A.browptr, loff, pl.uptr);
			B200_LAUNCHED();
			scalar_lists_kernel<<<grid, 256, 0, st>>>(nnzb, A.browptr, A.bcolind, A.diagind, A.browind,
			                                          pl.posptr, loff, pl.slmeta, pl.suall, pl.lcol, pl.ucol);
			B200_LAUNCHED();
			{
				DevBuf<int> d_ml;
				d_ml.alloc(2);
				B200_CUDA(cudaMemsetAsync(d_ml, 0, 2*sizeof(int), st));
				part_max_len_kernel<<<std::min(div_up(n, 256), 148*8), 256, 0, st>>>(n, A.browptr, A.diagind, d_ml);
				B200_LAUNCHED();
				int h[2] = {0, 0};
				B200_CUDA(cudaMemcpyAsync(h, d_ml, 2*sizeof(int), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
				pl.max_lower_len = h[0]; pl.max_upper_len = h[1];
			}
			const int *utpos = nullptr;
			static const bool want_ut = getenv("B200_UT") != nullptr;   // A/B switch (development)
			if(A.bs == 5 && want_ut) {
				// OPTIONAL: block factors keep a second copy of the strict upper part in COLUMN order (UT): the
				// U_kj partners of a product then sit next to each other - for a diagonal entry
				// (i,i) of a structurally symmetric matrix the run UT[column i] is index-aligned with
				// the run L[row i], so the dominant products stream instead of gathering 200-byte
				// blocks (measured: isolated 200 B gathers reach 2.2-3.9 TB/s, streams 6.2).
				// ut_order[d] = row-major strict-upper index of the d-th block in column order (stable
				// sort by column keeps rows ascending inside a column); utpos is its inverse.
				pl.ut_order.alloc(std::max<long long>(pl.nstrict, 1));
				pl.utpos.alloc(std::max<long long>(pl.nstrict, 1));
				if(pl.nstrict > 0) {
					DevBuf<int> ids, keys_out;
					ids.alloc(pl.nstrict); keys_out.alloc(pl.nstrict);
					iota_kernel<<<div_up(pl.nstrict, 256), 256, 0, st>>>((int)pl.nstrict, ids);
					B200_LAUNCHED();
					int bits = 1;
					while((1LL << bits) < n && bits < 32) bits++;
					size_t tbs = 0;
					cub::DeviceRadixSort::SortPairs(nullptr, tbs, pl.ucol.p, keys_out.p, ids.p, pl.ut_order.p,
					                                (int)pl.nstrict, 0, bits, st);
					DevBuf<char> tmps;
					tmps.alloc(tbs);
					B200_CUDA(cub::DeviceRadixSort::SortPairs(tmps.p, tbs, pl.ucol.p, keys_out.p, ids.p,
					                                          pl.ut_order.p, (int)pl.nstrict, 0, bits, st));
					g_launches.fetch_add(1);
					invert_order_kernel<<<div_up(pl.nstrict, 256), 256, 0, st>>>(pl.nstrict, pl.ut_order, pl.utpos);
					B200_LAUNCHED();
					upper_dest_transposed_kernel<<<div_up(pl.nupper, 256), 256, 0, st>>>(pl.nupper, pl.suall, pl.utpos);
					B200_LAUNCHED();
					B200_CUDA(cudaStreamSynchronize(st));      // the sort's temporaries go out of scope
				}
				utpos = pl.utpos.p;
			}
			pl.spairs.alloc(std::max<long long>(total, 1));
			if(total > 0) {
				scalar_pairs_kernel<<<div_up(total, 256), 256, 0, st>>>(total, pl.lowerp, pl.upperp,
					A.browptr, A.diagind, A.browind, loff, utpos, pl.spairs);
				B200_LAUNCHED();
			}
		}

		// upper entries that change from sweep to sweep: those with products, and the diagonal
		// entries (which refresh the compact inverse).  The rest satisfy U_ij = A_ij identically.
		pl.nuwork = 0;
		if(pl.nupper == 0) work_list->alloc(1);
		if(pl.nupper > 0) {
			DevBuf<char> flags;
			DevBuf<int> d_nsel;
			flags.alloc(pl.nupper);
			d_nsel.alloc(1);
			upper_work_flags_kernel<<<div_up(pl.nupper, 256), 256, 0, st>>>(pl.nupper, all_list->p, flags, true);
			B200_LAUNCHED();
			// the work list is sized exactly (counted first): on a 7-point 512^3 factor it holds the
			// 1.3e8 diagonals, not the 5.4e8 upper entries (6.4 GB less)
			{
				size_t tbc = 0;
				cub::DeviceReduce::Sum(nullptr, tbc, flags.p, d_nsel.p, (int)pl.nupper, st);
				DevBuf<char> tmpc;
				tmpc.alloc(tbc);
				B200_CUDA(cub::DeviceReduce::Sum(tmpc.p, tbc, flags.p, d_nsel.p, (int)pl.nupper, st));
				g_launches.fetch_add(1);
				int cnt = 0;
				B200_CUDA(cudaMemcpyAsync(&cnt, d_nsel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
				work_list->alloc(std::max(cnt, 1));
			}
			size_t tb3 = 0;
			cub::DeviceSelect::Flagged(nullptr, tb3, all_list->p, flags.p, work_list->p, d_nsel.p,
			                           (int)pl.nupper, st);
			DevBuf<char> tmp3;
			tmp3.alloc(tb3);
			B200_CUDA(cub::DeviceSelect::Flagged(tmp3.p, tb3, all_list->p, flags.p, work_list->p,
			                                     d_nsel.p, (int)pl.nupper, st));
			g_launches.fetch_add(1);
			int nsel = 0;
			B200_CUDA(cudaMemcpyAsync(&nsel, d_nsel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
			B200_CUDA(cudaStreamSynchronize(st));
			pl.nuwork = nsel;
		}
	}
	// bs = 5, column-ordered U copy present: are the products of every diagonal entry index-aligned
	// runs (pairs (l+k, u+k)) of at most 3 blocks, and are the diagonals the only upper entries that
	// change?  Then the upper launch can stage whole runs with TMA (factor.cu, "staged").
	pl.diag_runs_ok = pl.diag_u_runs = false;
	if(A.bs == 5 && pl.nuwork == A.nbrows && A.nbrows > 0) {
		pl.dmeta.alloc(A.nbrows);
		pl.dmeta_u.alloc(A.nbrows);
		DevBuf<int> d_bad;
		d_bad.alloc(2);
		B200_CUDA(cudaMemsetAsync(d_bad, 0, 2*sizeof(int), st));
		diag_runs_kernel<<<div_up(A.nbrows, 256), 256, 0, st>>>(A.nbrows, pl.uptr, pl.suall, pl.spairs, 3,
		                                                      pl.dmeta, pl.dmeta_u, d_bad);
		B200_LAUNCHED();
		int bad[2] = {1, 1};
		B200_CUDA(cudaMemcpyAsync(bad, d_bad.p, 2*sizeof(int), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		pl.diag_runs_ok = (bad[0] == 0);
		pl.diag_u_runs = (bad[0] == 0 && bad[1] == 0);
		if(!pl.diag_runs_ok) { pl.dmeta.release(); pl.dmeta_u.release(); }
	}
	// ... and the lower launch, when no lower entry has products and no row more than 3 lower entries
	pl.lower_rows_ok = false;
	if(A.bs == 5 && A.nbrows > 0 && pl.max_lower_len <= 3 && pl.nlower > 0) {
		long long stats[5];
		pattern_stats(pl, stats, st);
		if(stats[3] == 0) {
			pl.lrow_meta.alloc(A.nbrows);
			pl.lrow_cols.alloc(A.nbrows);
			lower_rows_kernel<<<div_up(A.nbrows, 256), 256, 0, st>>>(A.nbrows, A.browptr, pl.lptr, pl.lcol,
			                                                       pl.lrow_meta, pl.lrow_cols);
			B200_LAUNCHED();
			pl.lower_rows_ok = true;
		}
	}
	B200_CUDA(cudaStreamSynchronize(st));
	pl.built = true;
}

void part_max_lengths(const Mat& A, int out[2], cudaStream_t st)
{
	out[0] = out[1] = 0;
	if(A.nbrows == 0) return;
	DevBuf<int> d_ml;
	d_ml.alloc(2);
	B200_CUDA(cudaMemsetAsync(d_ml, 0, 2*sizeof(int), st));
	part_max_len_kernel<<<std::min(div_up(A.nbrows, 256), 148*8), 256, 0, st>>>(A.nbrows, A.browptr, A.diagind, d_ml);
	B200_LAUNCHED();
	B200_CUDA(cudaMemcpyAsync(out, d_ml.p, 2*sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------ pattern statistics

__global__ void __launch_bounds__(256)
lower_products_kernel(const long long n, const int4 *__restrict__ lmeta, unsigned long long *__restrict__ total)
{
	unsigned long long s = 0;
	for(long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n;
	    i += (long long)gridDim.x*blockDim.x) {
		const int4 m = lmeta[i];
		s += (unsigned long long)(m.w - m.z);
	}
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
	if((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

/// {lower entries, upper entries incl. diagonals, upper work entries, products of lower entries,
/// products of upper entries}: what the byte formulas of the factor launches are made of
void pattern_stats(const IluPattern& pl, long long out[5], cudaStream_t st)
{
	DevBuf<unsigned long long> d;
	d.alloc(1);
	B200_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), st));
	if(pl.nlower > 0) {
		lower_products_kernel<<<(int)std::min<long long>(div_up(pl.nlower, 256), 148*8), 256, 0, st>>>(pl.nlower, pl.slmeta, d);
		B200_LAUNCHED();
	}
	unsigned long long h = 0;
	B200_CUDA(cudaMemcpyAsync(&h, d.p, sizeof(h), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	out[0] = pl.nlower; out[1] = pl.nupper; out[2] = pl.nuwork;
	out[3] = (long long)h; out[4] = pl.npos - (long long)h;
}

// ------------------------------------------------------------------ split CSR only (scalar SGS)

__global__ void __launch_bounds__(256)
split_entries_kernel(const long long nnz, const int *__restrict__ rowptr,
                     const int *__restrict__ colind, const int *__restrict__ diagind,
                     const int *__restrict__ rowind, const int *__restrict__ loff,
                     int *__restrict__ lcol, int *__restrict__ ucol, int *__restrict__ lentry,
                     int *__restrict__ uentry)
{
	const long long j = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(j >= nnz) return;
	const int row = rowind[j], col = colind[j];
	const int rs = rowptr[row], dg = diagind[row], lo = loff[row];
	if(j < dg) {
		const int t = lo + (int)(j - rs);
		lcol[t] = col; lentry[t] = (int)j;
	} else if(j > dg) {
		const int ui = (rs - lo - row) + (int)(j - dg - 1);
		ucol[ui] = col; uentry[ui] = (int)j;
	}
}

__global__ void gather_kernel(const long long n, const int *__restrict__ idx,
                              const double *__restrict__ in, double *__restrict__ out)
{
	const long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(t < n) out[t] = in[idx[t]];
}

__global__ void gather_blocks_kernel(const long long n, const int bs2, const int *__restrict__ idx,
                                     const double *__restrict__ in, double *__restrict__ out)
{
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= n*bs2) return;
	const long long t = e / bs2;
	out[e] = in[(size_t)idx[t]*bs2 + (e - t*bs2)];
}

void gather_split_values(const IluPattern& pl, const double *vals, double *lval, double *uval,
                         cudaStream_t st, int bs)
{
	const int bs2 = bs*bs;
	if(pl.nlower > 0) {
		if(bs == 1) gather_kernel<<<div_up(pl.nlower, 256), 256, 0, st>>>(pl.nlower, pl.lentry, vals, lval);
		else gather_blocks_kernel<<<div_up(pl.nlower*bs2, 256), 256, 0, st>>>(pl.nlower, bs2, pl.lentry, vals, lval);
		B200_LAUNCHED();
	}
	if(pl.nstrict > 0) {
		if(bs == 1) gather_kernel<<<div_up(pl.nstrict, 256), 256, 0, st>>>(pl.nstrict, pl.uentry, vals, uval);
		else gather_blocks_kernel<<<div_up(pl.nstrict*bs2, 256), 256, 0, st>>>(pl.nstrict, bs2, pl.uentry, vals, uval);
		B200_LAUNCHED();
	}
}

void build_split_csr(const Mat& A, IluPattern& pl, cudaStream_t st)
{
	const int n = A.nbrows;
	const long long nnz = A.nnzb;
	DevBuf<int> nl_row;
	nl_row.alloc((size_t)n + 1);
	pl.lptr.alloc((size_t)n + 1);
	pl.uptr.alloc((size_t)n + 1);
	lower_count_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, A.browptr, A.diagind, nl_row);
	B200_LAUNCHED();
	size_t tb = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tb, nl_row.p, pl.lptr.p, n + 1, st);
	DevBuf<char> tmp;
	tmp.alloc(tb);
	B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, nl_row.p, pl.lptr.p, n + 1, st));
	g_launches.fetch_add(1);
	int nl = 0;
	B200_CUDA(cudaMemcpyAsync(&nl, pl.lptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	pl.nlower = nl;
	pl.nupper = nnz - nl;
	pl.nstrict = pl.nupper - n;
	pl.lcol.alloc(std::max<long long>(pl.nlower, 1));
	pl.lentry.alloc(std::max<long long>(pl.nlower, 1));
	pl.ucol.alloc(std::max<long long>(pl.nstrict, 1));
	pl.uentry.alloc(std::max<long long>(pl.nstrict, 1));
	scalar_uptr_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, A.browptr, pl.lptr, pl.uptr);
	B200_LAUNCHED();
	if(nnz > 0) {
		split_entries_kernel<<<div_up(nnz, 256), 256, 0, st>>>(nnz, A.browptr, A.bcolind, A.diagind,
			A.browind, pl.lptr, pl.lcol, pl.ucol, pl.lentry, pl.uentry);
		B200_LAUNCHED();
	}
	DevBuf<int> d_ml;
	d_ml.alloc(2);
	B200_CUDA(cudaMemsetAsync(d_ml, 0, 2*sizeof(int), st));
	if(n > 0) {
		part_max_len_kernel<<<std::min(div_up(n, 256), 148*8), 256, 0, st>>>(n, A.browptr, A.diagind, d_ml);
		B200_LAUNCHED();
	}
	int h[2] = {0, 0};
	B200_CUDA(cudaMemcpyAsync(h, d_ml, 2*sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	pl.max_lower_len = h[0]; pl.max_upper_len = h[1];
	pl.split_built = true;
}

// ------------------------------------------------------------------ structural symmetry check

__global__ void symmetry_check_kernel(const long long nnzb, const int *__restrict__ browptr,
                                      const int *__restrict__ bcolind,
                                      const int *__restrict__ browind, int *__restrict__ nbad)
{
	const long long j = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(j >= nnzb) return;
	const int row = __ldg(browind + j), col = __ldg(bcolind + j);
	if(search_sorted(bcolind, __ldg(browptr + col), __ldg(browptr + col + 1), row) < 0)
		atomicAdd(nbad, 1);
}

// ------------------------------------------------------------------ contiguous levels

// computeLevels (src/levelschedule.cpp:12-71) reduces, for sorted columns and a structurally
// symmetric pattern, to: level k starts at s_k, s_0 = 0, s_{k+1} = min{ r > s_k : d(r) >= s_k } with
// d(r) = the largest column below r in row r (-1 if none).  d(r) >= s implies r > s, hence
//     next(s) = min{ r : d(r) >= s } = min over v >= s of m[v],   m[v] = min{ r : d(r) = v },
// a scatter-min followed by a suffix-min SCAN, and the level starts are the orbit of 0 under next,
// marked by pointer doubling (round k marks what is 2^k steps behind everything marked so far).
// All of it runs on the whole grid in O(n log n) work; the reference (and round 1: one CTA) walks
// the rows one after the other.  Bit-identical output (tests/test_gpu_setup.py).

/// mr[n-1-d(r)] = min(mr[..], r): m mirrored, so that an inclusive min-scan gives the suffix minima
__global__ void __launch_bounds__(256)
level_scatter_min_kernel(const int n, const int *__restrict__ browptr, const int *__restrict__ bcolind,
                         const int *__restrict__ diagind, int *__restrict__ mr)
{
	const int r = blockIdx.x*blockDim.x + threadIdx.x;
	if(r >= n) return;
	const int dp = diagind[r];
	if(dp > browptr[r]) atomicMin(mr + (n - 1 - bcolind[dp-1]), r);
}

/// J[s] = next(s) = sr[n-1-s] for s < n, J[n] = n; flag = {0}
__global__ void __launch_bounds__(256)
level_next_kernel(const int n, const int *__restrict__ sr, int *__restrict__ J, char *__restrict__ flag)
{
	const int s = blockIdx.x*blockDim.x + threadIdx.x;
	if(s > n) return;
	J[s] = (s < n) ? sr[n - 1 - s] : n;
	flag[s] = (s == 0) ? 1 : 0;
}

/// one doubling round: mark J[u] for every marked u, J2 = J o J
__global__ void __launch_bounds__(256)
level_double_kernel(const int n, const int *__restrict__ J, int *__restrict__ J2, char *flag)
{
	const int u = blockIdx.x*blockDim.x + threadIdx.x;
	if(u > n) return;
	const int j = J[u];
	if(flag[u] && j < n) flag[j] = 1;
	J2[u] = J[j];
}

// ------------------------------------------------------------------ DAG levels

__global__ void __launch_bounds__(256)
dag_relax_kernel(const int nbrows, const int *__restrict__ browptr, const int *__restrict__ bcolind,
                 const int *__restrict__ diagind, int *level, int *__restrict__ changed)
{
	const int row = blockIdx.x*blockDim.x + threadIdx.x;
	if(row >= nbrows) return;
	int lv = 0;
	const int s = __ldg(browptr + row), e = __ldg(diagind + row);
	for(int jj = s; jj < e; jj++) {
		const int l = __ldcg(level + __ldg(bcolind + jj)) + 1;
		lv = max(lv, l);
	}
	if(lv != __ldcg(level + row)) {
		level[row] = lv;
		*changed = 1;
	}
}

/// DAG levels in ONE forward pass: level[i] = 1 + max_{j<i, a_ij != 0} level[j], level = -1 until
/// known.  A thread per row in natural order; a row reads only lower-numbered rows, CTAs take
/// their rows from a ticket counter, so everything a row waits for belongs to a CTA that has
/// started.  Every pass of the loop each unfinished lane looks at its next unknown dependency
/// once (polled at L2) - lanes of one warp may depend on each other (row i on row i-1), so nobody
/// spins alone.  Replaces relax-until-nothing-changes (8 launches per host round trip, #levels/8
/// round trips: 156 ms on the 7-point 256^3 matrix).
__global__ void __launch_bounds__(256)
dag_levels_syncfree_kernel(const int nbrows, const int *__restrict__ browptr,
                           const int *__restrict__ bcolind, const int *__restrict__ diagind,
                           int *level, int *__restrict__ ticket, int *__restrict__ err)
{
	__shared__ int s_cta;
	if(threadIdx.x == 0) s_cta = atomicAdd(ticket, 1);
	__syncthreads();
	const int row = s_cta*blockDim.x + threadIdx.x;
	bool done = row >= nbrows;
	int jj = 0, je = 0, lv = 0;
	if(!done) { jj = __ldg(browptr + row); je = __ldg(diagind + row); }
	int spins = 0;
	while(true) {
		if(!done) {
			// consume every dependency that is already known, stop at the first unknown one
			while(jj < je) {
				int l;
				asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(l) : "l"(level + __ldg(bcolind + jj)) : "memory");
				if(l < 0) break;
				lv = max(lv, l + 1);
				jj++;
			}
			if(jj == je) {
				*((volatile int*)(level + row)) = lv;
				done = true;
			}
		}
		if(__all_sync(0xffffffffu, done)) break;
		if(++spins > (1 << 22)) { *err = 1; break; }
	}
}

__global__ void fill_int_kernel(int n, int v, int *out)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) out[i] = v;
}

__global__ void iota_kernel(int n, int *out)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) out[i] = i;
}

__global__ void level_bounds_kernel(int n, const int *__restrict__ sorted_level, int *__restrict__ level_ptr)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	const int l = sorted_level[i];
	if(i == 0 || sorted_level[i-1] != l) level_ptr[l] = i;
	if(i == n-1) level_ptr[l+1] = n;
}

void build_levels(const Mat& A, Levels& lv, int mode, cudaStream_t st)
{
	const int n = A.nbrows;
	lv.mode = mode;
	lv.nlevels = 0;
	lv.level_ptr.assign(1, 0);
	if(n == 0) { lv.built = true; return; }

	if(mode == B200_LEVELS_CONTIGUOUS) {
		// "(jnode must be found because the sparsity structure is symmetric)" levelschedule.cpp:55-57
		DevBuf<int> d_bad;
		d_bad.alloc(1);
		B200_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
		symmetry_check_kernel<<<div_up(A.nnzb, 256), 256, 0, st>>>(A.nnzb, A.browptr, A.bcolind,
		                                                          A.browind, d_bad);
		B200_LAUNCHED();
		int bad = 0;
		B200_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		if(bad) throw Error("Faulty dependency list!");

		{
			DevBuf<int> mr, sr, J, J2, d_nl;
			DevBuf<char> flag;
			mr.alloc(n); sr.alloc(n); J.alloc((size_t)n + 1); J2.alloc((size_t)n + 1); flag.alloc((size_t)n + 1);
			d_nl.alloc(1);
			// m[v] = n ("no row") everywhere, then the scatter-min
			fill_int_kernel<<<div_up(n, 256), 256, 0, st>>>(n, n, mr);
			B200_LAUNCHED();
			level_scatter_min_kernel<<<div_up(n, 256), 256, 0, st>>>(n, A.browptr, A.bcolind, A.diagind, mr);
			B200_LAUNCHED();
			size_t tb = 0;
			cub::DeviceScan::InclusiveScan(nullptr, tb, mr.p, sr.p, cuda::minimum<>{}, n, st);
			DevBuf<char> tmp;
			tmp.alloc(tb);
			B200_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, mr.p, sr.p, cuda::minimum<>{}, n, st));
			g_launches.fetch_add(1);
			level_next_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, sr, J, flag);
			B200_LAUNCHED();
			int *Ja = J.p, *Jb = J2.p;
			for(long long reach = 1; reach <= (long long)n; reach *= 2) {
				level_double_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(n, Ja, Jb, flag);
				B200_LAUNCHED();
				std::swap(Ja, Jb);
			}
			// level starts = the marked rows in ascending order, then n
			lv.d_level_ptr.alloc((size_t)n + 1);
			size_t tb2 = 0;
			thrust::counting_iterator<int> ids(0);
			cub::DeviceSelect::Flagged(nullptr, tb2, ids, flag.p, lv.d_level_ptr.p, d_nl.p, n, st);
			DevBuf<char> tmp2;
			tmp2.alloc(tb2);
			B200_CUDA(cub::DeviceSelect::Flagged(tmp2.p, tb2, ids, flag.p, lv.d_level_ptr.p, d_nl.p, n, st));
			g_launches.fetch_add(1);
			B200_CUDA(cudaMemcpyAsync(&lv.nlevels, d_nl, sizeof(int), cudaMemcpyDeviceToHost, st));
			B200_CUDA(cudaStreamSynchronize(st));
			B200_CUDA(cudaMemcpyAsync(lv.d_level_ptr.p + lv.nlevels, &n, sizeof(int), cudaMemcpyHostToDevice, st));
		}
		lv.level_ptr.resize(lv.nlevels + 1);
		B200_CUDA(cudaMemcpyAsync(lv.level_ptr.data(), lv.d_level_ptr, (lv.nlevels+1)*sizeof(int),
		                          cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		lv.level_rows.release();
		lv.built = true;
		return;
	}

	// DAG wavefronts.  The backward substitutions walk the same levels in reverse, which is only a
	// valid order when the pattern is structurally symmetric - the assumption the reference states
	// at levelschedule.cpp:55-57 (and throws on); checked here as well.
	{
		DevBuf<int> d_bad;
		d_bad.alloc(1);
		B200_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
		symmetry_check_kernel<<<div_up(A.nnzb, 256), 256, 0, st>>>(A.nnzb, A.browptr, A.bcolind,
		                                                          A.browind, d_bad);
		B200_LAUNCHED();
		int bad = 0;
		B200_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		if(bad) throw Error("Faulty dependency list!");
	}
	DevBuf<int> level, changed, rows_in, level_sorted;
	level.alloc(n);
	changed.alloc(2);
	const int grid = div_up(n, 256);
	static const bool relax = getenv("B200_LEVELS_RELAX") != nullptr;     // A/B switch (development)
	if(!relax) {
		B200_CUDA(cudaMemsetAsync(level, 0xff, n*sizeof(int), st));          // -1 = not known yet
		B200_CUDA(cudaMemsetAsync(changed, 0, 2*sizeof(int), st));           // {ticket, error}
		dag_levels_syncfree_kernel<<<grid, 256, 0, st>>>(n, A.browptr, A.bcolind, A.diagind, level,
		                                                 changed.p, changed.p + 1);
		B200_LAUNCHED();
		int h_err = 0;
		B200_CUDA(cudaMemcpyAsync(&h_err, changed.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
		if(h_err) throw Error("DAG levels: a dependency never became known");
	} else {
		B200_CUDA(cudaMemsetAsync(level, 0, n*sizeof(int), st));
		int h_changed = 1, rounds = 0;
		while(h_changed) {
			B200_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
			for(int rep = 0; rep < 8; rep++) {
				dag_relax_kernel<<<grid, 256, 0, st>>>(n, A.browptr, A.bcolind, A.diagind, level, changed);
				B200_LAUNCHED();
			}
			B200_CUDA(cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
			B200_CUDA(cudaStreamSynchronize(st));
			if(++rounds > n + 8) throw Error("DAG level relaxation did not converge");
		}
	}
	rows_in.alloc(n);
	level_sorted.alloc(n);
	lv.level_rows.alloc(n);
	iota_kernel<<<grid, 256, 0, st>>>(n, rows_in);
	B200_LAUNCHED();
	size_t tb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb, level.p, level_sorted.p, rows_in.p, lv.level_rows.p,
	                                n, 0, 32, st);
	DevBuf<char> tmp;
	tmp.alloc(tb);
	B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, level.p, level_sorted.p, rows_in.p,
	                                          lv.level_rows.p, n, 0, 32, st));
	g_launches.fetch_add(1);
	int maxlevel = 0;
	B200_CUDA(cudaMemcpyAsync(&maxlevel, level_sorted.p + (n-1), sizeof(int), cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	lv.nlevels = maxlevel + 1;
	lv.d_level_ptr.alloc((size_t)lv.nlevels + 1);
	level_bounds_kernel<<<grid, 256, 0, st>>>(n, level_sorted, lv.d_level_ptr);
	B200_LAUNCHED();
	lv.level_ptr.resize(lv.nlevels + 1);
	B200_CUDA(cudaMemcpyAsync(lv.level_ptr.data(), lv.d_level_ptr, (lv.nlevels+1)*sizeof(int),
	                          cudaMemcpyDeviceToHost, st));
	B200_CUDA(cudaStreamSynchronize(st));
	lv.built = true;
}

}  // namespace b200
