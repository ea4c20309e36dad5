/** \file blockops.cuh
 * \brief Register-level access to small dense blocks and vector segments.
 *
 * Device layout of a block (decided per block size, hidden behind BlkIO):
 *   bs == 4 : ROW-major, 128-byte aligned.  Lane r of a 4-lane group owns row r and moves it with ONE
 *             256-bit access (LDG.E.ENL2.256 / STG.E.ENL2.256, new on sm_100): a group touches each
 *             128-byte line with a single instruction, which is what keeps the L1 tag stage (one
 *             line per cycle per SM) off the critical path of these HBM-bound kernels.
 *   others  : column-major, dense (bs=5: 200-byte blocks, only 8-byte aligned), 64-bit accesses; a
 *             group reads one block column (bs*8 contiguous bytes) per instruction.
 * The caller's layout (Eigen ColMajor / RowMajor) is converted at the boundary (storage.cu).
 *
 * Arithmetic is written against logical indices: a row is v[c] = B(r,c); a whole block is handed
 * out "column-major logical", d[c*BS + m] = B(m,c), whatever the storage order.
 */
#ifndef B200_BLOCKOPS_CUH
#define B200_BLOCKOPS_CUH

#include <cuda_runtime.h>

namespace b200 {

// ------------------------------------------------------------------ 256-bit global accesses

/// read-only data (matrix values, lists): non-coherent path
__device__ __forceinline__ void ld256_nc(const double *p, double &a, double &b, double &c, double &d)
{
	asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
/// data other CTAs may be rewriting during this launch (the chaotic iterate): read at L2.
/// Relaxed: not `volatile`, no memory clobber, so the compiler may batch these gathers with the
/// surrounding loads (more requests in flight); chaotic iteration does not care which of the
/// concurrently written values a gather observes.
__device__ __forceinline__ void ld256_cg(const double *p, double &a, double &b, double &c, double &d)
{
	asm("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
	    : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
/// same load, ordered against the surrounding stores (read-before-overwrite comparisons)
__device__ __forceinline__ void ld256_cg_ordered(const double *p, double &a, double &b, double &c, double &d)
{
	asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
	             : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <bool ITER> __device__ __forceinline__ double ld1(const double *p)
{
	return ITER ? __ldcg(p) : __ldg(p);
}

// ------------------------------------------------------------------ blocks

template <int BS>
struct BlkIO {
	static constexpr bool ROWMAJOR = false;
	/// v[c] = B(r,c)
	template <bool ITER>
	static __device__ __forceinline__ void load_row(const double *blk, const int r, double (&v)[BS])
	{
#pragma unroll
		for(int c = 0; c < BS; c++) v[c] = ld1<ITER>(blk + c*BS + r);
	}
	static __device__ __forceinline__ void load_row_ordered(const double *blk, const int r, double (&v)[BS])
	{
#pragma unroll
		for(int c = 0; c < BS; c++) v[c] = *((const volatile double*)(blk + c*BS + r));
	}
	static __device__ __forceinline__ void store_row(double *blk, const int r, const double (&v)[BS])
	{
#pragma unroll
		for(int c = 0; c < BS; c++) blk[c*BS + r] = v[c];
	}
	/// d[c*BS + m] = B(m,c)
	template <bool ITER>
	static __device__ __forceinline__ void load_full(const double *blk, double (&d)[BS*BS])
	{
#pragma unroll
		for(int e = 0; e < BS*BS; e++) d[e] = ld1<ITER>(blk + e);
	}
	/// position of B(r,c) inside the stored block
	static __device__ __forceinline__ int at(const int r, const int c) { return c*BS + r; }
	/// Whole block for every lane of the group, WARP-COLLECTIVE (all 32 lanes must call it):
	/// each lane loads only its own row (blk == nullptr: zeros) and the rows are exchanged with
	/// shuffles.  25 broadcast loads per lane would cost bs^2 L1 tag look-ups per block; this
	/// costs bs.  d[c*BS + m] = B(m,c).
	template <bool ITER>
	static __device__ __forceinline__ void load_full_group(const double *blk, const int r,
	                                                       const int gbase, double (&d)[BS*BS])
	{
		double row[BS];
#pragma unroll
		for(int c = 0; c < BS; c++) row[c] = 0;
		if(blk) load_row<ITER>(blk, r, row);
#pragma unroll
		for(int c = 0; c < BS; c++)
#pragma unroll
			for(int m = 0; m < BS; m++)
				d[c*BS + m] = __shfl_sync(0xffffffffu, row[c], min(gbase + m, 31));
	}
};

template <>
struct BlkIO<4> {
	static constexpr bool ROWMAJOR = true;
	template <bool ITER>
	static __device__ __forceinline__ void load_row(const double *blk, const int r, double (&v)[4])
	{
		if(ITER) ld256_cg(blk + 4*r, v[0], v[1], v[2], v[3]);
		else ld256_nc(blk + 4*r, v[0], v[1], v[2], v[3]);
	}
	static __device__ __forceinline__ void load_row_ordered(const double *blk, const int r, double (&v)[4])
	{
		ld256_cg_ordered(blk + 4*r, v[0], v[1], v[2], v[3]);
	}
	static __device__ __forceinline__ void store_row(double *blk, const int r, const double (&v)[4])
	{
		st256(blk + 4*r, v[0], v[1], v[2], v[3]);
	}
	template <bool ITER>
	static __device__ __forceinline__ void load_full(const double *blk, double (&d)[16])
	{
#pragma unroll
		for(int m = 0; m < 4; m++) {
			double t0, t1, t2, t3;
			if(ITER) ld256_cg(blk + 4*m, t0, t1, t2, t3);
			else ld256_nc(blk + 4*m, t0, t1, t2, t3);
			d[0*4 + m] = t0; d[1*4 + m] = t1; d[2*4 + m] = t2; d[3*4 + m] = t3;
		}
	}
	static __device__ __forceinline__ int at(const int r, const int c) { return r*4 + c; }
	/// same interface as the generic version; 4 broadcast 256-bit loads are cheap enough (4 tags)
	template <bool ITER>
	static __device__ __forceinline__ void load_full_group(const double *blk, const int, const int,
	                                                       double (&d)[16])
	{
		if(blk) load_full<ITER>(blk, d);
		else {
#pragma unroll
			for(int e = 0; e < 16; e++) d[e] = 0;
		}
	}
};

// ------------------------------------------------------------------ group products via shuffles
//
// A group of BS lanes holds a block one row per lane.  Products against a partner block are
// streamed: the partner's row held by lane m is broadcast entry by entry with warp shuffles, so no
// lane ever keeps a whole bs x bs partner in registers (register pressure, hence occupancy, is
// what bounds these latency-sensitive kernels) and no lane issues bs^2 broadcast loads (L1 tag
// stage).  WARP-COLLECTIVE: all 32 lanes must call these with the same control flow.

/// acc(r,:) -= L(r,:) * U  with  lrow = L(r,:) of this lane and urow = U(r,:) of this lane
template <int BS>
__device__ __forceinline__ void group_mul_sub(double (&acc)[BS], const double (&lrow)[BS],
                                              const double (&urow)[BS], const int gbase)
{
#pragma unroll
	for(int c = 0; c < BS; c++)
#pragma unroll
		for(int m = 0; m < BS; m++)
			acc[c] = fma(-lrow[m], __shfl_sync(0xffffffffu, urow[c], min(gbase + m, 31)), acc[c]);
}

/// out(r,:) = S(r,:) * D  with  srow = S(r,:) and drow = D(r,:) of this lane
template <int BS>
__device__ __forceinline__ void group_mul(double (&out)[BS], const double (&srow)[BS],
                                          const double (&drow)[BS], const int gbase)
{
#pragma unroll
	for(int c = 0; c < BS; c++) {
		double a = 0;
#pragma unroll
		for(int m = 0; m < BS; m++)
			a = fma(srow[m], __shfl_sync(0xffffffffu, drow[c], min(gbase + m, 31)), a);
		out[c] = a;
	}
}

/// v[idx] with a lane-dependent idx (a chain of selects; the array stays in registers)
template <int BS>
__device__ __forceinline__ double pick(const double (&v)[BS], const int idx)
{
	double x = v[0];
#pragma unroll
	for(int i = 1; i < BS; i++) x = (idx == i) ? v[i] : x;
	return x;
}

/// Whether the device stores blocks of this size row-major
inline bool device_rowmajor(const int bs) { return bs == 4; }

// ------------------------------------------------------------------ vector segments

/// xv[c] = x[c], c < BS.  VEC: the segment is 32-byte aligned (bs == 4 and an aligned base).
template <int BS, bool ITER, bool VEC>
__device__ __forceinline__ void load_seg(const double *x, double (&xv)[BS])
{
	if(BS == 4 && VEC) {
		if(ITER) ld256_cg(x, xv[0], xv[1], xv[2], xv[3]);
		else ld256_nc(x, xv[0], xv[1], xv[2], xv[3]);
	} else {
#pragma unroll
		for(int c = 0; c < BS; c++) xv[c] = ld1<ITER>(x + c);
	}
}

inline bool aligned32(const void *p) { return (reinterpret_cast<size_t>(p) & 31) == 0; }

// ------------------------------------------------------------------ cooperative block inverse

/// Inverse of a bs x bs block held one row per lane (lane base + m holds row m in a[]), by
/// Gauss-Jordan elimination with partial pivoting across the group: the pivot row travels by
/// shuffles, every lane eliminates in its own row.  10 (bs 5) / 8 (bs 4) doubles of state per lane
/// instead of the bs^2 + 2 bs of the redundant per-lane elimination (solve_right), which is what lets
/// the factor launches refresh U_ii^-1 themselves without losing a resident CTA.
///
/// Shuffles share the L1 data pipe with the global loads (ncu on the bs = 5 factor launches:
/// l1tex__data_pipe_lsu_wavefronts at 81 % of peak, half of it shuffles - a 64-bit shuffle is two
/// wavefronts), so the exchange is kept minimal: pivot candidates are compared on the high words of
/// |a_mk| (32-bit shuffles; candidates equal in their upper 32 bits tie, lowest row wins), columns
/// <= k of the pivot row are not exchanged (they are never read again): 105 (bs 5) shuffle
/// wavefronts per inverse instead of 160.  (redux.sync + ballot on per-group masks for the pivot
/// search was measured: 15-35 % SLOWER launches - sub-warp masks serialise.)
/// On return `inv` holds row `prow` of the inverse (rows end up where their pivots were found).
/// Must be called by all lanes of the warp; `base` = first lane of the caller's group.
template <int BS>
__device__ __forceinline__ void group_inverse(double (&a)[BS], double (&inv)[BS], const int base,
                                              const int r, int& prow)
{
#pragma unroll
	for(int c = 0; c < BS; c++) inv[c] = (c == r) ? 1.0 : 0.0;
	prow = -1;
#pragma unroll
	for(int k = 0; k < BS; k++) {
		// pivot: largest |a_mk| among the rows not used yet, compared on the high words (one
		// 32-bit shuffle per candidate instead of two); lowest row on ties, same on every lane
		const int mine = (prow < 0) ? __double2hiint(fabs(a[k])) : -1;
		int best = -2, bl = 0;
#pragma unroll
		for(int m = 0; m < BS; m++) {
			const int v = __shfl_sync(0xffffffffu, mine, min(base + m, 31));
			if(v > best) { best = v; bl = m; }
		}
		const int src = min(base + bl, 31);
		const bool is_pivot = (base + r == src);
		const double rinv = 1.0/__shfl_sync(0xffffffffu, a[k], src);
		const double f = a[k]*rinv;
#pragma unroll
		for(int c = k + 1; c < BS; c++) {
			const double pa = __shfl_sync(0xffffffffu, a[c], src);
			a[c] = is_pivot ? pa*rinv : fma(-f, pa, a[c]);
		}
#pragma unroll
		for(int c = 0; c < BS; c++) {
			const double pi = __shfl_sync(0xffffffffu, inv[c], src);
			inv[c] = is_pivot ? pi*rinv : fma(-f, pi, inv[c]);
		}
		if(is_pivot) prow = k;
	}
}

// ------------------------------------------------------------------ dense solve in registers

/// Solves x * D = s for the row vector x, with D handed over as d[c*BS + m] = D(m,c).
/// Equivalent to M x^T = s^T with M = D^T, M(i,j) = d[i*BS + j]: Gaussian elimination with partial
/// pivoting, fully unrolled, done redundantly by every lane of a group for its own right-hand side.
/// Stands in for `sum * U_jj.inverse()` (src/kernels/kernels_ilu0_factorize.hpp:91 of the
/// reference) and for `.inverse()` itself (row r of D^-1 solves x D = e_r).
// Elimination step P of solve_right as a template so that every loop bound is a compile-time
// constant: with runtime-triangular bounds the compiler stops unrolling for BS = 5 and the block
// falls out of registers into local memory.
template <int BS, int P>
struct SolveStep {
	static __device__ __forceinline__ void eliminate(double (&d)[BS*BS], double (&s)[BS],
	                                                 double (&pinv)[BS])
	{
		// bring the largest |M(q,P)|, q >= P, to row P by successive conditional swaps
#pragma unroll
		for(int q = P+1; q < BS; q++) {
			const bool sw = fabs(d[q*BS+P]) > fabs(d[P*BS+P]);
#pragma unroll
			for(int j = P; j < BS; j++) {
				const double a = d[P*BS+j], b = d[q*BS+j];
				d[P*BS+j] = sw ? b : a;
				d[q*BS+j] = sw ? a : b;
			}
			const double a = s[P], b = s[q];
			s[P] = sw ? b : a;
			s[q] = sw ? a : b;
		}
		pinv[P] = 1.0/d[P*BS+P];
#pragma unroll
		for(int i = P+1; i < BS; i++) {
			const double f = d[i*BS+P]*pinv[P];
#pragma unroll
			for(int j = P+1; j < BS; j++)
				d[i*BS+j] = fma(-f, d[P*BS+j], d[i*BS+j]);
			s[i] = fma(-f, s[P], s[i]);
		}
		SolveStep<BS,P+1>::eliminate(d, s, pinv);
	}
	static __device__ __forceinline__ void substitute(const double (&d)[BS*BS], const double (&s)[BS],
	                                                  const double (&pinv)[BS], double (&x)[BS])
	{
		// rows BS-1-P: back substitution from the bottom
		constexpr int i = BS - 1 - P;
		double t = s[i];
#pragma unroll
		for(int j = i+1; j < BS; j++)
			t = fma(-d[i*BS+j], x[j], t);
		x[i] = t*pinv[i];
		SolveStep<BS,P+1>::substitute(d, s, pinv, x);
	}
};
template <int BS>
struct SolveStep<BS,BS> {
	static __device__ __forceinline__ void eliminate(double (&)[BS*BS], double (&)[BS], double (&)[BS]) {}
	static __device__ __forceinline__ void substitute(const double (&)[BS*BS], const double (&)[BS],
	                                                  const double (&)[BS], double (&)[BS]) {}
};

template <int BS>
__device__ __forceinline__ void solve_right(double (&d)[BS*BS], double (&s)[BS], double (&x)[BS])
{
	double pinv[BS];
	SolveStep<BS,0>::eliminate(d, s, pinv);
	SolveStep<BS,0>::substitute(d, s, pinv, x);
}

}  // namespace b200
#endif
