/** \file blasted_oracle.c
 * \brief Plain-C CPU restatement of the BLASTed asynchronous-preconditioner hot path.
 *
 * TEST INFRASTRUCTURE ONLY - see blasted_oracle.h for the rules on who may call this and for the
 * parity status (PINNED against the unmodified reference build in oracle/_ref and tests/golden/).
 * Citations are relative to /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "blasted_oracle.h"

#define BIDX(rowmajor, bs, r, c) ((rowmajor) ? (r)*(bs)+(c) : (c)*(bs)+(r))

/* ---------- small dense helpers ---------- */

/* y += A x for one block */
static void blk_gemv_acc(int bs, int rm, const double *a, const double *x, double *y)
{
	for(int r = 0; r < bs; r++) {
		double s = 0;
		for(int c = 0; c < bs; c++)
			s += a[BIDX(rm,bs,r,c)] * x[c];
		y[r] += s;
	}
}

/* c = a*b */
static void blk_mul(int bs, int rm, const double *a, const double *b, double *c)
{
	for(int j = 0; j < bs; j++)
		for(int i = 0; i < bs; i++) {
			double s = 0;
			for(int k = 0; k < bs; k++)
				s += a[BIDX(rm,bs,i,k)] * b[BIDX(rm,bs,k,j)];
			c[BIDX(rm,bs,i,j)] = s;
		}
}

/* Gauss-Jordan with partial pivoting.  The reference calls Eigen's .inverse()
 * (src/kernels/kernels_ilu0_factorize.hpp:91, src/async_blockilu_factor.cpp:146,
 * src/solverops_jacobi.cpp:45); Eigen is an un-vendored dependency (README.md:13-16,
 * "Eigen 3.3.4 or later"), so the published algorithm class (pivoted elimination) is restated. */
void orc_block_inverse(int bs, int rm, const double *ain, double *ainv)
{
	double a[ORC_MAX_BS][2*ORC_MAX_BS];
	for(int i = 0; i < bs; i++)
		for(int j = 0; j < bs; j++) {
			a[i][j] = ain[BIDX(rm,bs,i,j)];
			a[i][bs+j] = (i == j) ? 1.0 : 0.0;
		}
	for(int c = 0; c < bs; c++) {
		int p = c;
		double best = fabs(a[c][c]);
		for(int i = c+1; i < bs; i++)
			if(fabs(a[i][c]) > best) { best = fabs(a[i][c]); p = i; }
		if(p != c)
			for(int j = 0; j < 2*bs; j++) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
		const double piv = 1.0/a[c][c];
		for(int j = 0; j < 2*bs; j++) a[c][j] *= piv;
		for(int i = 0; i < bs; i++) {
			if(i == c) continue;
			const double f = a[i][c];
			for(int j = 0; j < 2*bs; j++) a[i][j] -= f*a[c][j];
		}
	}
	for(int i = 0; i < bs; i++)
		for(int j = 0; j < bs; j++)
			ainv[BIDX(rm,bs,i,j)] = a[i][bs+j];
}

/* src/kernels/kernels_ilu0_factorize.hpp:61-69 */
static void scale_block(int bs, int rm, const double *scale, int brow, int bcol, double *blk)
{
	for(int j = 0; j < bs; j++)
		for(int i = 0; i < bs; i++)
			blk[BIDX(rm,bs,i,j)] *= scale[brow*bs+i] * scale[bcol*bs+j];
}

/* ---------- K9 SpMV ---------- */

/* src/blas/matvecs.cpp:25-48 (BSR), :78-91 (CSR) */
void orc_spmv(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
              const double *vals, const double *x, double *y)
{
	if(bs == 1) {
		for(int i = 0; i < nbrows; i++) {
			double s = 0;
			for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
				s += vals[jj]*x[bcolind[jj]];
			y[i] = s;
		}
		return;
	}
	const int bs2 = bs*bs;
	for(int i = 0; i < nbrows; i++) {
		double acc[ORC_MAX_BS];
		for(int r = 0; r < bs; r++) acc[r] = 0;
		for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
			blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, x + (size_t)bcolind[jj]*bs, acc);
		for(int r = 0; r < bs; r++) y[(size_t)i*bs+r] = acc[r];
	}
}

/* src/blas/matvecs.cpp:51-75 (BSR), :94-108 (CSR): z = a A x + b y */
void orc_gemv3(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
               const double *vals, double a, const double *x, double b, const double *y, double *z)
{
	if(bs == 1) {
		for(int i = 0; i < nbrows; i++) {
			double s = b*y[i];
			for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
				s += a*vals[jj]*x[bcolind[jj]];
			z[i] = s;
		}
		return;
	}
	const int bs2 = bs*bs;
	for(int i = 0; i < nbrows; i++) {
		double acc[ORC_MAX_BS];
		for(int r = 0; r < bs; r++) acc[r] = b*y[(size_t)i*bs+r];
		for(int jj = browptr[i]; jj < browptr[i+1]; jj++) {
			double t[ORC_MAX_BS];
			for(int r = 0; r < bs; r++) t[r] = 0;
			blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, x + (size_t)bcolind[jj]*bs, t);
			for(int r = 0; r < bs; r++) acc[r] += a*t[r];
		}
		for(int r = 0; r < bs; r++) z[(size_t)i*bs+r] = acc[r];
	}
}

/* ---------- T4 ILU positions ---------- */

/* src/helper_algorithms.hpp:39-49 */
static int inner_search(const int *aind, int start, int end, int tofind)
{
	for(int j = start; j < end; j++)
		if(aind[j] == tofind) return j;
	return -1;
}

/* src/ilu_pattern.cpp:32-163.  For the entry at position j = (irow, col): every k-position in row
 * irow with column kc < min(irow, col) such that row kc stores column col at or after its diagonal;
 * pairs are listed in ascending k. */
long long orc_ilu_positions(int nbrows, const int *browptr, const int *bcolind, const int *diagind,
                            int *posptr, int *lowerp, int *upperp)
{
	long long total = 0;
	posptr[0] = 0;
	for(int irow = 0; irow < nbrows; irow++)
		for(int j = browptr[irow]; j < browptr[irow+1]; j++) {
			const int col = bcolind[j];
			const int lim = (irow > col) ? col : irow;      /* :46-48 lower, :62-63 upper */
			for(int k = browptr[irow]; k < browptr[irow+1] && bcolind[k] < lim; k++) {
				const int kc = bcolind[k];
				const int ipos = inner_search(bcolind, diagind[kc], browptr[kc+1], col);
				if(ipos > -1) {
					if(lowerp) { lowerp[total] = k; upperp[total] = ipos; }
					total++;
				}
			}
			posptr[j+1] = (int)total;                       /* inclusive scan, :87-88 */
		}
	return total;
}

/* ---------- T5 levels ---------- */

static int has_col(const int *browptr, const int *bcolind, int row, int col)
{
	return inner_search(bcolind, browptr[row], browptr[row+1], col) >= 0;
}

/* src/levelschedule.cpp:12-71.  With sorted columns and a structurally symmetric pattern the
 * std::list bookkeeping reduces to: after all rows < s are in finished levels, the front of row r's
 * dependency list is its smallest column >= s; the level starting at s extends over consecutive
 * rows r that have no column in [s, r). */
int orc_compute_levels(int nbrows, const int *browptr, const int *bcolind, int *levels)
{
	/* "(jnode must be found because the sparsity structure is symmetric)" :55-57 */
	for(int i = 0; i < nbrows; i++)
		for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
			if(!has_col(browptr, bcolind, bcolind[jj], i))
				return -1;

	int nl = 0;
	levels[nl++] = 0;
	int inode = 0;
	while(inode < nbrows) {
		const int s = inode;
		while(inode < nbrows) {
			int dep = 0;
			for(int jj = browptr[inode]; jj < browptr[inode+1]; jj++) {
				const int c = bcolind[jj];
				if(c >= s && c < inode) { dep = 1; break; }
			}
			if(dep) break;
			inode++;
		}
		levels[nl++] = inode;
	}
	return nl;
}

int orc_dag_levels(int nbrows, const int *browptr, const int *bcolind, const int *diagind,
                   int *level_of_row)
{
	int nlev = 0;
	for(int i = 0; i < nbrows; i++) {
		int lv = 0;
		for(int jj = browptr[i]; jj < diagind[i]; jj++) {
			const int l = level_of_row[bcolind[jj]] + 1;
			if(l > lv) lv = l;
		}
		level_of_row[i] = lv;
		if(lv+1 > nlev) nlev = lv+1;
	}
	return nlev;
}

/* ---------- T6 scaling ---------- */

/* src/rawsrmatrixutils.cpp:343-350 */
void orc_scaling_vector(int bs, int nbrows, const double *vals, const int *diagind, double *scale)
{
	for(int i = 0; i < nbrows; i++)
		for(int j = 0; j < bs; j++)
			scale[(size_t)i*bs+j] = 1.0/sqrt(vals[(size_t)diagind[i]*bs*bs + j*bs + j]);
}

/* ---------- K4 init ---------- */

void orc_ilu0_init(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *scale, int fact_init,
                   double *ilu)
{
	const int bs2 = bs*bs;
	const size_t nn = (size_t)browptr[nbrows]*bs2;

	if(fact_init == ORC_INIT_F_NONE) return;

	if(bs == 1) {
		/* src/async_ilu_factor.cpp:47-58 : INIT_F_ZERO falls through into INIT_F_ORIGINAL */
		if(fact_init == ORC_INIT_F_ZERO || fact_init == ORC_INIT_F_ORIGINAL) {
			/* :136-151 */
			for(int i = 0; i < nbrows; i++)
				for(int j = browptr[i]; j < browptr[i+1]; j++)
					ilu[j] = scale ? scale[i]*vals[j]*scale[bcolind[j]] : vals[j];
		}
		else {
			/* :110-133.  NB the reference indexes `scale` with diagind[col] (a position in the
			 * non-zero array, out of bounds for a length-nbrows vector); the intended value
			 * scale[col] is used here.  Unscaled branch is as written. */
			for(int i = 0; i < nbrows; i++) {
				for(int j = browptr[i]; j < browptr[i+1]; j++)
					ilu[j] = scale ? scale[i]*vals[j]*scale[bcolind[j]] : vals[j];
				for(int j = browptr[i]; j < diagind[i]; j++) {
					const int c = bcolind[j];
					if(scale)
						ilu[j] *= 1.0/(vals[diagind[c]]*scale[c]*scale[c]);
					else
						ilu[j] *= 1.0/vals[diagind[c]];
				}
			}
		}
		return;
	}

	/* block: src/async_blockilu_factor.cpp:63-94 */
	if(fact_init == ORC_INIT_F_ZERO) {
		for(size_t i = 0; i < nn; i++) ilu[i] = 0;
		return;
	}
	for(int i = 0; i < nbrows; i++)
		for(int jj = browptr[i]; jj < browptr[i+1]; jj++) {
			memcpy(ilu + (size_t)jj*bs2, vals + (size_t)jj*bs2, bs2*sizeof(double));
			if(scale) scale_block(bs, rm, scale, i, bcolind[jj], ilu + (size_t)jj*bs2);
		}
	if(fact_init == ORC_INIT_F_SGS) {
		/* :207-254 : L' = L D^-1 with D the (scaled) diagonal blocks */
		double *dinv = (double*)malloc((size_t)nbrows*bs2*sizeof(double));
		for(int i = 0; i < nbrows; i++)
			orc_block_inverse(bs, rm, ilu + (size_t)diagind[i]*bs2, dinv + (size_t)i*bs2);
		for(int i = 0; i < nbrows; i++)
			for(int jj = browptr[i]; jj < diagind[i]; jj++) {
				double t[ORC_MAX_BS*ORC_MAX_BS];
				blk_mul(bs, rm, ilu + (size_t)jj*bs2, dinv + (size_t)bcolind[jj]*bs2, t);
				memcpy(ilu + (size_t)jj*bs2, t, bs2*sizeof(double));
			}
		free(dinv);
	}
}

/* ---------- K1/K2 factor sweeps ---------- */

/* one entry of the ILU(0) fixed-point map, reading from `src`, result in `out` (bs2 doubles).
 * scalar: src/kernels/kernels_ilu0_factorize.hpp:26-52; block: :77-97 */
static void ilu0_entry(int bs, int rm, int irow, int jpos, const int *bcolind, const double *vals,
                       const int *diagind, const int *posptr, const int *lowerp, const int *upperp,
                       const double *scale, const double *src, double *out)
{
	const int bs2 = bs*bs;
	const int col = bcolind[jpos];
	if(bs == 1) {
		double sum = vals[jpos];
		if(scale) { sum *= scale[irow]; sum *= scale[col]; }
		for(int k = posptr[jpos]; k < posptr[jpos+1]; k++)
			sum -= src[lowerp[k]]*src[upperp[k]];
		if(irow > col)
			sum = sum / src[diagind[col]];
		out[0] = sum;
		return;
	}
	double sum[ORC_MAX_BS*ORC_MAX_BS];
	memcpy(sum, vals + (size_t)jpos*bs2, bs2*sizeof(double));
	if(scale) scale_block(bs, rm, scale, irow, col, sum);
	for(int k = posptr[jpos]; k < posptr[jpos+1]; k++) {
		double t[ORC_MAX_BS*ORC_MAX_BS];
		blk_mul(bs, rm, src + (size_t)lowerp[k]*bs2, src + (size_t)upperp[k]*bs2, t);
		for(int e = 0; e < bs2; e++) sum[e] -= t[e];
	}
	if(irow > col) {
		double dinv[ORC_MAX_BS*ORC_MAX_BS];
		orc_block_inverse(bs, rm, src + (size_t)diagind[col]*bs2, dinv);
		blk_mul(bs, rm, sum, dinv, out);
	}
	else
		memcpy(out, sum, bs2*sizeof(double));
}

/* src/async_ilu_factor.cpp:154-177, src/async_blockilu_factor.cpp:187-204 with one thread */
void orc_ilu0_sweeps(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                     const double *vals, const int *diagind, const int *posptr, const int *lowerp,
                     const int *upperp, const double *scale, int nsweeps, double *ilu)
{
	const int bs2 = bs*bs;
	for(int isweep = 0; isweep < nsweeps; isweep++)
		for(int irow = 0; irow < nbrows; irow++)
			for(int j = browptr[irow]; j < browptr[irow+1]; j++) {
				double out[ORC_MAX_BS*ORC_MAX_BS];
				ilu0_entry(bs, rm, irow, j, bcolind, vals, diagind, posptr, lowerp, upperp, scale,
				           ilu, out);
				memcpy(ilu + (size_t)j*bs2, out, bs2*sizeof(double));
			}
}

void orc_ilu0_sweep_synchronous(int bs, int rm, int nbrows, const int *browptr,
                                const int *bcolind, const double *vals, const int *diagind,
                                const int *posptr, const int *lowerp, const int *upperp,
                                const double *scale, const double *ilu_old, double *ilu_new)
{
	const int bs2 = bs*bs;
	for(int irow = 0; irow < nbrows; irow++)
		for(int j = browptr[irow]; j < browptr[irow+1]; j++)
			ilu0_entry(bs, rm, irow, j, bcolind, vals, diagind, posptr, lowerp, upperp, scale,
			           ilu_old, ilu_new + (size_t)j*bs2);
}

/* src/async_blockilu_factor.cpp:144-146 */
void orc_ilu0_invert_diag(int bs, int rm, int nbrows, const int *diagind, double *ilu)
{
	if(bs == 1) return;
	const int bs2 = bs*bs;
	for(int i = 0; i < nbrows; i++) {
		double t[ORC_MAX_BS*ORC_MAX_BS];
		orc_block_inverse(bs, rm, ilu + (size_t)diagind[i]*bs2, t);
		memcpy(ilu + (size_t)diagind[i]*bs2, t, bs2*sizeof(double));
	}
}

/* ---------- K10 residual ---------- */

/* src/async_ilu_factor.cpp:180-217; src/async_blockilu_factor.cpp:257-297 */
double orc_ilu0_nonlinear_res(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                              const double *vals, const int *diagind, const int *posptr,
                              const int *lowerp, const int *upperp, const double *scale,
                              const double *ilu)
{
	const int bs2 = bs*bs;
	double resnorm = 0;
	for(int irow = 0; irow < nbrows; irow++)
		for(int j = browptr[irow]; j < browptr[irow+1]; j++) {
			const int col = bcolind[j];
			if(bs == 1) {
				double sum = vals[j];
				if(scale) { sum *= scale[irow]; sum *= scale[col]; }
				for(int k = posptr[j]; k < posptr[j+1]; k++)
					sum -= ilu[lowerp[k]]*ilu[upperp[k]];
				if(irow > col) sum -= ilu[j]*ilu[diagind[col]];
				else sum -= ilu[j];
				resnorm += fabs(sum);
				continue;
			}
			double sum[ORC_MAX_BS*ORC_MAX_BS], t[ORC_MAX_BS*ORC_MAX_BS];
			memcpy(sum, vals + (size_t)j*bs2, bs2*sizeof(double));
			if(scale) scale_block(bs, rm, scale, irow, col, sum);
			for(int k = posptr[j]; k < posptr[j+1]; k++) {
				blk_mul(bs, rm, ilu + (size_t)lowerp[k]*bs2, ilu + (size_t)upperp[k]*bs2, t);
				for(int e = 0; e < bs2; e++) sum[e] -= t[e];
			}
			if(irow > col) {
				blk_mul(bs, rm, ilu + (size_t)j*bs2, ilu + (size_t)diagind[col]*bs2, t);
				for(int e = 0; e < bs2; e++) sum[e] -= t[e];
			}
			else
				for(int e = 0; e < bs2; e++) sum[e] -= ilu[(size_t)j*bs2+e];
			double bn = 0;
			for(int e = 0; e < bs2; e++) bn += fabs(sum[e]);
			resnorm += bn;
		}
	return resnorm;
}

double orc_matrix_abs_sum(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                          const double *vals, const double *scale)
{
	const int bs2 = bs*bs;
	double s = 0;
	for(int irow = 0; irow < nbrows; irow++)
		for(int j = browptr[irow]; j < browptr[irow+1]; j++) {
			double blk[ORC_MAX_BS*ORC_MAX_BS];
			memcpy(blk, vals + (size_t)j*bs2, bs2*sizeof(double));
			if(scale) scale_block(bs, rm, scale, irow, bcolind[j], blk);
			for(int e = 0; e < bs2; e++) s += fabs(blk[e]);
		}
	return s;
}

/* ---------- K5 ILU apply ---------- */

int orc_ilu0_apply(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                   const int *diagind, const double *ilu, const double *scale, int napplysweeps,
                   int apply_init, const double *r, double *z, double *y)
{
	const int bs2 = bs*bs;
	const size_t n = (size_t)nbrows*bs;

	/* z := S r  (src/solverops_ilu0.cpp:76-87, :248-257) */
	for(size_t i = 0; i < n; i++) z[i] = scale ? scale[i]*r[i] : r[i];

	/* :89-100, :259-269 */
	if(apply_init == ORC_INIT_A_JACOBI || apply_init == ORC_INIT_A_ZERO)
		for(size_t i = 0; i < n; i++) y[i] = 0;

	/* L y = S r : src/kernels/kernels_ilu_apply.hpp:15-27, :54-67 */
	for(int isweep = 0; isweep < napplysweeps; isweep++)
		for(int i = 0; i < nbrows; i++) {
			if(bs == 1) {
				double inter = 0;
				for(int jj = browptr[i]; jj < diagind[i]; jj++)
					inter += ilu[jj]*y[bcolind[jj]];
				y[i] = z[i] - inter;
			} else {
				double inter[ORC_MAX_BS];
				for(int q = 0; q < bs; q++) inter[q] = 0;
				for(int jj = browptr[i]; jj < diagind[i]; jj++)
					blk_gemv_acc(bs, rm, ilu + (size_t)jj*bs2, y + (size_t)bcolind[jj]*bs, inter);
				for(int q = 0; q < bs; q++) y[(size_t)i*bs+q] = z[(size_t)i*bs+q] - inter[q];
			}
		}

	/* :112-128, :285-300 */
	if(apply_init == ORC_INIT_A_JACOBI)
		for(size_t i = 0; i < n; i++) z[i] = y[i];
	else if(apply_init == ORC_INIT_A_ZERO)
		for(size_t i = 0; i < n; i++) z[i] = 0;
	else
		return 1;

	/* U z = y : src/kernels/kernels_ilu_apply.hpp:30-42 (scalar, divides by the diagonal entry,
	 * src/solverops_ilu0.cpp:311-312), :79-94 (block, multiplies by the pre-inverted diagonal) */
	for(int isweep = 0; isweep < napplysweeps; isweep++)
		for(int i = nbrows-1; i >= 0; i--) {
			if(bs == 1) {
				double inter = 0;
				for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++)
					inter += ilu[jj]*z[bcolind[jj]];
				z[i] = (1.0/ilu[diagind[i]]) * (y[i] - inter);
			} else {
				double inter[ORC_MAX_BS], t[ORC_MAX_BS], o[ORC_MAX_BS];
				for(int q = 0; q < bs; q++) inter[q] = 0;
				for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++)
					blk_gemv_acc(bs, rm, ilu + (size_t)jj*bs2, z + (size_t)bcolind[jj]*bs, inter);
				for(int q = 0; q < bs; q++) { t[q] = y[(size_t)i*bs+q] - inter[q]; o[q] = 0; }
				blk_gemv_acc(bs, rm, ilu + (size_t)diagind[i]*bs2, t, o);
				for(int q = 0; q < bs; q++) z[(size_t)i*bs+q] = o[q];
			}
		}

	/* :144-147, :316-320 */
	if(scale)
		for(size_t i = 0; i < n; i++) z[i] = z[i]*scale[i];
	return 0;
}

/* ---------- K3 Jacobi ---------- */

void orc_jacobi_setup(int bs, int rm, int nbrows, const double *vals, const int *diagind,
                      double *dblocks)
{
	const int bs2 = bs*bs;
	for(int i = 0; i < nbrows; i++) {
		if(bs == 1) dblocks[i] = 1.0/vals[diagind[i]];                  /* :141-147 */
		else orc_block_inverse(bs, rm, vals + (size_t)diagind[i]*bs2, dblocks + (size_t)i*bs2);  /* :43-45 */
	}
}

void orc_jacobi_apply(int bs, int rm, int nbrows, const double *dblocks, const double *r, double *z)
{
	const int bs2 = bs*bs;
	for(int i = 0; i < nbrows; i++) {
		if(bs == 1) { z[i] = dblocks[i]*r[i]; continue; }
		double o[ORC_MAX_BS];
		for(int q = 0; q < bs; q++) o[q] = 0;
		blk_gemv_acc(bs, rm, dblocks + (size_t)i*bs2, r + (size_t)i*bs, o);
		for(int q = 0; q < bs; q++) z[(size_t)i*bs+q] = o[q];
	}
}

/* ---------- K6 SGS apply ---------- */

void orc_sgs_apply(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *dblocks, int napplysweeps,
                   int apply_init, const double *r, double *z, double *y)
{
	const int bs2 = bs*bs;
	const size_t n = (size_t)nbrows*bs;

	if(apply_init == ORC_INIT_A_JACOBI || apply_init == ORC_INIT_A_ZERO)
		for(size_t i = 0; i < n; i++) y[i] = 0;

	/* forward: y_i = D_i^-1 (r_i - sum_{j<i} A_ij y_j)  (kernels_sgs.hpp:17-30, :47-60) */
	for(int isweep = 0; isweep < napplysweeps; isweep++)
		for(int i = 0; i < nbrows; i++) {
			if(bs == 1) {
				double inter = 0;
				for(int jj = browptr[i]; jj < diagind[i]; jj++)
					inter += vals[jj]*y[bcolind[jj]];
				y[i] = dblocks[i]*(r[i] - inter);
			} else {
				double inter[ORC_MAX_BS], t[ORC_MAX_BS], o[ORC_MAX_BS];
				for(int q = 0; q < bs; q++) inter[q] = 0;
				for(int jj = browptr[i]; jj < diagind[i]; jj++)
					blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, y + (size_t)bcolind[jj]*bs, inter);
				for(int q = 0; q < bs; q++) { t[q] = r[(size_t)i*bs+q] - inter[q]; o[q] = 0; }
				blk_gemv_acc(bs, rm, dblocks + (size_t)i*bs2, t, o);
				for(int q = 0; q < bs; q++) y[(size_t)i*bs+q] = o[q];
			}
		}

	if(apply_init == ORC_INIT_A_JACOBI)
		for(size_t i = 0; i < n; i++) z[i] = y[i];
	else if(apply_init == ORC_INIT_A_ZERO)
		for(size_t i = 0; i < n; i++) z[i] = 0;

	/* backward: z_i = y_i - D_i^-1 sum_{j>i} A_ij z_j  (kernels_sgs.hpp:33-44, :63-76) */
	for(int isweep = 0; isweep < napplysweeps; isweep++)
		for(int i = nbrows-1; i >= 0; i--) {
			if(bs == 1) {
				double inter = 0;
				for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++)
					inter += vals[jj]*z[bcolind[jj]];
				z[i] = y[i] - dblocks[i]*inter;
			} else {
				double inter[ORC_MAX_BS], o[ORC_MAX_BS];
				for(int q = 0; q < bs; q++) { inter[q] = 0; o[q] = 0; }
				for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++)
					blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, z + (size_t)bcolind[jj]*bs, inter);
				blk_gemv_acc(bs, rm, dblocks + (size_t)i*bs2, inter, o);
				for(int q = 0; q < bs; q++) z[(size_t)i*bs+q] = y[(size_t)i*bs+q] - o[q];
			}
		}
}

/* ---------- K7 relaxation ---------- */

/* src/kernels/kernels_relaxation.hpp:17-54 : x_i <- D_i^-1 (b_i - sum_{j != i} A_ij x_j) */
static void relax_row(int bs, int rm, int i, const int *browptr, const int *bcolind,
                      const double *vals, const int *diagind, const double *dblocks,
                      const double *b, const double *xsrc, double *xdst)
{
	const int bs2 = bs*bs;
	if(bs == 1) {
		double inter = 0;
		for(int jj = browptr[i]; jj < diagind[i]; jj++) inter += vals[jj]*xsrc[bcolind[jj]];
		for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++) inter += vals[jj]*xsrc[bcolind[jj]];
		xdst[i] = dblocks[i]*(b[i] - inter);
		return;
	}
	double inter[ORC_MAX_BS], t[ORC_MAX_BS], o[ORC_MAX_BS];
	for(int q = 0; q < bs; q++) inter[q] = 0;
	for(int jj = browptr[i]; jj < diagind[i]; jj++)
		blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, xsrc + (size_t)bcolind[jj]*bs, inter);
	for(int jj = diagind[i]+1; jj < browptr[i+1]; jj++)
		blk_gemv_acc(bs, rm, vals + (size_t)jj*bs2, xsrc + (size_t)bcolind[jj]*bs, inter);
	for(int q = 0; q < bs; q++) { t[q] = b[(size_t)i*bs+q] - inter[q]; o[q] = 0; }
	blk_gemv_acc(bs, rm, dblocks + (size_t)i*bs2, t, o);
	for(int q = 0; q < bs; q++) xdst[(size_t)i*bs+q] = o[q];
}

/* src/solverops_sgs.cpp:86-116, :180-203 */
void orc_sgs_relax(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                   const double *vals, const int *diagind, const double *dblocks, int maxits,
                   const double *b, double *x)
{
	for(int step = 0; step < maxits; step++) {
		for(int i = 0; i < nbrows; i++)
			relax_row(bs, rm, i, browptr, bcolind, vals, diagind, dblocks, b, x, x);
		for(int i = nbrows-1; i >= 0; i--)
			relax_row(bs, rm, i, browptr, bcolind, vals, diagind, dblocks, b, x, x);
	}
}

/* src/relaxation_chaotic.cpp:22-123 */
void orc_gs_relax(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                  const double *vals, const int *diagind, const double *dblocks, int nsweeps,
                  const double *b, double *x)
{
	for(int step = 0; step < nsweeps; step++)
		for(int i = 0; i < nbrows; i++)
			relax_row(bs, rm, i, browptr, bcolind, vals, diagind, dblocks, b, x, x);
}

/* src/solverops_jacobi.cpp:66-121, :174-220 with ctol == false */
void orc_jacobi_relax(int bs, int rm, int nbrows, const int *browptr, const int *bcolind,
                      const double *vals, const int *diagind, const double *dblocks, int maxits,
                      const double *b, double *x, double *xtemp)
{
	const size_t n = (size_t)nbrows*bs;
	for(int step = 0; step < maxits; step++) {
		for(int i = 0; i < nbrows; i++)
			relax_row(bs, rm, i, browptr, bcolind, vals, diagind, dblocks, b, x, xtemp);
		for(size_t i = 0; i < n; i++) x[i] = xtemp[i];
	}
}

/* ---------- diagnostics ---------- */

/* src/matrix_properties.cpp:11-77 */
void orc_diagonal_dominance(int bs, int rm, int nbrows, const int *browptr, const int *diagind,
                            const double *vals, double out[4])
{
	const int bs2 = bs*bs;
	double uddavg = 0, uddmin = 1e30, lddavg = 0, lddmin = 1e30;
	for(int irow = 0; irow < nbrows; irow++) {
		double rowddu[ORC_MAX_BS], rowddl[ORC_MAX_BS];
		for(int i = 0; i < bs; i++) { rowddl[i] = 0; rowddu[i] = 0; }
		const int diagp = diagind[irow];
		const double *db = vals + (size_t)diagp*bs2;
		for(int i = 0; i < bs; i++)
			for(int j = 0; j < bs; j++)
				if(i != j) rowddu[i] += fabs(db[BIDX(rm,bs,i,j)]);
		for(int jj = diagp+1; jj < browptr[irow+1]; jj++)
			for(int i = 0; i < bs; i++)
				for(int j = 0; j < bs; j++)
					rowddu[i] += fabs(vals[(size_t)jj*bs2 + BIDX(rm,bs,i,j)]);
		for(int jj = browptr[irow]; jj < diagp; jj++)
			for(int i = 0; i < bs; i++)
				for(int j = 0; j < bs; j++)
					rowddl[i] += fabs(vals[(size_t)jj*bs2 + BIDX(rm,bs,i,j)]);
		for(int i = 0; i < bs; i++) {
			rowddl[i] = 1.0 - rowddl[i];
			rowddu[i] = 1.0 - rowddu[i]/fabs(db[BIDX(rm,bs,i,i)]);
		}
		for(int i = 0; i < bs; i++) {
			if(uddmin > rowddu[i]) uddmin = rowddu[i];
			if(lddmin > rowddl[i]) lddmin = rowddl[i];
			lddavg += rowddl[i];
			uddavg += rowddu[i];
		}
	}
	out[0] = lddavg/((double)nbrows*bs);
	out[1] = lddmin;
	out[2] = uddavg/((double)nbrows*bs);
	out[3] = uddmin;
}

/* ---------- front end ---------- */

typedef struct { int r, c; double v; long long k; } orc_entry;

static int cmp_entry(const void *a, const void *b)
{
	const orc_entry *x = (const orc_entry*)a, *y = (const orc_entry*)b;
	if(x->r != y->r) return x->r < y->r ? -1 : 1;
	if(x->c != y->c) return x->c < y->c ? -1 : 1;
	return x->k < y->k ? -1 : (x->k > y->k);            /* input order among duplicates */
}

/* src/coomatrix.cpp:222-403 */
long long orc_coo_convert(int nrows, long long nnz, const int *rowind, const int *colind,
                          const double *val, int bs, int rm, int *browptr, int *bcolind,
                          int *diagind, double *vals)
{
	orc_entry *e = (orc_entry*)malloc((size_t)(nnz > 0 ? nnz : 1)*sizeof(orc_entry));
	for(long long i = 0; i < nnz; i++) { e[i].r = rowind[i]; e[i].c = colind[i]; e[i].v = val[i]; e[i].k = i; }
	qsort(e, (size_t)nnz, sizeof(orc_entry), cmp_entry);
	int *rowptr = (int*)calloc((size_t)nrows + 1, sizeof(int));
	for(long long i = 0; i < nnz; i++) rowptr[e[i].r + 1]++;
	for(int i = 0; i < nrows; i++) rowptr[i+1] += rowptr[i];

	long long bnnz = 0;
	if(bs == 1) {
		/* convertToCSR :262-297 */
		bnnz = nnz;
		if(bcolind) {
			for(int i = 0; i <= nrows; i++) browptr[i] = rowptr[i];
			for(int i = 0; i < nrows; i++) diagind[i] = -1;
			for(long long j = 0; j < nnz; j++) {
				bcolind[j] = e[j].c; vals[j] = e[j].v;
				if(e[j].c == e[j].r) diagind[e[j].r] = (int)j;
			}
		}
	} else {
		/* convertToBSR :299-403 */
		const int nbrows = nrows/bs, bs2 = bs*bs;
		int *bptr = (int*)calloc((size_t)nbrows + 1, sizeof(int));
		int *bcol = (int*)malloc((size_t)(nnz > 0 ? nnz : 1)*sizeof(int));
		int *bdiag = (int*)malloc((size_t)(nbrows > 0 ? nbrows : 1)*sizeof(int));
		char *tally = (char*)calloc((size_t)nbrows + 1, 1);
		for(int i = 0; i < nbrows; i++) bdiag[i] = -1;
		for(int irow = 0; irow < nrows; irow++) {
			const int brow = irow/bs;
			for(int j = rowptr[irow]; j < rowptr[irow+1]; j++) {
				const int bc = e[j].c/bs;
				if(!tally[brow]) { bptr[brow] = (int)bnnz; tally[brow] = 1; }
				long long it = bptr[brow];
				while(it < bnnz && bcol[it] != bc) it++;
				if(it == bnnz) {
					bcol[bnnz] = bc;
					if(bc == brow) bdiag[brow] = (int)bnnz;
					bnnz++;
				}
			}
		}
		bptr[nbrows] = (int)bnnz;
		for(int i = nbrows-1; i > 0; i--)
			if(bptr[i] == 0) bptr[i] = bptr[i+1];
		if(bcolind) {
			for(int i = 0; i <= nbrows; i++) browptr[i] = bptr[i];
			for(int i = 0; i < nbrows; i++) diagind[i] = bdiag[i];
			for(long long i = 0; i < bnnz; i++) bcolind[i] = bcol[i];
			for(long long i = 0; i < bnnz*bs2; i++) vals[i] = 0;
			for(int irow = 0; irow < nrows; irow++) {
				const int brow = irow/bs;
				for(int j = rowptr[irow]; j < rowptr[irow+1]; j++) {
					const int c = e[j].c, bc = c/bs;
					const int off = rm ? (irow - brow*bs)*bs + c - bc*bs : (c - bc*bs)*bs + irow - brow*bs;
					int p = bptr[brow];
					while(p < bptr[brow+1] && bcol[p] != bc) p++;
					vals[(size_t)p*bs2 + off] = e[j].v;
				}
			}
		}
		free(bptr); free(bcol); free(bdiag); free(tally);
	}
	free(e); free(rowptr);
	return bnnz;
}

/* internal::sortBlockInnerDimension (src/helper_algorithms.hpp): sort a row's blocks by column */
static void sort_row_blocks(int bs2, int n, int *col, double *v)
{
	double *tmp = (double*)malloc((size_t)bs2*sizeof(double));
	for(int i = 1; i < n; i++) {
		const int c = col[i];
		memcpy(tmp, v + (size_t)i*bs2, (size_t)bs2*sizeof(double));
		int j = i - 1;
		while(j >= 0 && col[j] > c) {
			col[j+1] = col[j];
			memcpy(v + (size_t)(j+1)*bs2, v + (size_t)j*bs2, (size_t)bs2*sizeof(double));
			j--;
		}
		col[j+1] = c;
		memcpy(v + (size_t)(j+1)*bs2, tmp, (size_t)bs2*sizeof(double));
	}
	free(tmp);
}

/* src/reorderingscaling.cpp:77-205 */
void orc_reorder_matrix(int bs, int nbrows, int *browptr, int *bcolind, double *vals,
                        const int *rord, const int *cord, int inverse)
{
	const int bs2 = bs*bs;
	const int nnzb = browptr[nbrows];
	if(rord) {
		int *tptr = (int*)malloc(((size_t)nbrows + 1)*sizeof(int));
		int *tcol = (int*)malloc((size_t)(nnzb > 0 ? nnzb : 1)*sizeof(int));
		double *tval = (double*)malloc((size_t)(nnzb > 0 ? nnzb : 1)*bs2*sizeof(double));
		if(!inverse) {
			/* new row i is old row rord[i] (:83-115) */
			int pos = 0;
			for(int i = 0; i < nbrows; i++) {
				tptr[i] = pos;
				for(int jj = browptr[rord[i]]; jj < browptr[rord[i]+1]; jj++, pos++) {
					tcol[pos] = bcolind[jj];
					memcpy(tval + (size_t)pos*bs2, vals + (size_t)jj*bs2, (size_t)bs2*sizeof(double));
				}
			}
			tptr[nbrows] = pos;
		} else {
			/* old row i moves to row rord[i] (:144-178) */
			int *len = (int*)calloc((size_t)nbrows + 1, sizeof(int));
			for(int i = 0; i < nbrows; i++) len[rord[i]] = browptr[i+1] - browptr[i];
			tptr[0] = 0;
			for(int i = 0; i < nbrows; i++) tptr[i+1] = tptr[i] + len[i];
			for(int i = 0; i < nbrows; i++) {
				int pos = tptr[rord[i]];
				for(int jj = browptr[i]; jj < browptr[i+1]; jj++, pos++) {
					tcol[pos] = bcolind[jj];
					memcpy(tval + (size_t)pos*bs2, vals + (size_t)jj*bs2, (size_t)bs2*sizeof(double));
				}
			}
			free(len);
		}
		memcpy(browptr, tptr, ((size_t)nbrows + 1)*sizeof(int));
		memcpy(bcolind, tcol, (size_t)nnzb*sizeof(int));
		memcpy(vals, tval, (size_t)nnzb*bs2*sizeof(double));
		free(tptr); free(tcol); free(tval);
	}
	if(cord) {
		/* forward renames with the inverse permutation, inverse with cord itself (:117-139,181-203) */
		int *map = (int*)malloc((size_t)(nbrows > 0 ? nbrows : 1)*sizeof(int));
		if(!inverse) for(int i = 0; i < nbrows; i++) map[cord[i]] = i;
		else for(int i = 0; i < nbrows; i++) map[i] = cord[i];
		for(int i = 0; i < nbrows; i++) {
			for(int jj = browptr[i]; jj < browptr[i+1]; jj++) bcolind[jj] = map[bcolind[jj]];
			sort_row_blocks(bs2, browptr[i+1] - browptr[i], bcolind + browptr[i],
			                vals + (size_t)browptr[i]*bs2);
		}
		free(map);
	}
}

/* src/reorderingscaling.cpp:211-266 */
void orc_reorder_vector(int bs, int n, const int *ord, int inverse, double *vec)
{
	if(!ord || n <= 0) return;
	double *tv = (double*)malloc((size_t)n*bs*sizeof(double));
	memcpy(tv, vec, (size_t)n*bs*sizeof(double));
	for(int i = 0; i < n; i++)
		for(int k = 0; k < bs; k++) {
			if(!inverse) vec[(size_t)i*bs + k] = tv[(size_t)ord[i]*bs + k];
			else vec[(size_t)ord[i]*bs + k] = tv[(size_t)i*bs + k];
		}
	free(tv);
}

/* src/reorderingscaling.cpp:282-337 */
void orc_scale_matrix(int bs, int nbrows, const int *browptr, const int *bcolind, double *vals,
                      const double *rowscale, const double *colscale, int inverse)
{
	const int bs2 = bs*bs;
	if(rowscale)
		for(int i = 0; i < nbrows; i++)
			for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
				for(int k = 0; k < bs2; k++) {
					if(!inverse) vals[(size_t)jj*bs2 + k] *= rowscale[i];
					else vals[(size_t)jj*bs2 + k] /= rowscale[i];
				}
	if(colscale)
		for(int i = 0; i < nbrows; i++)
			for(int jj = browptr[i]; jj < browptr[i+1]; jj++)
				for(int k = 0; k < bs2; k++) {
					if(!inverse) vals[(size_t)jj*bs2 + k] *= colscale[bcolind[jj]];
					else vals[(size_t)jj*bs2 + k] /= colscale[bcolind[jj]];
				}
}

/* src/reorderingscaling.cpp:340-368 */
void orc_scale_vector(int bs, int n, const double *scale, int inverse, double *vec)
{
	if(!scale) return;
	for(int i = 0; i < n; i++)
		for(int k = 0; k < bs; k++) {
			if(!inverse) vec[(size_t)i*bs + k] *= scale[i];
			else vec[(size_t)i*bs + k] /= scale[i];
		}
}
