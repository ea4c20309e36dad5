/** \file blas1.cu
 * \brief Vector kernels of the Krylov test drivers (K11): axpby, axpbypcz, fused multi-dot.
 *
 * Replaces the OpenMP loops axpby / axpbypcz / dot / vecassign of tests/solvers.cpp:20-60 of the
 * reference.  Dots are fused (several pairs per pass, one deterministic two-stage reduction) and
 * their results stay on the device until the driver needs them on the host.
 */
#include "common.cuh"

namespace b200 {

__global__ void axpby_kernel(const long long n, const double p, double *z, const double q,
                             const double *x)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) z[i] = p*z[i] + q*x[i];
}

__global__ void axpbypcz_kernel(const long long n, const double p, double *z, const double q,
                                const double *x, const double r, const double *y)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) z[i] = p*z[i] + q*x[i] + r*y[i];
}

void launch_axpby(long long n, double p, double *z, double q, const double *x, cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	if(n == 0) return;
	axpby_kernel<<<div_up(n,256),256,0,st>>>(n, p, z, q, x);
	B200_LAUNCHED();
}

void launch_axpbypcz(long long n, double p, double *z, double q, const double *x, double r,
                     const double *y, cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	if(n == 0) return;
	axpbypcz_kernel<<<div_up(n,256),256,0,st>>>(n, p, z, q, x, r, y);
	B200_LAUNCHED();
}

struct DotPtrs { const double *a[MAX_DOTS]; const double *b[MAX_DOTS]; };

template <int ND>
__global__ void __launch_bounds__(256)
multi_dot_stage1(const long long n, const DotPtrs p, double *__restrict__ partial)
{
	double acc[ND];
#pragma unroll
	for(int d = 0; d < ND; d++) acc[d] = 0;
	for(long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n;
	    i += (long long)gridDim.x*blockDim.x)
	{
#pragma unroll
		for(int d = 0; d < ND; d++) acc[d] = fma(p.a[d][i], p.b[d][i], acc[d]);
	}
	__shared__ double sm[ND][8];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for(int d = 0; d < ND; d++) {
		double v = acc[d];
#pragma unroll
		for(int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
		if(lane == 0) sm[d][w] = v;
	}
	__syncthreads();
	if(threadIdx.x < ND) {
		double t = 0;
		for(int i = 0; i < 8; i++) t += sm[threadIdx.x][i];
		partial[threadIdx.x*DOT_BLOCKS + blockIdx.x] = t;
	}
}

/// The Gram-Schmidt case: every product is against the same vector w.  w is read once per element
/// (the general kernel would request it ND times), two elements per thread and iteration as one
/// 16-byte load per vector.
template <int ND>
__global__ void __launch_bounds__(256)
multi_dot_common_stage1(const long long n, const DotPtrs p, double *__restrict__ partial)
{
	double acc[ND];
#pragma unroll
	for(int d = 0; d < ND; d++) acc[d] = 0;
	const double *__restrict__ w = p.b[0];
	const long long n2 = n >> 1;
	for(long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n2;
	    i += (long long)gridDim.x*blockDim.x)
	{
		const double2 wv = __ldg(reinterpret_cast<const double2*>(w) + i);
		double2 av[ND];
#pragma unroll
		for(int d = 0; d < ND; d++) av[d] = __ldg(reinterpret_cast<const double2*>(p.a[d]) + i);
#pragma unroll
		for(int d = 0; d < ND; d++) acc[d] = fma(av[d].y, wv.y, fma(av[d].x, wv.x, acc[d]));
	}
	if((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
		for(int d = 0; d < ND; d++) acc[d] = fma(p.a[d][n-1], w[n-1], acc[d]);
	}
	__shared__ double sm[ND][8];
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
	for(int d = 0; d < ND; d++) {
		double v = acc[d];
#pragma unroll
		for(int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
		if(lane == 0) sm[d][wp] = v;
	}
	__syncthreads();
	if(threadIdx.x < ND) {
		double t = 0;
		for(int i = 0; i < 8; i++) t += sm[threadIdx.x][i];
		partial[threadIdx.x*DOT_BLOCKS + blockIdx.x] = t;
	}
}

__global__ void __launch_bounds__(256)
multi_dot_stage2(const int nd, const int nblocks, const double *__restrict__ partial,
                 double *__restrict__ out)
{
	__shared__ double sm[8];
	for(int d = 0; d < nd; d++) {
		double v = 0;
		for(int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partial[d*DOT_BLOCKS + i];
#pragma unroll
		for(int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
		if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
		__syncthreads();
		if(threadIdx.x == 0) {
			double t = 0;
			for(int i = 0; i < 8; i++) t += sm[i];
			out[d] = t;
		}
		__syncthreads();
	}
}

void launch_multi_dot(long long n, int nd, const double *const *a, const double *const *b,
                      double *d_partial, double *d_out, cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	if(nd < 1 || nd > MAX_DOTS) throw Error("multi_dot: bad count");
	DotPtrs p;
	for(int d = 0; d < MAX_DOTS; d++) { p.a[d] = a[d < nd ? d : 0]; p.b[d] = b[d < nd ? d : 0]; }
	const int grid = (int)std::max<long long>(1, std::min<long long>(DOT_BLOCKS, div_up(n, 256)));
	// all products against one 16-byte aligned vector: the Gram-Schmidt form
	bool common = nd > 1 && (reinterpret_cast<size_t>(b[0]) & 15) == 0;
	for(int d = 0; d < nd && common; d++)
		common = (b[d] == b[0]) && (reinterpret_cast<size_t>(a[d]) & 15) == 0;
#define B200_DOT_CASE(K) \
	case K: if(common) multi_dot_common_stage1<K><<<grid,256,0,st>>>(n, p, d_partial); \
	        else multi_dot_stage1<K><<<grid,256,0,st>>>(n, p, d_partial); break;
	switch(nd) {
	B200_DOT_CASE(1) B200_DOT_CASE(2) B200_DOT_CASE(3) B200_DOT_CASE(4) B200_DOT_CASE(5)
	B200_DOT_CASE(6) B200_DOT_CASE(7) B200_DOT_CASE(8) B200_DOT_CASE(9) B200_DOT_CASE(10)
	B200_DOT_CASE(11) B200_DOT_CASE(12) B200_DOT_CASE(13) B200_DOT_CASE(14) B200_DOT_CASE(15)
	default: if(common) multi_dot_common_stage1<16><<<grid,256,0,st>>>(n, p, d_partial);
	         else multi_dot_stage1<16><<<grid,256,0,st>>>(n, p, d_partial); break;
	}
#undef B200_DOT_CASE
	B200_LAUNCHED();
	multi_dot_stage2<<<1,256,0,st>>>(nd, grid, d_partial, d_out);
	B200_LAUNCHED();
}

struct AxpyPtrs { const double *v[32]; };

__global__ void __launch_bounds__(256)
multi_axpy_kernel(const long long n, const int nv, const AxpyPtrs p, const double *__restrict__ coef,
                  double *y, const double sign)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double s = y[i];
	for(int l = 0; l < nv; l++) s = fma(sign*__ldg(coef + l), p.v[l][i], s);
	y[i] = s;
}

/// out = (y + sign * sum_l coef[l] v_l) * (*scale): the Gram-Schmidt update and the normalisation of
/// the new basis vector in one pass (y itself is not rewritten)
__global__ void __launch_bounds__(256)
multi_axpy_scaled_kernel(const long long n, const int nv, const AxpyPtrs p,
                         const double *__restrict__ coef, const double *__restrict__ y,
                         double *__restrict__ out, const double sign, const double *__restrict__ scale)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double s = y[i];
	for(int l = 0; l < nv; l++) s = fma(sign*__ldg(coef + l), p.v[l][i], s);
	out[i] = s*__ldg(scale);
}

/// After the fused Gram-Schmidt multi-dot of FGMRES iteration j: dots[0..j] = v_i.w, dots[j+1] = w.w.
/// Writes column j of the Hessenberg matrix {h_0..h_j, |w - sum h_i v_i|, cancellation flag} and
/// the reciprocal norm for the normalisation - on the device, so that the host need not wait.
__global__ void fgmres_column_kernel(const int j, const double *__restrict__ dots,
                                     double *__restrict__ hcol, const int hn_slot,
                                     double *__restrict__ inv_hn)
{
	if(threadIdx.x != 0 || blockIdx.x != 0) return;
	double sumsq = 0;
	for(int i = 0; i <= j; i++) { const double h = dots[i]; hcol[i] = h; sumsq = fma(h, h, sumsq); }
	const double ww = dots[j+1], hn2 = ww - sumsq;
	// |w_new|^2 = w.w - sum h_i^2 (V orthonormal); more than four digits lost to cancellation (or
	// a non-finite value) is flagged for the host, which redoes the iteration with an explicit norm
	const bool ok = hn2 > 1e-4*ww;
	const double hn = ok ? sqrt(hn2) : 0.0;
	hcol[hn_slot] = hn;
	hcol[hn_slot + 1] = ok ? 0.0 : 1.0;
	*inv_hn = hn > 0 ? 1.0/hn : 0.0;
}

__global__ void vec_scal_kernel(const long long n, const double alpha, const double *__restrict__ in,
                                double *__restrict__ out)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i < n) out[i] = alpha*in[i];
}

__global__ void __launch_bounds__(256)
update_diffnorm_kernel(const long long n, const double *__restrict__ xtemp, double *x,
                       double *__restrict__ out)
{
	double acc = 0;
	for(long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n;
	    i += (long long)gridDim.x*blockDim.x)
	{
		const double t = xtemp[i], d = t - x[i];
		acc = fma(d, d, acc);
		x[i] = t;
	}
#pragma unroll
	for(int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
	__shared__ double sm[8];
	if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
	__syncthreads();
	if(threadIdx.x == 0) {
		double t = 0;
		for(int i = 0; i < 8; i++) t += sm[i];
		atomicAdd(out, t);
	}
}

void launch_update_diffnorm(long long n, const double *xtemp, double *x, double *d_out, cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	B200_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
	if(n == 0) return;
	const int grid = (int)std::max<long long>(1, std::min<long long>(DOT_BLOCKS, div_up(n, 256)));
	update_diffnorm_kernel<<<grid,256,0,st>>>(n, xtemp, x, d_out);
	B200_LAUNCHED();
}

void launch_vec_scal(long long n, double alpha, const double *in, double *out, cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	if(n == 0) return;
	vec_scal_kernel<<<div_up(n,256),256,0,st>>>(n, alpha, in, out);
	B200_LAUNCHED();
}

void launch_multi_axpy_scaled(long long n, int nv, const double *const *v, const double *d_coef,
                              const double *y, double *out, const double *d_scale, cudaStream_t st,
                              double sign)
{
	ProfScope ps(KC_BLAS1, st);
	if(n == 0) return;
	if(nv > 32) throw Error("multi_axpy: too many vectors");
	AxpyPtrs p;
	for(int l = 0; l < 32; l++) p.v[l] = v[l < nv ? l : 0];
	multi_axpy_scaled_kernel<<<div_up(n,256),256,0,st>>>(n, nv, p, d_coef, y, out, sign, d_scale);
	B200_LAUNCHED();
}

void launch_fgmres_column(int j, const double *d_dots, double *d_hcol, int hn_slot, double *d_inv_hn,
                          cudaStream_t st)
{
	ProfScope ps(KC_BLAS1, st);
	fgmres_column_kernel<<<1,32,0,st>>>(j, d_dots, d_hcol, hn_slot, d_inv_hn);
	B200_LAUNCHED();
}

void launch_multi_axpy(long long n, int nv, const double *const *v, const double *d_coef, double *y,
                       cudaStream_t st, double sign)
{
	ProfScope ps(KC_BLAS1, st);
	if(n == 0 || nv == 0) return;
	if(nv > 32) throw Error("multi_axpy: too many vectors");
	AxpyPtrs p;
	for(int l = 0; l < 32; l++) p.v[l] = v[l < nv ? l : 0];
	multi_axpy_kernel<<<div_up(n,256),256,0,st>>>(n, nv, p, d_coef, y, sign);
	B200_LAUNCHED();
}

}  // namespace b200
