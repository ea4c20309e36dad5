/** \file krylov.cu
 * \brief Device-resident Krylov test drivers: right-preconditioned BiCGSTAB, restarted GCR
 * (== FGMRES in exact arithmetic) and Richardson.
 *
 * Mirrors the operation order of the reference's drivers so that iteration counts are comparable:
 *   RichardsonSolver::solve   tests/solvers.cpp:90-133
 *   BiCGSTAB::solve           tests/solvers.cpp:140-244   (iters = step+1)
 *   GCR::solve                tests/solvers.cpp:252-352   (tests/solvers.hpp:108-110)
 * All vectors stay in HBM; dots are fused where the algorithm allows (t.r and t.t; the GCR
 * orthogonalisation coefficients) and reduced deterministically; only the scalars the recurrences
 * need travel to the host.
 */
#include "common.cuh"
#include <cmath>

namespace b200 {

namespace {

struct Timer {
	cudaEvent_t a = nullptr, b = nullptr;
	cudaStream_t st;
	explicit Timer(cudaStream_t s) : st(s) {
		B200_CUDA(cudaEventCreate(&a));
		B200_CUDA(cudaEventCreate(&b));
	}
	~Timer() { if(a) cudaEventDestroy(a); if(b) cudaEventDestroy(b); }
	void start() { B200_CUDA(cudaEventRecord(a, st)); }
	double stop() {
		B200_CUDA(cudaEventRecord(b, st));
		B200_CUDA(cudaEventSynchronize(b));
		float ms = 0;
		B200_CUDA(cudaEventElapsedTime(&ms, a, b));
		return ms;
	}
};

double dot1(KrylovOps& ops, const double *a, const double *b)
{
	double out;
	const double *aa[1] = {a}, *bb[1] = {b};
	ops.dots(1, aa, bb, &out);
	return out;
}

void richardson(KrylovOps& ops, const double *b, double *x, double tol, int maxiter,
                b200_solve_info *info)
{
	const long long n = ops.n;
	cudaStream_t st = ops.stream;
	DevBuf<double> s, ddu;
	s.alloc(n); ddu.alloc(n);
	B200_CUDA(cudaMemsetAsync(ddu, 0, n*sizeof(double), st));
	const double bnorm = std::sqrt(dot1(ops, b, b));
	double resnorm = 0;
	int step = 0;
	while(step < maxiter) {
		ops.gemv3(-1.0, x, 1.0, b, s);
		resnorm = std::sqrt(dot1(ops, s, s));
		if(resnorm/bnorm < tol) break;
		ops.prec(s, ddu);
		launch_axpby(n, 1.0, x, 1.0, ddu, st);
		step++;
	}
	info->iters = step;
	info->resnorm = resnorm;
	info->bnorm = bnorm;
	info->converged = resnorm/bnorm < tol;
}

void bicgstab(KrylovOps& ops, const double *rhs, double *x, double tol, int maxiter,
              b200_solve_info *info)
{
	const long long n = ops.n;
	cudaStream_t st = ops.stream;
	DevBuf<double> rhat, r, p, v, y, z, t;
	rhat.alloc(n); r.alloc(n); p.alloc(n); v.alloc(n); y.alloc(n); z.alloc(n); t.alloc(n);
	B200_CUDA(cudaMemsetAsync(p, 0, n*sizeof(double), st));
	B200_CUDA(cudaMemsetAsync(v, 0, n*sizeof(double), st));
	B200_CUDA(cudaMemsetAsync(y, 0, n*sizeof(double), st));
	B200_CUDA(cudaMemsetAsync(z, 0, n*sizeof(double), st));

	double resnorm = 100.0;
	int step = 0;
	double omega = 1.0, rho, rhoold = 1.0, alpha = 1.0, beta;

	ops.gemv3(-1.0, x, 1.0, rhs, r);                       // r := rhs - A x
	const double bnorm = std::sqrt(dot1(ops, rhs, rhs));
	B200_CUDA(cudaMemcpyAsync(rhat, r, n*sizeof(double), cudaMemcpyDeviceToDevice, st));

	while(step < maxiter) {
		rho = dot1(ops, rhat, r);
		beta = rho*alpha/(rhoold*omega);
		launch_axpbypcz(n, beta, p, 1.0, r, -beta*omega, v, st);   // p <- r + beta p - beta omega v
		ops.prec(p, y);                                            // y <- Minv p
		ops.spmv(y, v);                                            // v <- A y
		alpha = rho/dot1(ops, rhat, v);
		launch_axpby(n, 1.0, r, -alpha, v, st);                    // s <- r - alpha v (in r)
		ops.prec(r, z);                                            // z <- Minv s
		ops.spmv(z, t);                                            // t <- A z
		{
			double d2[2];
			const double *aa[2] = {t, t}, *bb[2] = {r, t};
			ops.dots(2, aa, bb, d2);
			omega = d2[0]/d2[1];
		}
		launch_axpbypcz(n, 1.0, x, alpha, y, omega, z, st);        // x <- x + alpha y + omega z
		launch_axpby(n, 1.0, r, -omega, t, st);                    // r <- r - omega t
		resnorm = std::sqrt(dot1(ops, r, r));
		if(resnorm/bnorm < tol) break;
		rhoold = rho;
		step++;
	}
	info->iters = step+1;
	info->resnorm = resnorm;
	info->bnorm = bnorm;
	info->converged = resnorm/bnorm < tol;
}

void gcr(KrylovOps& ops, const double *b, double *x, double tol, int maxiter, int nrestart,
         b200_solve_info *info)
{
	const long long n = ops.n;
	cudaStream_t st = ops.stream;
	if(nrestart < 1) throw Error("GCR: restart length must be positive");
	DevBuf<double> dcoef;
	// basis vectors come from the operator's store as they are first needed (p_i, q_i interleaved)
	double *const res = ops.work_vec(0), *const z = ops.work_vec(1);
	dcoef.alloc(2*(size_t)nrestart);
	std::vector<double*> p(nrestart, nullptr), q(nrestart, nullptr);
	auto need = [&](int i) { if(!p[i]) { p[i] = ops.work_vec(2 + 2*i); q[i] = ops.work_vec(3 + 2*i); } };
	need(0);
	std::vector<double> qq(nrestart, 0.0), beta(nrestart, 0.0);

	const double bnorm = std::sqrt(dot1(ops, b, b));
	double resnorm = 1.0;
	int step = 0;

	while(step < maxiter) {
		ops.gemv3(-1.0, x, 1.0, b, res);                   // r := b - A x
		ops.prec(res, p[0]);
		ops.spmv(p[0], q[0]);
		for(int k = 0; k < nrestart; k++) {
			double d3[2];
			{
				const double *aa[2] = {res, q[k]}, *bb[2] = {q[k], q[k]};
				ops.dots(2, aa, bb, d3);
			}
			qq[k] = d3[1];
			const double alpha = d3[0]/d3[1];
			launch_axpby(n, 1.0, x, alpha, p[k], st);
			launch_axpby(n, 1.0, res, -alpha, q[k], st);
			resnorm = std::sqrt(dot1(ops, res, res));
			step++;
			if(resnorm/bnorm < tol) break;
			if(k == nrestart-1) break;
			if(step >= maxiter) break;

			need(k+1);
			ops.prec(res, z);
			ops.spmv(z, q[k+1]);
			B200_CUDA(cudaMemcpyAsync(p[k+1], z, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
			// beta_i = -(q_{k+1}.q_i)/(q_i.q_i), i <= k: fused multi-dots (tests/solvers.cpp:319-322)
			for(int i0 = 0; i0 <= k; i0 += MAX_DOTS) {
				const int nd = std::min(MAX_DOTS, k + 1 - i0);
				const double *aa[MAX_DOTS], *bb[MAX_DOTS];
				double out[MAX_DOTS];
				for(int i = 0; i < nd; i++) { aa[i] = q[k+1]; bb[i] = q[i0+i]; }
				ops.dots(nd, aa, bb, out);
				for(int i = 0; i < nd; i++) beta[i0+i] = -out[i]/qq[i0+i];
			}
			B200_CUDA(cudaMemcpyAsync(dcoef, beta.data(), (k+1)*sizeof(double),
			                          cudaMemcpyHostToDevice, st));
			B200_CUDA(cudaStreamSynchronize(st));           // beta is reused on the host
			for(int l0 = 0; l0 <= k; l0 += 32) {
				const int nv = std::min(32, k + 1 - l0);
				launch_multi_axpy(n, nv, p.data() + l0, dcoef.p + l0, p[k+1], st);
				launch_multi_axpy(n, nv, q.data() + l0, dcoef.p + l0, q[k+1], st);
			}
		}
		if(resnorm/bnorm < tol) break;
	}
	info->converged = resnorm/bnorm <= tol;
	info->iters = step;
	info->resnorm = resnorm;
	info->bnorm = bnorm;
}

/// Flexible GMRES(m): right preconditioning with a preconditioner that may change between
/// applications (what an asynchronous sweep is), classical Gram-Schmidt with one fused multi-dot
/// and one fused multi-axpy per iteration (PETSc's -ksp_type fgmres default), Givens rotations on
/// the host.  In exact arithmetic its iterates coincide with GCR's (tests/solvers.hpp:108-110) at
/// about half the orthogonalisation traffic: only the Krylov basis V is orthogonalised, the
/// preconditioned vectors Z are kept for the solution update.
void fgmres(KrylovOps& ops, const double *b, double *x, double tol, int maxiter, int m,
            b200_solve_info *info)
{
	const long long n = ops.n;
	cudaStream_t st = ops.stream;
	if(m < 1) throw Error("FGMRES: restart length must be positive");
	if(m + 1 > MAX_KRYLOV_DOTS)
		throw Error("FGMRES: restart length above " + std::to_string(MAX_KRYLOV_DOTS - 1) + " not supported");
	// basis vectors come from the operator's store as they are first needed (w, v_0, z_0, v_1, ...)
	double *const w = ops.work_vec(0);
	std::vector<double*> V(m+1, nullptr), Z(m, nullptr);
	auto need = [&](int i) {
		if(!V[i]) V[i] = ops.work_vec(1 + 2*i);
		if(i < m && !Z[i]) Z[i] = ops.work_vec(2 + 2*i);
	};
	need(0);
	std::vector<double> H((size_t)(m+1)*m, 0.0), cs(m), sn(m), g(m+1), y(m);
	DevBuf<double> dy;
	dy.alloc(m);

	const double bnorm = std::sqrt(dot1(ops, b, b));
	double resnorm = bnorm;
	int step = 0;
	bool done = false;
	while(step < maxiter && !done) {
		ops.gemv3(-1.0, x, 1.0, b, w);                         // r = b - A x
		const double beta = std::sqrt(dot1(ops, w, w));
		resnorm = beta;
		if(beta/bnorm < tol || beta == 0.0) break;
		launch_vec_scal(n, 1.0/beta, w, V[0], st);
		std::fill(g.begin(), g.end(), 0.0);
		g[0] = beta;
		int j = 0;
		for(; j < m && step < maxiter; j++) {
			need(j+1);
			ops.prec(V[j], Z[j]);                              // z_j = M_j^-1 v_j
			ops.spmv(Z[j], w);                                 // w = A z_j
			// classical Gram-Schmidt with ONE reduction per iteration: h_i = v_i . w and w . w in the
			// same fused multi-dot (one all-reduce, one host synchronisation), w -= sum h_i v_i, and
			// |w_new|^2 = w.w - sum h_i^2 because V is orthonormal.  When that difference loses
			// more than four digits to cancellation the norm is computed explicitly instead.
			double hn;
			{
				const double *aa[MAX_KRYLOV_DOTS], *bb[MAX_KRYLOV_DOTS];
				double out[MAX_KRYLOV_DOTS];
				for(int i = 0; i <= j; i++) { aa[i] = V[i]; bb[i] = w; }
				aa[j+1] = w; bb[j+1] = w;
				const double *dh = ops.dots(j + 2, aa, bb, out);
				double sumsq = 0;
				for(int i = 0; i <= j; i++) { H[(size_t)i*m + j] = out[i]; sumsq += out[i]*out[i]; }
				for(int l0 = 0; l0 <= j; l0 += 32)
					launch_multi_axpy(n, std::min(32, j + 1 - l0), V.data() + l0, dh + l0, w, st, -1.0);
				const double ww = out[j+1], hn2 = ww - sumsq;
				hn = (hn2 > 1e-4*ww) ? std::sqrt(hn2) : std::sqrt(dot1(ops, w, w));
			}
			H[(size_t)(j+1)*m + j] = hn;
			if(hn > 0) launch_vec_scal(n, 1.0/hn, w, V[j+1], st);
			// Givens rotations on column j
			for(int i = 0; i < j; i++) {
				const double t = cs[i]*H[(size_t)i*m + j] + sn[i]*H[(size_t)(i+1)*m + j];
				H[(size_t)(i+1)*m + j] = -sn[i]*H[(size_t)i*m + j] + cs[i]*H[(size_t)(i+1)*m + j];
				H[(size_t)i*m + j] = t;
			}
			const double a = H[(size_t)j*m + j], c = H[(size_t)(j+1)*m + j];
			const double d = std::hypot(a, c);
			cs[j] = d > 0 ? a/d : 1.0;
			sn[j] = d > 0 ? c/d : 0.0;
			H[(size_t)j*m + j] = d;
			H[(size_t)(j+1)*m + j] = 0.0;
			g[j+1] = -sn[j]*g[j];
			g[j] = cs[j]*g[j];
			resnorm = std::fabs(g[j+1]);
			step++;
			if(resnorm/bnorm < tol || hn == 0.0) { j++; done = true; break; }
		}
		// y = R^-1 g, x += Z y
		for(int i = j-1; i >= 0; i--) {
			double t = g[i];
			for(int k = i+1; k < j; k++) t -= H[(size_t)i*m + k]*y[k];
			y[i] = t/H[(size_t)i*m + i];
		}
		if(j > 0) {
			B200_CUDA(cudaMemcpyAsync(dy, y.data(), j*sizeof(double), cudaMemcpyHostToDevice, st));
			for(int l0 = 0; l0 < j; l0 += 32)
				launch_multi_axpy(n, std::min(32, j - l0), Z.data() + l0, dy.p + l0, x, st);
			B200_CUDA(cudaStreamSynchronize(st));
		}
	}
	info->iters = step;
	info->resnorm = resnorm;
	info->bnorm = bnorm;
	info->converged = resnorm/bnorm < tol;
}

}  // namespace

void krylov_solve(const std::string& solver, KrylovOps& ops, const double *d_b, double *d_x,
                  double tol, int maxiter, int restart, b200_solve_info *info)
{
	b200_solve_info local;
	if(!info) info = &local;
	*info = b200_solve_info();
	Timer tm(ops.stream);
	tm.start();
	if(solver == "bicgstab") bicgstab(ops, d_b, d_x, tol, maxiter, info);
	else if(solver == "gcr") gcr(ops, d_b, d_x, tol, maxiter, restart, info);
	else if(solver == "fgmres") fgmres(ops, d_b, d_x, tol, maxiter, restart, info);
	else if(solver == "richardson") richardson(ops, d_b, d_x, tol, maxiter, info);
	else throw Error("unknown solver '" + solver + "'");
	info->device_ms = tm.stop();
}

}  // namespace b200
