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
	if(ops.prec_reads_output()) {
		// the reference's work vectors are value-initialised (tests/solvers.cpp:264-270); a
		// preconditioner that sweeps in place must not start from what an earlier solve left
		B200_CUDA(cudaMemsetAsync(z, 0, n*sizeof(double), st));
		B200_CUDA(cudaMemsetAsync(p[0], 0, n*sizeof(double), st));
	}
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

/// Pinned host buffer (the Hessenberg columns travel through it without a pageable staging copy)
struct PinnedBuf {
	double *p = nullptr;
	explicit PinnedBuf(size_t n) { B200_CUDA(cudaMallocHost((void**)&p, std::max<size_t>(n, 1)*sizeof(double))); }
	~PinnedBuf() { if(p) cudaFreeHost(p); }
	PinnedBuf(const PinnedBuf&) = delete;
	PinnedBuf& operator=(const PinnedBuf&) = delete;
};

/// Flexible GMRES(m): right preconditioning with a preconditioner that may change between
/// applications (what an asynchronous sweep is), classical Gram-Schmidt with one fused multi-dot
/// and one fused update per iteration (PETSc's -ksp_type fgmres default), Givens rotations on
/// the host.  In exact arithmetic its iterates coincide with GCR's (tests/solvers.hpp:108-110) at
/// about half the orthogonalisation traffic: only the Krylov basis V is orthogonalised, the
/// preconditioned vectors Z are kept for the solution update.
///
/// The host is NOT in the iteration: the Gram-Schmidt coefficients stay on the device (the update
/// reads them there), a one-thread kernel turns them into the Hessenberg column and the
/// reciprocal norm of the new basis vector, and iterations are enqueued back to back.  Columns are
/// handed to the host a few iterations at a time (one copy + one synchronisation per hand-over;
/// the interval follows the observed convergence rate), where the rotations and the convergence
/// test run; iterations enqueued past the converged one are simply not used.  With eight ranks
/// this removes the all-ranks-drain-their-queues point that every iteration used to end in.
void fgmres(KrylovOps& ops, const double *b, double *x, double tol, int maxiter, int m,
            b200_solve_info *info)
{
	const long long n = ops.n;
	cudaStream_t st = ops.stream;
	if(m < 1) throw Error("FGMRES: restart length must be positive");
	if(m + 1 > MAX_KRYLOV_DOTS)
		throw Error("FGMRES: restart length above " + std::to_string(MAX_KRYLOV_DOTS - 1) + " not supported");
	static const bool every_iteration = getenv("B200_FGMRES_SYNC") != nullptr;   // A/B switch (development)
	const bool zero_z = ops.prec_reads_output();
	// basis vectors come from the operator's store as they are first needed (w, v_0, z_0, v_1, ...)
	double *const w = ops.work_vec(0);
	std::vector<double*> V(m+1, nullptr), Z(m, nullptr);
	auto need = [&](int i) {
		if(!V[i]) V[i] = ops.work_vec(1 + 2*i);
		if(i < m && !Z[i]) {
			Z[i] = ops.work_vec(2 + 2*i);
			// a preconditioner that sweeps in place starts from zero, as on the reference's
			// value-initialised vectors (tests/solvers.cpp:264-270), not from a previous solve
			if(zero_z) B200_CUDA(cudaMemsetAsync(Z[i], 0, n*sizeof(double), st));
		}
	};
	need(0);
	const int HS = m + 2;                    // column stride: h_0..h_j, slot m: h_{j+1,j}, slot m+1: flag
	std::vector<double> H((size_t)(m+1)*m, 0.0), cs(m), sn(m), g(m+1), y(m);
	DevBuf<double> dy, dH, dinv;
	dy.alloc(m); dH.alloc((size_t)m*HS); dinv.alloc(m);
	PinnedBuf hH((size_t)m*HS);

	const double bnorm = std::sqrt(dot1(ops, b, b));
	double resnorm = bnorm;
	int step = 0;
	bool done = false;

	// iteration j, enqueued only
	auto enqueue = [&](int j) {
		need(j+1);
		ops.prec(V[j], Z[j]);                              // z_j = M_j^-1 v_j
		ops.spmv(Z[j], w);                                 // w = A z_j
		const double *aa[MAX_KRYLOV_DOTS], *bb[MAX_KRYLOV_DOTS];
		for(int i = 0; i <= j; i++) { aa[i] = V[i]; bb[i] = w; }
		aa[j+1] = w; bb[j+1] = w;
		const double *dh = ops.dots_device(j + 2, aa, bb);     // h_i = v_i.w, w.w: one reduction
		launch_fgmres_column(j, dh, dH.p + (size_t)j*HS, m, dinv.p + j, st);
		int l0 = 0;
		for(; j + 1 - l0 > 32; l0 += 32)
			launch_multi_axpy(n, 32, V.data() + l0, dh + l0, w, st, -1.0);
		// v_{j+1} = (w - sum h_i v_i) / |.|
		launch_multi_axpy_scaled(n, j + 1 - l0, V.data() + l0, dh + l0, w, V[j+1], dinv.p + j, st, -1.0);
	};
	// iteration j with the host in the loop and an explicit norm (after a flagged cancellation)
	auto careful = [&](int j, double *hcol) {
		need(j+1);
		ops.prec(V[j], Z[j]);
		ops.spmv(Z[j], w);
		const double *aa[MAX_KRYLOV_DOTS], *bb[MAX_KRYLOV_DOTS];
		double out[MAX_KRYLOV_DOTS];
		for(int i = 0; i <= j; i++) { aa[i] = V[i]; bb[i] = w; }
		const double *dh = ops.dots(j + 1, aa, bb, out);
		for(int l0 = 0; l0 <= j; l0 += 32)
			launch_multi_axpy(n, std::min(32, j + 1 - l0), V.data() + l0, dh + l0, w, st, -1.0);
		const double hn = std::sqrt(dot1(ops, w, w));
		for(int i = 0; i <= j; i++) hcol[i] = out[i];
		hcol[m] = hn; hcol[m+1] = 0.0;
		if(hn > 0) launch_vec_scal(n, 1.0/hn, w, V[j+1], st);
	};

	int interval = 1;                                      // iterations enqueued per hand-over
	while(step < maxiter && !done) {
		ops.gemv3(-1.0, x, 1.0, b, w);                         // r = b - A x
		const double beta = std::sqrt(dot1(ops, w, w));
		resnorm = beta;
		if(beta/bnorm < tol || beta == 0.0) break;
		launch_vec_scal(n, 1.0/beta, w, V[0], st);
		std::fill(g.begin(), g.end(), 0.0);
		g[0] = beta;
		int jq = 0;                                        // columns enqueued
		int j = 0;                                         // columns accepted (rotated) on the host
		bool cycle_careful = false;
		while(j < m && step < maxiter && !done) {
			const int room = std::min(m - j, maxiter - step);
			const int target = j + std::max(1, std::min(room, every_iteration ? 1 : interval));
			if(!cycle_careful) {
				for(; jq < target; jq++) enqueue(jq);
				B200_CUDA(cudaMemcpyAsync(hH.p + (size_t)j*HS, dH.p + (size_t)j*HS,
				                          (size_t)(jq - j)*HS*sizeof(double), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
			}
			const int jend = cycle_careful ? j + 1 : jq;
			double prev = resnorm;
			for(; j < jend && !done; j++) {
				double *hcol = hH.p + (size_t)j*HS;
				if(cycle_careful) careful(j, hcol);
				else if(hcol[m+1] != 0.0) {
					// cancellation in |w|^2 - sum h^2: what was enqueued from here on used a bad
					// norm; redo this iteration (and the rest of the cycle) with explicit norms
					cycle_careful = true;
					careful(j, hcol);
					jq = j + 1;
				}
				for(int i = 0; i <= j; i++) H[(size_t)i*m + j] = hcol[i];
				const double hn = hcol[m];
				H[(size_t)(j+1)*m + j] = hn;
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
				prev = resnorm;
				resnorm = std::fabs(g[j+1]);
				step++;
				if(resnorm/bnorm < tol || hn == 0.0 || !std::isfinite(resnorm)) done = true;
			}
			// next hand-over: about 0.7 of the iterations the current rate still needs, 1..10
			if(!done) {
				const double rate = (prev > 0) ? resnorm/prev : 1.0;
				double needit = 10.0;
				if(rate < 1.0 && rate > 0.0 && resnorm > 0)
					needit = std::log(tol*bnorm/resnorm)/std::log(rate);
				interval = (int)std::max(1.0, std::min(10.0, std::ceil(0.7*needit)));
			}
		}
		// y = R^-1 g, x += Z y   (columns enqueued past j are not used)
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
	ops.check_prec();          // a failed exact substitution inside the solve is an error, not a result
}

}  // namespace b200
