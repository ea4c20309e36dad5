/** \file precond.cu
 * \brief compute() / apply() / apply_relax() of the device preconditioner objects.
 *
 * Host-side sweep drivers (the L2 layer of SURVEY.md section 1), one per reference class:
 *   AsyncILU0_SRPreconditioner / AsyncBlockILU0_SRPreconditioner   src/solverops_ilu0.cpp:56-383
 *   scalar_ilu0_factorize / block_ilu0_factorize                   src/async_ilu_factor.cpp:37-97,
 *                                                                  src/async_blockilu_factor.cpp:47-149
 *   AsyncSGS_SRPreconditioner / AsyncBlockSGS_SRPreconditioner     src/solverops_sgs.cpp:32-203
 *   JacobiSRPreconditioner / BJacobiSRPreconditioner               src/solverops_jacobi.cpp:31-220
 *   ChaoticRelaxation / ChaoticBlockRelaxation                     src/relaxation_chaotic.cpp:22-123
 *   Level_SGS / Level_BSGS                                         src/solverops_levels_sgs.cpp:54-221
 *   Async_Level_ILU0 / Async_Level_BlockILU0                       src/solverops_levels_ilu0.cpp:58-192
 *   NoPreconditioner                                               src/solverops_base.cpp:27-43
 *
 * "Sequential" variants (BLASTED_SEQILU0 / SFILU0 / SAPILU0: threadedfactor/threadedapply false,
 * src/solverfactory.cpp:93-107,163-180) are the EXACT operations on the device: the factorisation
 * iterates asynchronous sweeps until no entry changes bitwise (the ILU(0) fixed point, which one
 * sequential pass of the same row kernel produces in the reference), and the triangular solves run
 * level-scheduled substitution.
 */
#include "common.cuh"
#include <cmath>

namespace b200 {

static void ensure_events(Prec& P)
{
	if(!P.ev0) {
		B200_CUDA(cudaEventCreate(&P.ev0));
		B200_CUDA(cudaEventCreate(&P.ev1));
		B200_CUDA(cudaEventCreate(&P.evc0));
		B200_CUDA(cudaEventCreate(&P.evc1));
	}
}

// ------------------------------------------------------------------ compute

static bool exact_in_one_launch(Prec& P);

bool prec_init_is_chunkable(const Prec& P)
{
	return P.computed && P.is_ilu && P.A->bs > 1 && P.pl.built && !P.s.scale &&
	       P.s.fact_inittype == B200_INIT_F_ORIGINAL;
}

void prec_compute(Prec& P, double precinfo[6], bool init_done)
{
	Mat& A = *P.A;
	cudaStream_t st = P.stream;
	const int type = P.s.prectype;
	ensure_events(P);
	if(precinfo) for(int i = 0; i < 6; i++) precinfo[i] = 0;     // PrecInfo() value-initialised

	if(type == B200_NO_PREC) { P.computed = true; return; }
	if(A.nbrows == 0) { P.computed = true; return; }          // empty subdomain: nothing to build
	if(!A.has_diag) throw Error("preconditioner needs a structurally non-zero diagonal");
	if(!P.scratch.p) P.scratch.alloc(8);

	const bool first = !P.computed;
	B200_CUDA(cudaEventRecord(P.evc0, st));

	if(P.is_jacobi_family) {
		// BJacobiSRPreconditioner::compute, solverops_jacobi.cpp:31-48 / scalar_jacobi_setup :141-147
		if(!P.dinv.p) P.dinv.alloc((size_t)A.nbrows*A.bs*A.bs);
		launch_invert_diag_blocks(A, A.vals, A.diagind, P.dinv, true, st);
		if(first) {
			if(A.bs > 1) {
				int ml[2];
				part_max_lengths(A, ml, st);
				P.a_max_lower = ml[0]; P.a_max_upper = ml[1];
			}
			// ytemp allocated and zeroed once: solverops_sgs.cpp:36-43, solverops_levels_sgs.cpp:37-41
			P.ytemp.alloc(A.dim());
			B200_CUDA(cudaMemsetAsync(P.ytemp, 0, A.dim()*sizeof(double), st));
			if(P.uses_levels) build_levels(A, P.levels, P.s.level_mode, st);
			// scalar async SGS: sweeps run over a split copy of A (L part, strict U part) so that
			// each sweep streams one contiguous array through the staged kernel (csrstream.cu)
			if(type == B200_SGS && A.bs == 1 && stream_supported(A.max_row_len)) {
				build_split_csr(A, P.pl, st);
				P.sf.lval.alloc(std::max<long long>(P.pl.nlower, 1));
				P.sf.uval.alloc(std::max<long long>(P.pl.nstrict, 1));
			}
			// block async SGS (bs 4, 5): the same - the forward sweep streams the L parts, the backward
			// sweep the strict U parts, each a contiguous array (C3: 0.83 -> ~1.0 of peak per sweep);
			// the copy is refreshed at every compute() (one pass over A)
			if(type == B200_SGS && (A.bs == 4 || A.bs == 5)) {
				build_split_csr(A, P.pl, st);
				const size_t b2 = (size_t)A.bs*A.bs;
				P.sf.lval.alloc(std::max<size_t>((size_t)P.pl.nlower*b2, 1));
				P.sf.uval.alloc(std::max<size_t>((size_t)P.pl.nstrict*b2, 1));
			}
		}
		if(P.pl.split_built) gather_split_values(P.pl, A.vals, P.sf.lval, P.sf.uval, st, A.bs);
	}
	else if(P.is_ilu) {
		const bool scalar = (A.bs == 1);       // scalar factors live in split form (P.sf), see scalar_ilu.cu
		if(first) {
			// setup_storage + compute_ILU_positions_CSR_CSR on first call: solverops_ilu0.cpp:190-196,358-363
			P.ytemp.alloc(A.dim());
			B200_CUDA(cudaMemsetAsync(P.ytemp, 0, A.dim()*sizeof(double), st));
			if(P.s.scale) P.scale.alloc(A.dim());
			P.flag.alloc(1);
			// levels first, as Async_Level_*::compute does (solverops_levels_ilu0.cpp:52-56,139-145)
			if(P.uses_levels || !P.threadedapply || !P.threadedfactor)
				build_levels(A, P.levels, P.uses_levels ? P.s.level_mode : B200_LEVELS_DAG, st);
			build_ilu_pattern(A, P.pl, st);
			// the reference copies A into iluvals at allocation (solverops_ilu0.cpp:160-164,333-337);
			// this is what INIT_F_NONE then starts from
			if(scalar) scalar_ilu0_init(A, P.pl, nullptr, B200_INIT_F_ORIGINAL, P.sf, st);
			else {
				block_factor_alloc(A, P.pl, P.sf);
				launch_ilu0_init(A, P.pl, nullptr, B200_INIT_F_ORIGINAL, P.sf, st);
			}
			B200_CUDA(cudaEventRecord(P.evc0, st));      // time the factorisation proper
		}
		const double *scale = nullptr;
		if(P.s.scale) {
			launch_scaling_vector(A, P.scale, st);
			scale = P.scale;
		}
		// compact inverses of the diagonal blocks: written by the initialisation, kept current by
		// the upper launches of the sweeps
		double *dinv = nullptr;
		if(A.bs > 1) {
			if(!P.dinv.p) P.dinv.alloc((size_t)A.nbrows*A.bs*A.bs);
			dinv = P.dinv;
		}
		if(scalar) scalar_ilu0_init(A, P.pl, scale, P.s.fact_inittype, P.sf, st);
		else {
			// (inverting the initial diagonal blocks inside the init launch was measured slower: every
			// warp then runs the elimination, 322 us against 189 + 74 us for the two passes on C2)
			if(!init_done) launch_ilu0_init(A, P.pl, scale, P.s.fact_inittype, P.sf, st);
			launch_invert_diag_blocks(A, P.sf.udiag, nullptr, dinv, true, st);
		}

		// Async_Level_ILU0 (scalar) passes `threadedfactor`=true into the compute_info slot
		// (solverops_levels_ilu0.cpp:129-130): it always gathers PrecInfo.  Replicated.
		const bool info = P.s.compute_precinfo ||
			(type == B200_ASYNC_LEVEL_ILU0 && A.bs == 1);
		if(info && precinfo)
			precinfo[1] = scalar ? scalar_ilu0_residual(A, P.pl, scale, P.sf, P.scratch, st)
			                     : ilu0_residual(A, P.pl, scale, P.sf, P.scratch, st);

		// INIT_F_ORIGINAL / INIT_F_SGS already leave U_ij = (scaled) A_ij in every upper entry; the
		// entries without products never change from that, so only the first sweep after another
		// kind of initial guess has to touch them
		const bool const_upper_set = (P.s.fact_inittype == B200_INIT_F_ORIGINAL ||
		                              P.s.fact_inittype == B200_INIT_F_SGS ||
		                              (A.bs == 1 && P.s.fact_inittype == B200_INIT_F_ZERO));
		auto sweep = [&](int sw, int *flag) {
			const bool all_upper = (sw == 0 && !const_upper_set);
			if(scalar) scalar_ilu0_sweep(A, P.pl, scale, P.sf, flag, all_upper, st);
			else launch_ilu0_sweep(A, P.pl, scale, P.sf, dinv, flag, all_upper, st);
		};
		bool wrote_all_upper = !const_upper_set;      // every upper entry was (re)written by the sweeps
		if(P.threadedfactor) {
			for(int sw = 0; sw < P.s.nbuildsweeps; sw++) sweep(sw, nullptr);
			P.factor_sweeps_done = P.s.nbuildsweeps;
		}
		else if(P.s.nbuildsweeps > 0 && scalar && exact_in_one_launch(P)) {
			// exact factorisation: one launch over the level-sorted rows (scalar_ilu.cu)
			if(!P.rowdone.p) P.rowdone.alloc(std::max(A.nbrows, 1));
			B200_CUDA(cudaMemsetAsync(P.sync_flags, 0, 2*sizeof(int), st));
			scalar_ilu0_exact(A, P.pl, P.levels.level_rows, scale, P.sf, P.rowdone, P.sync_flags, st);
			int failed = 0;      // the sweeps path synchronises too (every fourth sweep)
			B200_CUDA(cudaMemcpyAsync(&failed, P.sync_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
			B200_CUDA(cudaStreamSynchronize(st));
			if(failed) throw Error("exact factorisation did not complete: a dependency never arrived");
			P.factor_sweeps_done = 1;
		}
		else if(P.s.nbuildsweeps > 0 && !scalar && exact_in_one_launch(P)) {
			// exact block factorisation: one launch over level-sorted, warp-padded rows (factor.cu)
			if(!P.rowdone.p) P.rowdone.alloc(std::max(A.nbrows, 1));
			if(!P.exact_slots.p) P.n_exact_slots = block_exact_slots(A, P.levels, P.exact_slots, st);
			B200_CUDA(cudaMemsetAsync(P.sync_flags, 0, 2*sizeof(int), st));
			launch_ilu0_exact(A, P.pl, P.exact_slots, P.n_exact_slots, scale, P.sf, dinv, P.rowdone,
			                  P.sync_flags, st);
			int failed = 0;
			B200_CUDA(cudaMemcpyAsync(&failed, P.sync_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
			B200_CUDA(cudaStreamSynchronize(st));
			if(failed) throw Error("exact factorisation did not complete: a dependency never arrived");
			P.factor_sweeps_done = 1;
			wrote_all_upper = true;
		}
		else if(P.s.nbuildsweeps > 0) {
			// exact factorisation: iterate to the bitwise fixed point
			// Every sweep makes at least one more dependency level final (an entry is a deterministic
			// function of final inputs once its row's predecessors are final), so nlevels sweeps reach
			// the fixed point; entries are compared bitwise, so a NaN (singular pivot) cannot keep
			// the loop alive.  Not converging within that bound is an error, not a result.
			int changed = 1, sw = 0;
			const int maxsw = P.levels.nlevels + 8;
			while(changed && sw < maxsw) {
				B200_CUDA(cudaMemsetAsync(P.flag, 0, sizeof(int), st));
				for(int rep = 0; rep < 4; rep++, sw++) sweep(sw, P.flag);
				B200_CUDA(cudaMemcpyAsync(&changed, P.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
			}
			P.factor_sweeps_done = sw;
			if(changed)
				throw Error("exact ILU(0) factorisation did not reach its fixed point within " +
				            std::to_string(sw) + " sweeps (" + std::to_string(P.levels.nlevels) + " levels)");
		}

		// blocks: the sweeps maintain the column-order copy of the strict upper part; the row-order
		// copy the triangular solves read catches up here (no entries at all for star stencils)
		if(!scalar && P.factor_sweeps_done > 0 && P.s.nbuildsweeps > 0)
			launch_sync_upper(A, P.pl, P.sf, wrote_all_upper, st);

		if(info && precinfo) {
			precinfo[0] = scalar ? scalar_ilu0_residual(A, P.pl, scale, P.sf, P.scratch, st)
			                     : ilu0_residual(A, P.pl, scale, P.sf, P.scratch, st);
			double dd[4];
			if(scalar) {
				DevBuf<double> tmp;
				tmp.alloc(std::max<long long>(A.nnzb, 1));
				scalar_ilu0_gather(A, P.pl, P.sf, tmp, st);
				diag_dominance(A, tmp, dd, P.scratch, st);
			} else {
				DevBuf<double> tmp;
				tmp.alloc((size_t)A.nnzb*A.bs*A.bs);
				block_factor_assemble(A, P.pl, P.sf, P.sf.udiag, tmp, st);
				diag_dominance(A, tmp, dd, P.scratch, st);
			}
			// PrecInfo layout: [2] upper min, [3] upper avg, [4] lower min, [5] lower avg
			// (preconditioner_diagnostics.hpp:20-30; arr = {lavg, lmin, uavg, umin})
			precinfo[5] = dd[0]; precinfo[4] = dd[1]; precinfo[3] = dd[2]; precinfo[2] = dd[3];
		}

		// "invert diagonal blocks in place" (async_blockilu_factor.cpp:144-146): the inverses of the
		// final U_ii are already in the compact array `dinv`, which is what the device triangular
		// solves read; the factor keeps U_ii itself and b200_prec_get_factor() substitutes the
		// inverses when the reference-layout factor is asked for.
	}
	else throw Error("Invalid preconditioner!");

	// asynchronous: the elapsed time is read when somebody asks (b200_prec_last_times)
	B200_CUDA(cudaEventRecord(P.evc1, st));
	P.compute_timed = true;
	P.computed = true;
}

// ------------------------------------------------------------------ deferred errors

void prec_check(Prec& P)
{
	if(!P.sync_flags.p) return;
	int flag = 0;
	B200_CUDA(cudaMemcpyAsync(&flag, P.sync_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, P.stream));
	B200_CUDA(cudaStreamSynchronize(P.stream));
	if(flag) {
		B200_CUDA(cudaMemsetAsync(P.sync_flags.p + 1, 0, sizeof(int), P.stream));
		throw Error("exact substitution did not complete: a dependency never arrived");
	}
}

// ------------------------------------------------------------------ level-scheduled sweeps

/// Runs `body` (a sequence of per-level launches reading P.lev_r and writing P.lev_z) as a CUDA
/// graph: thousands of tiny dependent launches are the whole cost of a level-scheduled solve, and a
/// graph replays them without per-launch driver work.  Captured once per preconditioner and `slot`
/// on a private stream (the legacy default stream cannot be captured), replayed on the handle's
/// stream; input/output go through fixed internal buffers so the kernel arguments never change.
template <typename Body>
static void run_level_graph(Prec& P, int slot, const double *r, double *z, Body body)
{
	const long long n = P.A->dim();
	cudaStream_t st = P.stream;
	if(!P.lev_r.p) { P.lev_r.alloc(n); P.lev_z.alloc(n); }
	B200_CUDA(cudaMemcpyAsync(P.lev_r, r, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
	if(P.levels.nlevels > 16384) {
		// degenerate schedules (the reference's contiguous levels on a stencil matrix give ~N
		// levels): a graph of that size is pointless, launch directly
		body();
		B200_CUDA(cudaMemcpyAsync(z, P.lev_z, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
		return;
	}
	if(!P.level_graph[slot]) {
		if(!P.cap_stream) B200_CUDA(cudaStreamCreateWithFlags(&P.cap_stream, cudaStreamNonBlocking));
		B200_CUDA(cudaStreamSynchronize(st));
		const bool prof = g_prof.enabled;
		g_prof.enabled = false;                 // no timing events inside a capture
		cudaStream_t saved = P.stream;
		P.stream = P.cap_stream;
		cudaGraph_t graph = nullptr;
		try {
			B200_CUDA(cudaStreamBeginCapture(P.cap_stream, cudaStreamCaptureModeThreadLocal));
			body();
			B200_CUDA(cudaStreamEndCapture(P.cap_stream, &graph));
		} catch(...) {
			cudaStreamEndCapture(P.cap_stream, &graph);
			if(graph) cudaGraphDestroy(graph);
			P.stream = saved; g_prof.enabled = prof;
			throw;
		}
		P.stream = saved; g_prof.enabled = prof;
		cudaGraphExec_t exec = nullptr;
		const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
		cudaGraphDestroy(graph);
		if(e != cudaSuccess) throw Error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
		P.level_graph[slot] = exec;
	}
	B200_CUDA(cudaGraphLaunch((cudaGraphExec_t)P.level_graph[slot], st));
	g_launches.fetch_add(2*(long long)P.levels.nlevels, std::memory_order_relaxed);
	B200_CUDA(cudaMemcpyAsync(z, P.lev_z, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
}

/// Exact triangular solves over a DAG schedule run as ONE launch per triangle over the level-sorted
/// rows (apply.cu::tri_syncfree_kernel); the reference-identical contiguous schedule, whose levels
/// are row ranges, keeps per-level launches replayed as a CUDA graph.  Measured per apply (L + U),
/// graph -> one launch: 7-point 256^3 (766 levels) 6.3 -> 3.5 ms, C2 (2047 levels of 512 block
/// rows) 16.5 -> 5.3 ms, bs=5 96^3 (286 levels) 3.25 -> 1.1 ms, 27-point 160^3 7.7 -> 6.1 ms;
/// results are bit-identical.  B200_LEVEL_GRAPH=1 forces the graph form (development).
static bool one_launch_levels(Prec& P)
{
	static const bool force_graph = getenv("B200_LEVEL_GRAPH") != nullptr;
	if(force_graph || P.levels.mode != B200_LEVELS_DAG || !P.levels.level_rows.p) return false;
	if(!P.sync_flags.p) {
		P.sync_flags.alloc(2);
		B200_CUDA(cudaMemsetAsync(P.sync_flags, 0, 2*sizeof(int), P.stream));
	}
	return true;
}

/// The exact scalar factorisation has a one-launch form too (scalar_ilu.cu::scalar_exact_kernel).
/// Its cost is levels x (one row's dependent load chain, ~20 us on a 7-point row), the sweeps' cost
/// is (sweeps to the bitwise fixed point) x (sweep time): measured 7-point 256^3 34.7 -> 15.4 ms,
/// 27-point 160^3 865 -> 173 ms, but 7-point 128^3 4.8 -> 6.7 ms - so it is used from a million
/// rows up, where the number of sweeps is what hurts.  B200_EXACT_SWEEPS=1 / B200_EXACT_ONE_LAUNCH=1
/// force either form (development, tests).
static bool exact_in_one_launch(Prec& P)
{
	static const bool force_sweeps = getenv("B200_EXACT_SWEEPS") != nullptr;
	static const bool force_one = getenv("B200_EXACT_ONE_LAUNCH") != nullptr;
	if(force_sweeps) return false;
	// blocks: a sweep costs the same whatever the size and the fixed point needs one per level, so
	// the one-launch form wins as soon as there are more than a few levels
	if(!force_one && P.A->bs == 1 && P.A->nbrows < (1 << 20)) return false;
	if(!force_one && P.A->bs > 1 && P.levels.nlevels < 8) return false;
	return one_launch_levels(P);
}

static void exact_pair(Prec& P, TriKind lower, TriKind upper, TriArgs aL, TriArgs aU,
                       const double *r, double *z)
{
	// lower: rhs r -> ytemp ; upper: rhs ytemp -> z, both over all rows in level order
	const int n = P.A->nbrows;
	aL.rows = aU.rows = P.levels.level_rows.p;
	aL.row_begin = aU.row_begin = 0; aL.row_end = aU.row_end = n;
	aL.rhs = r; aL.x = P.ytemp; aL.descending = false;
	aU.rhs = P.ytemp; aU.x = z; aU.descending = true;
	// the error flag is sticky: it is cleared only by whoever reads and reports it (prec_check: the
	// *_host entry points, the Krylov drivers, b200_prec_check); the ticket is reset per launch
	double *zz = z;
	if(z == r) {                       // in-place call: the upper solve may not overwrite r early
		if(!P.lev_z.p) { P.lev_r.alloc(P.A->dim()); P.lev_z.alloc(P.A->dim()); }
		aU.x = zz = P.lev_z;
	}
	launch_tri_syncfree(*P.A, lower, aL, P.sync_flags.p, P.sync_flags.p + 1, P.stream);
	launch_tri_syncfree(*P.A, upper, aU, P.sync_flags.p, P.sync_flags.p + 1, P.stream);
	if(zz != z)
		B200_CUDA(cudaMemcpyAsync(z, zz, P.A->dim()*sizeof(double), cudaMemcpyDeviceToDevice, P.stream));
}

static void level_sweep(Prec& P, TriKind kind, TriArgs a, bool backward)
{
	const Mat& A = *P.A;
	const Levels& lv = P.levels;
	a.rows = (lv.mode == B200_LEVELS_DAG) ? lv.level_rows.p : nullptr;
	a.descending = backward;
	if(!backward)
		for(int l = 0; l < lv.nlevels; l++) {
			a.row_begin = lv.level_ptr[l]; a.row_end = lv.level_ptr[l+1];
			launch_tri_sweep(A, kind, a, P.stream);
		}
	else
		for(int l = lv.nlevels-1; l >= 0; l--) {
			a.row_begin = lv.level_ptr[l]; a.row_end = lv.level_ptr[l+1];
			launch_tri_sweep(A, kind, a, P.stream);
		}
}

// ------------------------------------------------------------------ apply

void prec_apply(Prec& P, const double *r, double *z)
{
	Mat& A = *P.A;
	cudaStream_t st = P.stream;
	const int type = P.s.prectype;
	const long long n = A.dim();
	ensure_events(P);
	if(type != B200_NO_PREC && !P.computed) throw Error("apply() called before compute()");
	if(n == 0) return;
	B200_CUDA(cudaEventRecord(P.ev0, st));

	if(type == B200_NO_PREC) {
		// NoPreconditioner::apply, solverops_base.cpp:33-38
		if(r != z) B200_CUDA(cudaMemcpyAsync(z, r, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
	}
	else if(type == B200_JACOBI) {
		launch_jacobi_apply(A, P.dinv, r, z, st);
	}
	else if(type == B200_GS) {
		// ChaoticRelaxation::apply, relaxation_chaotic.cpp:22-45,92-106: forward sweeps in place on
		// whatever z holds on entry
		TriArgs a; a.vals = A.vals; a.dinv = P.dinv; a.rhs = r; a.x = z;
		a.row_begin = 0; a.row_end = A.nbrows;
		for(int sw = 0; sw < P.s.napplysweeps; sw++) launch_tri_sweep(A, TRI_RELAX, a, st);
	}
	else if(type == B200_SGS) {
		// AsyncSGS_SRPreconditioner::apply, solverops_sgs.cpp:48-83,148-177
		const int ai = P.s.apply_inittype;
		if(ai == B200_INIT_A_JACOBI || ai == B200_INIT_A_ZERO)
			B200_CUDA(cudaMemsetAsync(P.ytemp, 0, n*sizeof(double), st));
		TriArgs a; a.vals = A.vals; a.dinv = P.dinv; a.row_begin = 0; a.row_end = A.nbrows;
		a.rhs = r; a.x = P.ytemp; a.descending = false;
		a.max_part_len = P.a_max_lower;
		const bool split = P.pl.split_built && A.bs == 1;
		if(P.pl.split_built && A.bs > 1) {        // block sweeps over the split copy of A
			a.vals = P.sf.lval; a.part_ptr = P.pl.lptr; a.part_col = P.pl.lcol;
		}
		StreamArgs sa;
		if(split) {
			sa.ptr = P.pl.lptr; sa.col = P.pl.lcol; sa.val = P.sf.lval; sa.x = P.ytemp; sa.out = P.ytemp;
			sa.rhs = r; sa.diag = P.dinv; sa.row_end = A.nbrows; sa.chain = P.chain_sweeps;
		}
		for(int sw = 0; sw < P.s.napplysweeps; sw++) {
			if(split) {
				ProfScope ps(KC_TRI_LOWER, st);
				launch_csr_stream(STREAM_SGS_FWD, sa, std::max(P.pl.max_lower_len, 1), st);
			} else launch_tri_sweep(A, TRI_SGS_FWD, a, st);
		}
		if(ai == B200_INIT_A_JACOBI)
			B200_CUDA(cudaMemcpyAsync(z, P.ytemp, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
		else if(ai == B200_INIT_A_ZERO)
			B200_CUDA(cudaMemsetAsync(z, 0, n*sizeof(double), st));
		a.rhs = P.ytemp; a.x = z; a.descending = true;
		a.max_part_len = P.a_max_upper;
		if(P.pl.split_built && A.bs > 1) {
			a.vals = P.sf.uval; a.part_ptr = P.pl.uptr; a.part_col = P.pl.ucol;
		}
		if(split) {
			sa = StreamArgs();
			sa.ptr = P.pl.uptr; sa.col = P.pl.ucol; sa.val = P.sf.uval; sa.x = z; sa.out = z;
			sa.rhs = P.ytemp; sa.diag = P.dinv; sa.row_end = A.nbrows; sa.descending = 1;
			sa.chain = P.chain_sweeps;
		}
		for(int sw = 0; sw < P.s.napplysweeps; sw++) {
			if(split) {
				ProfScope ps(KC_TRI_UPPER, st);
				launch_csr_stream(STREAM_SGS_BWD, sa, std::max(P.pl.max_upper_len, 1), st);
			} else launch_tri_sweep(A, TRI_SGS_BWD, a, st);
		}
	}
	else if(type == B200_LEVEL_SGS) {
		// Level_SGS::apply, solverops_levels_sgs.cpp:54-87,160-189
		if(one_launch_levels(P)) {
			TriArgs a; a.vals = A.vals; a.dinv = P.dinv;
			exact_pair(P, TRI_SGS_FWD, TRI_SGS_BWD, a, a, r, z);
		}
		else run_level_graph(P, 0, r, z, [&] {
			TriArgs a; a.vals = A.vals; a.dinv = P.dinv;
			a.rhs = P.lev_r; a.x = P.ytemp;
			level_sweep(P, TRI_SGS_FWD, a, false);
			a.rhs = P.ytemp; a.x = P.lev_z;
			level_sweep(P, TRI_SGS_BWD, a, true);
		});
	}
	else if(P.is_ilu) {
		const double *scale = P.s.scale ? P.scale.p : nullptr;
		const bool levelled = P.uses_levels || !P.threadedapply;
		const bool scalar = (A.bs == 1);
		TriArgs a; a.row_begin = 0; a.row_end = A.nbrows;
		a.dinv = (A.bs > 1) ? P.dinv.p : nullptr;       // compact U_ii^-1 (blocks)
		// factors are stored split: L part, strict U part (each its own CSR / block CSR), diagonal
		TriArgs aL = a, aU = a;
		aL.vals = P.sf.lval; aL.part_ptr = P.pl.lptr; aL.part_col = P.pl.lcol;
		aU.vals = P.sf.uval; aU.part_ptr = P.pl.uptr; aU.part_col = P.pl.ucol;
		aL.max_part_len = P.pl.max_lower_len; aU.max_part_len = P.pl.max_upper_len;
		if(scalar) aU.part_diag = P.sf.udiag;
		const bool stream = scalar && stream_supported(A.max_row_len);
		if(levelled) {
			// Async_Level_ILU0::apply (solverops_levels_ilu0.cpp:58-105,148-192), and the exact
			// triangular solves of the sequential variants
			if(!P.uses_levels && P.s.apply_inittype == B200_INIT_A_NONE)
				throw Error(" scalar_ilu0_apply: Invalid init type!");
			if(one_launch_levels(P)) {
				aL.rscale = scale; aU.rscale = nullptr;
				exact_pair(P, TRI_ILU_LOWER, TRI_ILU_UPPER, aL, aU, r, z);
			}
			else run_level_graph(P, 0, r, z, [&] {
				aL.rhs = P.lev_r; aL.rscale = scale; aL.x = P.ytemp;
				level_sweep(P, TRI_ILU_LOWER, aL, false);
				aU.rhs = P.ytemp; aU.rscale = nullptr; aU.x = P.lev_z;
				level_sweep(P, TRI_ILU_UPPER, aU, true);
			});
		}
		else {
			// scalar_ilu0_apply / block_ilu0_apply, solverops_ilu0.cpp:56-148,240-321.
			// z := S r is folded into the L sweep (rhs scaled on the fly).
			const int ai = P.s.apply_inittype;
			if(ai == B200_INIT_A_NONE) throw Error(" scalar_ilu0_apply: Invalid init type!");
			// The first two lower sweeps are ONE pass: from y = 0 the first gives y = S r, so the
			// second gathers S r itself (no memset, one pass over L less); likewise the first upper
			// sweep gathers its initial guess z = y from ytemp instead of from a copy.
			static const bool fuse_env = getenv("B200_NO_FUSE_FIRST") == nullptr;   // A/B switch
			const bool fuse = fuse_env && stream && P.s.napplysweeps >= 2;
			if(!fuse) B200_CUDA(cudaMemsetAsync(P.ytemp, 0, n*sizeof(double), st));
			aL.rhs = r; aL.rscale = scale; aL.x = P.ytemp; aL.descending = false;
			StreamArgs sa;
			if(stream) {
				sa.ptr = P.pl.lptr; sa.col = P.pl.lcol; sa.val = P.sf.lval; sa.x = P.ytemp;
				sa.out = P.ytemp; sa.rhs = r; sa.rscale = scale; sa.row_end = A.nbrows;
				sa.chain = P.chain_sweeps;
			}
			for(int sw = fuse ? 1 : 0; sw < P.s.napplysweeps; sw++) {
				if(stream) {
					ProfScope ps(KC_TRI_LOWER, st);
					sa.x = (fuse && sw == 1) ? r : P.ytemp;
					sa.xscale = (fuse && sw == 1) ? scale : nullptr;
					launch_csr_stream(STREAM_TRI_LOWER, sa, std::max(P.pl.max_lower_len, 1), st);
				} else launch_tri_sweep(A, TRI_ILU_LOWER, aL, st);
			}
			const bool fuse_u = fuse && ai == B200_INIT_A_JACOBI;
			if(fuse_u) {}
			else if(ai == B200_INIT_A_JACOBI)
				B200_CUDA(cudaMemcpyAsync(z, P.ytemp, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
			else
				B200_CUDA(cudaMemsetAsync(z, 0, n*sizeof(double), st));
			aU.rhs = P.ytemp; aU.rscale = nullptr; aU.x = z; aU.descending = true;
			if(stream) {
				sa = StreamArgs();
				sa.ptr = P.pl.uptr; sa.col = P.pl.ucol; sa.val = P.sf.uval; sa.x = z; sa.out = z;
				sa.rhs = P.ytemp; sa.diag = P.sf.udiag; sa.row_end = A.nbrows; sa.descending = 1;
				sa.chain = P.chain_sweeps;
			}
			for(int sw = 0; sw < P.s.napplysweeps; sw++) {
				if(stream) {
					ProfScope ps(KC_TRI_UPPER, st);
					sa.x = (fuse_u && sw == 0) ? P.ytemp.p : z;
					launch_csr_stream(STREAM_TRI_UPPER, sa, std::max(P.pl.max_upper_len, 1), st);
				} else launch_tri_sweep(A, TRI_ILU_UPPER, aU, st);
			}
		}
		if(scale) launch_vec_scale_copy(n, scale, z, z, st);      // z := S z
	}
	else throw Error("Invalid preconditioner!");

	B200_CUDA(cudaEventRecord(P.ev1, st));
}

// ------------------------------------------------------------------ relaxation

void prec_apply_relax(Prec& P, const double *b, double *x, int maxits)
{
	Mat& A = *P.A;
	cudaStream_t st = P.stream;
	const int type = P.s.prectype;
	const long long n = A.dim();
	if(type == B200_NO_PREC) return;                      // NoPreconditioner::apply_relax: nothing
	if(P.is_ilu) throw Error("ILU relaxation not implemented!");   // solverops_ilu0.cpp:215,382
	if(!P.computed) throw Error("apply_relax() called before compute()");
	if(n == 0) return;

	TriArgs a; a.vals = A.vals; a.dinv = P.dinv; a.rhs = b; a.x = x;
	a.row_begin = 0; a.row_end = A.nbrows;
	if(type == B200_JACOBI) {
		// solverops_jacobi.cpp:66-121,174-220 with ctol == false (what relax_local_blasted sets,
		// blasted_petsc.cpp:532): xtemp = relax(x); x = xtemp
		if(!P.xtemp.p) P.xtemp.alloc(n);
		double refdiffnorm = 1;
		for(int step = 0; step < maxits; step++) {
			a.x = P.xtemp; a.xsrc = x;
			launch_tri_sweep(A, TRI_RELAX, a, st);
			if(P.ctol) {
				// tolerance checks of solverops_jacobi.cpp:86-105, 193-212
				launch_update_diffnorm(n, P.xtemp, x, P.scratch, st);
				double d2 = 0;
				B200_CUDA(cudaMemcpyAsync(&d2, P.scratch.p, sizeof(double), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
				const double diffnorm = std::sqrt(d2);
				if(step == 0) refdiffnorm = diffnorm;
				if(diffnorm < P.atol || diffnorm/refdiffnorm < P.rtol || diffnorm/refdiffnorm > P.dtol)
					break;
			} else
				B200_CUDA(cudaMemcpyAsync(x, P.xtemp, n*sizeof(double), cudaMemcpyDeviceToDevice, st));
		}
	}
	else if(type == B200_GS) {
		for(int step = 0; step < maxits; step++) launch_tri_sweep(A, TRI_RELAX, a, st);
	}
	else if(type == B200_SGS) {
		// solverops_sgs.cpp:86-116,180-203
		for(int step = 0; step < maxits; step++) {
			a.descending = false; launch_tri_sweep(A, TRI_RELAX, a, st);
			a.descending = true;  launch_tri_sweep(A, TRI_RELAX, a, st);
		}
	}
	else if(type == B200_LEVEL_SGS) {
		// solverops_levels_sgs.cpp:90-126,192-221
		for(int step = 0; step < maxits; step++) {
			level_sweep(P, TRI_RELAX, a, false);
			level_sweep(P, TRI_RELAX, a, true);
		}
	}
	else throw Error("Invalid preconditioner!");
}

}  // namespace b200
