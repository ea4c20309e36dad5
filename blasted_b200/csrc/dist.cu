/** \file dist.cu
 * \brief Row-partitioned operator and Krylov drivers across GPUs: one subdomain per GPU, local
 * block-Jacobi preconditioner, NCCL send/recv halo exchange for SpMV, NCCL all-reduce for dots.
 *
 * The reference has no distributed code of its own: it is the local sub-preconditioner under PETSc's
 * `-pc_type bjacobi` (doc/user-doc.md:36-40; one block per rank, src/blasted_petsc.cpp:604-606), the
 * outer MatMult / VecDot being PETSc's.  This file supplies that outer layer for the device-resident
 * drivers of krylov.cu:
 *   y = A x   : pack boundary entries -> ncclSend/ncclRecv (one group) -> diagonal-block SpMV
 *               (spmv.cu / csrstream.cu) -> off-diagonal-block SpMV accumulating into y;
 *   dots      : fused local multi-dot (blas1.cu) -> one ncclAllReduce(sum, fp64) -> host scalars;
 *   M^-1      : the local preconditioner of the diagonal block, no communication.
 * NCCL is loaded at run time (dlopen of the libnccl the process already uses, e.g. torch's), so
 * the single-GPU library carries no link-time dependency on it.
 */
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <memory>


namespace b200 {

// ------------------------------------------------------------------ NCCL, loaded lazily

struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
	                          cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static void nccl_load(const char *path)
{
	if(g_nccl.handle) return;
	const char *cands[] = { path, getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
	for(const char *c : cands) {
		if(!c || !*c) continue;
		g_nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
		if(g_nccl.handle) break;
	}
	if(!g_nccl.handle) throw Error(std::string("cannot load NCCL: ") + dlerror());
#define B200_SYM(field, name)                                                          \
	*(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                             \
	if(!g_nccl.field) throw Error(std::string("NCCL symbol missing: ") + name);
	B200_SYM(GetUniqueId, "ncclGetUniqueId")
	B200_SYM(CommInitRank, "ncclCommInitRank")
	B200_SYM(CommDestroy, "ncclCommDestroy")
	B200_SYM(Send, "ncclSend")
	B200_SYM(Recv, "ncclRecv")
	B200_SYM(AllReduce, "ncclAllReduce")
	B200_SYM(GroupStart, "ncclGroupStart")
	B200_SYM(GroupEnd, "ncclGroupEnd")
	B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
}

#define B200_NCCL(call)                                                                    \
	do {                                                                                   \
		ncclResult_t r__ = (call);                                                         \
		if(r__ != ncclSuccess)                                                             \
			throw b200::Error(std::string(#call) + " failed: " + g_nccl.GetErrorString(r__)); \
	} while(0)

struct Comm {
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
};

// ------------------------------------------------------------------ distributed matrix

__global__ void pack_kernel(const long long n, const int bs, const int *__restrict__ idx,
                            const double *__restrict__ x, double *__restrict__ buf)
{
	const long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(e >= n*bs) return;
	const long long i = e / bs;
	const int c = (int)(e - i*bs);
	buf[e] = x[(size_t)idx[i]*bs + c];
}

struct DistMat {
	Comm *comm = nullptr;
	Mat *diag = nullptr, *offd = nullptr;       ///< borrowed; offd columns index the halo buffer
	int bs = 1;
	int nhalo = 0;                              ///< halo (block) entries received
	std::vector<int> neigh, send_count, recv_count;
	long long nsend = 0;
	DevBuf<int> send_idx;                       ///< local (block) rows to send, grouped by neighbour
	DevBuf<double> send_buf, halo;
	DevBuf<int> offd_rows;                      ///< block rows that couple to another subdomain
	int n_offd_rows = 0;
	cudaStream_t stream = 0;
	// the exchange runs on its own stream so that it overlaps the diagonal-block product
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_packed = nullptr, ev_halo = nullptr;
	void ensure_comm_stream() {
		if(comm_stream) return;
		// highest priority: the exchange kernels are tiny and must get onto the SMs while the
		// diagonal-block product (posted first, thousands of CTAs) is still being dispatched
		int lo = 0, hi = 0;
		B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		B200_CUDA(cudaStreamCreateWithPriority(&comm_stream, cudaStreamNonBlocking, hi));
		B200_CUDA(cudaEventCreateWithFlags(&ev_packed, cudaEventDisableTiming));
		B200_CUDA(cudaEventCreateWithFlags(&ev_halo, cudaEventDisableTiming));
	}
	~DistMat() {
		if(ev_packed) cudaEventDestroy(ev_packed);
		if(ev_halo) cudaEventDestroy(ev_halo);
		if(comm_stream) cudaStreamDestroy(comm_stream);
	}
};

/// Packs the boundary rows on the compute stream and posts the sends/receives on the exchange
/// stream; returns true if an exchange is in flight (the caller waits on ev_halo before it reads
/// the halo).  Buffer reuse is ordered by the same two events: the next pack follows the wait on
/// ev_halo (sends done), the next receive follows ev_packed (previous halo reads done).  Only one
/// NCCL operation of the communicator is ever in flight (the all-reduces follow the product).
static bool halo_exchange(DistMat& D, const double *x)
{
	if(D.nsend > 0) {
		ProfScope ps(KC_HALO_PACK, D.stream);
		pack_kernel<<<div_up(D.nsend*D.bs, 256), 256, 0, D.stream>>>(D.nsend, D.bs, D.send_idx, x, D.send_buf);
		B200_LAUNCHED();
	}
	if(D.neigh.empty()) return false;
	D.ensure_comm_stream();
	cudaStream_t st = D.comm_stream;
	B200_CUDA(cudaEventRecord(D.ev_packed, D.stream));
	B200_CUDA(cudaStreamWaitEvent(st, D.ev_packed, 0));
	B200_NCCL(g_nccl.GroupStart());
	try {
		size_t so = 0, ro = 0;
		for(size_t k = 0; k < D.neigh.size(); k++) {
			const size_t ns = (size_t)D.send_count[k]*D.bs, nr = (size_t)D.recv_count[k]*D.bs;
			if(ns) B200_NCCL(g_nccl.Send(D.send_buf.p + so, ns, ncclFloat64, D.neigh[k], D.comm->comm, st));
			if(nr) B200_NCCL(g_nccl.Recv(D.halo.p + ro, nr, ncclFloat64, D.neigh[k], D.comm->comm, st));
			so += ns; ro += nr;
		}
	} catch(...) {
		g_nccl.GroupEnd();                 // never leave the group open behind an exception
		throw;
	}
	B200_NCCL(g_nccl.GroupEnd());
	B200_CUDA(cudaEventRecord(D.ev_halo, st));
	return true;
}

static void dist_spmv(DistMat& D, double a, const double *x, double b, const double *y, double *z,
                      bool plain)
{
	const bool inflight = halo_exchange(D, x);
	if(plain) launch_spmv(*D.diag, x, z, D.stream);      // overlaps the exchange
	else launch_gemv3(*D.diag, a, x, b, y, z, D.stream);
	if(inflight) {
		ProfScope ps(KC_HALO_WAIT, D.stream);
		B200_CUDA(cudaStreamWaitEvent(D.stream, D.ev_halo, 0));
	}
	if(D.n_offd_rows > 0)                                // z += a * A_offd * halo, boundary rows only
		launch_gemv_add_rows(*D.offd, D.n_offd_rows, D.offd_rows, plain ? 1.0 : a, D.halo, z, D.stream);
}

struct DistOps : public KrylovOps {
	DistMat *D;
	Prec *M;
	DevBuf<double> partial, dout;
	DistOps(DistMat *D_, Prec *M_) : D(D_), M(M_) {
		n = D->diag->dim();
		stream = D->stream;
		if(D->diag->stream != D->stream || (M && M->stream != D->stream))
			throw Error("dist solve: the matrix, the partitioned operator and the preconditioner "
			            "must share one stream");
		ws = &D->diag->krylov_ws;
		partial.alloc((size_t)MAX_DOTS*DOT_BLOCKS);
		dout.alloc(MAX_KRYLOV_DOTS);
	}
	void spmv(const double *x, double *y) override { dist_spmv(*D, 1.0, x, 0.0, nullptr, y, true); }
	void gemv3(double a, const double *x, double b, const double *y, double *z) override {
		dist_spmv(*D, a, x, b, y, z, false);
	}
	void prec(const double *r, double *z) override {
		if(M) prec_apply(*M, r, z);
		else B200_CUDA(cudaMemcpyAsync(z, r, n*sizeof(double), cudaMemcpyDeviceToDevice, stream));
	}
	const double *dots(int nd, const double *const *a, const double *const *b, double *out) override {
		const double *d = dots_device(nd, a, b);
		B200_CUDA(cudaMemcpyAsync(out, d, nd*sizeof(double), cudaMemcpyDeviceToHost, stream));
		B200_CUDA(cudaStreamSynchronize(stream));
		return d;
	}
	const double *dots_device(int nd, const double *const *a, const double *const *b) override {
		if(nd > MAX_KRYLOV_DOTS) throw Error("dots: too many products");
		for(int o = 0; o < nd; o += MAX_DOTS)
			launch_multi_dot(n, std::min(MAX_DOTS, nd - o), a + o, b + o, partial, dout.p + o, stream);
		if(D->comm->world > 1) {
			ProfScope ps(KC_ALLREDUCE, stream);
			B200_NCCL(g_nccl.AllReduce(dout.p, dout.p, nd, ncclFloat64, ncclSum, D->comm->comm, stream));
		}
		return dout.p;
	}
	bool prec_reads_output() const override { return M && prec_sweeps_in_place(*M); }
	void check_prec() override { if(M) prec_check(*M); }
};

static thread_local std::string g_derr;
template <typename F>
static int dguarded(F&& f)
{
	try { f(); return 0; }
	catch(const std::exception& e) { set_error(e.what()); return 1; }
	catch(...) { set_error("unknown error"); return 1; }
}

}  // namespace b200

using namespace b200;

struct b200_comm { b200::Comm c; };
struct b200_dist_mat { b200::DistMat d; };

extern "C" {

int b200_nccl_load(const char *path) { return dguarded([&] { nccl_load(path); }); }

int b200_comm_unique_id(char id[128])
{
	return dguarded([&] {
		nccl_load(nullptr);
		ncclUniqueId u;
		B200_NCCL(g_nccl.GetUniqueId(&u));
		std::memcpy(id, u.internal, 128);
	});
}

int b200_comm_create(const char id[128], int rank, int world, b200_comm **out)
{
	return dguarded([&] {
		b200_comm *h = new b200_comm;
		h->c.rank = rank; h->c.world = world;
		if(world == 1) { *out = h; return; }          // a single subdomain never communicates: no NCCL
		nccl_load(nullptr);
		ncclUniqueId u;
		std::memcpy(u.internal, id, 128);
		ncclResult_t r = g_nccl.CommInitRank(&h->c.comm, world, u, rank);
		if(r != ncclSuccess) { delete h; throw Error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
		*out = h;
	});
}

void b200_comm_destroy(b200_comm *c)
{
	if(!c) return;
	if(c->c.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->c.comm);
	delete c;
}

int b200_dist_mat_create(b200_comm *comm, b200_mat *diag, b200_mat *offd, int nhalo, int nneigh,
                         const int *neigh_ranks, const int *send_counts, const int *send_idx,
                         const int *recv_counts, b200_dist_mat **out)
{
	return dguarded([&] {
		if(!comm || !diag || !out) throw Error("null argument");
		std::unique_ptr<b200_dist_mat> h(new b200_dist_mat);      // released to the caller on success
		DistMat& D = h->d;
		D.comm = &comm->c; D.diag = &diag->m; D.offd = offd ? &offd->m : nullptr;
		D.bs = diag->m.bs; D.nhalo = nhalo; D.stream = diag->m.stream;
		long long ns = 0, nr = 0;
		for(int k = 0; k < nneigh; k++) {
			D.neigh.push_back(neigh_ranks[k]);
			D.send_count.push_back(send_counts[k]);
			D.recv_count.push_back(recv_counts[k]);
			ns += send_counts[k]; nr += recv_counts[k];
		}
		if(nr != nhalo) throw Error("halo plan: receive counts do not add up to nhalo");
		D.nsend = ns;
		D.send_idx.alloc(std::max<long long>(ns, 1));
		if(ns) B200_CUDA(cudaMemcpy(D.send_idx, send_idx, ns*sizeof(int), cudaMemcpyHostToDevice));
		D.send_buf.alloc(std::max<long long>(ns*D.bs, 1));
		D.halo.alloc(std::max<long long>((long long)nhalo*D.bs, 1));
		B200_CUDA(cudaMemset(D.halo, 0, std::max<long long>((long long)nhalo*D.bs, 1)*sizeof(double)));
		if(D.offd && D.offd->nnzb > 0) {
			if(D.offd->nbrows != D.diag->nbrows || D.offd->bs != D.bs)
				throw Error("halo plan: the coupling part must have the rows and block size of the diagonal part");
			D.n_offd_rows = nonempty_rows(*D.offd, D.offd_rows, D.stream);
		}
		*out = h.release();
	});
}

void b200_dist_mat_destroy(b200_dist_mat *d) { delete d; }

int b200_dist_mat_apply(b200_dist_mat *d, const double *d_x, double *d_y)
{
	return dguarded([&] { dist_spmv(d->d, 1.0, d_x, 0.0, nullptr, d_y, true); });
}

int b200_dist_mat_apply_with_halo(b200_dist_mat *d, const double *d_x, const double *d_halo, double *d_y)
{
	return dguarded([&] {
		DistMat& D = d->d;
		if(D.nhalo > 0 && !d_halo) throw Error("apply_with_halo: null halo");
		launch_spmv(*D.diag, d_x, d_y, D.stream);
		if(D.n_offd_rows > 0)
			launch_gemv_add_rows(*D.offd, D.n_offd_rows, D.offd_rows, 1.0, d_halo, d_y, D.stream);
	});
}

int b200_dist_solve(const char *solver, b200_dist_mat *A, b200_prec *M, const double *d_b,
                    double *d_x, double tol, int maxiter, int restart, b200_solve_info *info)
{
	return dguarded([&] {
		DistOps ops(&A->d, M ? &M->p : nullptr);
		krylov_solve(solver, ops, d_b, d_x, tol, maxiter, restart, info);
	});
}

/// Sum over all ranks of a host scalar array (small; for global norms in tests/bench)
int b200_comm_allreduce_sum(b200_comm *c, double *vals, int n)
{
	return dguarded([&] {
		DevBuf<double> d;
		d.alloc(n);
		B200_CUDA(cudaMemcpy(d, vals, n*sizeof(double), cudaMemcpyHostToDevice));
		if(c->c.world > 1) B200_NCCL(g_nccl.AllReduce(d.p, d.p, n, ncclFloat64, ncclSum, c->c.comm, 0));
		B200_CUDA(cudaMemcpy(vals, d, n*sizeof(double), cudaMemcpyDeviceToHost));
	});
}

}  // extern "C"
