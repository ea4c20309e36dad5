/** \file api.cu
 * \brief The C ABI of libblasted_b200.so (include/blasted_b200.h): handles, copies at the
 * boundary, error translation.  No exception crosses this file's extern "C" functions.
 */
#include "common.cuh"
#include "blockops.cuh"
#include <cstring>
#include <cmath>


namespace b200 {

std::atomic<long long> g_launches{0};
Profiler g_prof;

cudaEvent_t Profiler::get()
{
	if(!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
	cudaEvent_t e;
	B200_CUDA(cudaEventCreate(&e));
	return e;
}
void Profiler::begin(int kc, cudaStream_t st)
{
	Rec r; r.a = get(); r.b = get(); r.kc = kc;
	B200_CUDA(cudaEventRecord(r.a, st));
	pending.push_back(r);
}
void Profiler::end(cudaStream_t st) { B200_CUDA(cudaEventRecord(pending.back().b, st)); }
void Profiler::collect()
{
	for(Rec& r : pending) {
		float t = 0;
		if(cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
			ms[r.kc] += t; count[r.kc]++;
		}
		pool.push_back(r.a); pool.push_back(r.b);
	}
	pending.clear();
}
void Profiler::reset()
{
	collect();
	for(int i = 0; i < KC_COUNT; i++) { ms[i] = 0; count[i] = 0; }
}
Profiler::~Profiler() { /* events are released with the context */ }
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

template <typename F>
static int guarded(F&& f)
{
	try { f(); return 0; }
	catch(const std::exception& e) { set_error(e.what()); return 1; }
	catch(...) { set_error("unknown error"); return 1; }
}

static void require_device()
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
		cudaGetLastError();
		throw Error("no CUDA device available: libblasted_b200 has no CPU fallback");
	}
}

void finish_matrix(Mat& A, const int *h_diagind, cudaStream_t st)
{
	// row statistics for kernel selection
	A.avg_row_len = A.nbrows ? (double)A.nnzb/A.nbrows : 0.0;
	A.max_row_len = max_row_length(A, st);
	A.browind.alloc(std::max<long long>(A.nnzb, 1));
	build_browind(A, st);
	A.diagind.alloc(std::max(A.nbrows, 1));
	if(h_diagind) {
		B200_CUDA(cudaMemcpyAsync(A.diagind, h_diagind, A.nbrows*sizeof(int), cudaMemcpyHostToDevice, st));
		bool ok = true;
		for(int i = 0; i < A.nbrows; i++) if(h_diagind[i] < 0) { ok = false; break; }
		A.has_diag = ok;
		B200_CUDA(cudaStreamSynchronize(st));
	} else
		A.has_diag = (find_diagonals(A, st) == 0);
}

static void upload_values(Mat& A, const double *vals, bool from_host, cudaStream_t st)
{
	const size_t n = (size_t)A.nnzb*A.bs*A.bs;
	if(n == 0) return;
	const cudaMemcpyKind kind = from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
	if(A.bs > 1 && (A.blockstorage == B200_ROWMAJOR) != device_rowmajor(A.bs)) {
		// the device layout is fixed per block size (blockops.cuh): transpose on the way in,
		// through a fixed staging buffer so that no matrix-sized temporary is needed
		if(!from_host) {
			transpose_blocks(A.bs, A.nnzb, vals, A.vals, st);
		} else {
			const long long bs2 = (long long)A.bs*A.bs;
			const long long chunk_blocks = std::max<long long>(1, (32LL << 20)/(bs2*8));   // 32 MiB
			A.stage.alloc((size_t)std::min<long long>(chunk_blocks, A.nnzb)*bs2*2);
			double *buf[2] = { A.stage.p, A.stage.p + (size_t)std::min<long long>(chunk_blocks, A.nnzb)*bs2 };
			int which = 0;
			for(long long b0 = 0; b0 < A.nnzb; b0 += chunk_blocks, which ^= 1) {
				const long long nb = std::min<long long>(chunk_blocks, A.nnzb - b0);
				B200_CUDA(cudaMemcpyAsync(buf[which], vals + b0*bs2, nb*bs2*sizeof(double), kind, st));
				transpose_blocks(A.bs, nb, buf[which], A.vals.p + b0*bs2, st);
			}
		}
		B200_CUDA(cudaStreamSynchronize(st));
	} else {
		B200_CUDA(cudaMemcpyAsync(A.vals, vals, n*sizeof(double), kind, st));
		if(from_host) B200_CUDA(cudaStreamSynchronize(st));
	}
}

/// Host values -> device, chunk by chunk on a copy stream of the matrix; after every chunk has
/// landed (and has been converted to the device block layout) `per_chunk(first block, blocks)` is
/// called to enqueue work on the chunk on the matrix's stream - so that the layout conversion and
/// whatever the caller enqueues run BEHIND the next chunk's copy instead of after the whole upload.
/// Returns when the host buffer has been consumed (the enqueued device work may still be running).
template <typename F>
static void upload_values_pipelined(Mat& A, const double *vals, F&& per_chunk)
{
	const long long bs2 = (long long)A.bs*A.bs;
	if(A.nnzb == 0) return;
	cudaStream_t st = A.stream;
	if(!A.copy_stream) {
		B200_CUDA(cudaStreamCreateWithFlags(&A.copy_stream, cudaStreamNonBlocking));
		for(int i = 0; i < 2; i++) {
			B200_CUDA(cudaEventCreateWithFlags(&A.ev_up[i], cudaEventDisableTiming));
			B200_CUDA(cudaEventCreateWithFlags(&A.ev_free[i], cudaEventDisableTiming));
		}
	}
	const bool convert = A.bs > 1 && (A.blockstorage == B200_ROWMAJOR) != device_rowmajor(A.bs);
	const long long chunk_blocks = std::max<long long>(1, (32LL << 20)/(bs2*8));          // 32 MiB
	const long long cb = std::min<long long>(chunk_blocks, A.nnzb);
	if(convert) A.stage.alloc((size_t)cb*bs2*2);
	// the copy stream starts after what is already enqueued on the matrix's stream (the previous
	// values may still be in use there)
	B200_CUDA(cudaEventRecord(A.ev_free[0], st));
	B200_CUDA(cudaStreamWaitEvent(A.copy_stream, A.ev_free[0], 0));
	int which = 0;
	long long k = 0;
	for(long long b0 = 0; b0 < A.nnzb; b0 += chunk_blocks, which ^= 1, k++) {
		const long long nb = std::min<long long>(chunk_blocks, A.nnzb - b0);
		double *dst = convert ? A.stage.p + (size_t)which*cb*bs2 : A.vals.p + b0*bs2;
		// a staging buffer is free again once the conversion of the chunk before last has run
		if(convert && k >= 2) B200_CUDA(cudaStreamWaitEvent(A.copy_stream, A.ev_free[which], 0));
		B200_CUDA(cudaMemcpyAsync(dst, vals + b0*bs2, nb*bs2*sizeof(double), cudaMemcpyHostToDevice, A.copy_stream));
		B200_CUDA(cudaEventRecord(A.ev_up[which], A.copy_stream));
		B200_CUDA(cudaStreamWaitEvent(st, A.ev_up[which], 0));
		if(convert) {
			transpose_blocks(A.bs, nb, dst, A.vals.p + b0*bs2, st);
			B200_CUDA(cudaEventRecord(A.ev_free[which], st));
		}
		per_chunk(b0, nb);
	}
	B200_CUDA(cudaStreamSynchronize(A.copy_stream));        // the host buffer has been read
}

/// Copies a device vector of block values out in the caller's block layout
static void download_blocks(const Mat& A, long long nblocks, const double *d_src, double *h_dst,
                            cudaStream_t st)
{
	const size_t n = (size_t)nblocks*A.bs*A.bs;
	if(n == 0) return;
	if(A.bs > 1 && (A.blockstorage == B200_ROWMAJOR) != device_rowmajor(A.bs)) {
		DevBuf<double> tmp;
		tmp.alloc(n);
		transpose_blocks(A.bs, nblocks, d_src, tmp, st);
		B200_CUDA(cudaMemcpyAsync(h_dst, tmp, n*sizeof(double), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
	} else {
		B200_CUDA(cudaMemcpyAsync(h_dst, d_src, n*sizeof(double), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
	}
}

/// Single-GPU Krylov operations
struct LocalOps : public KrylovOps {
	const Mat *A;
	Prec *M;
	DevBuf<double> partial, dout;
	double prec_ms = 0;
	LocalOps(const Mat *A_, Prec *M_) : A(A_), M(M_) {
		n = A->dim();
		stream = A->stream;
		if(M && M->stream != A->stream)
			throw Error("solve: the matrix and the preconditioner are on different streams "
			            "(set both with b200_mat_set_stream / b200_prec_set_stream)");
		ws = &A->krylov_ws;
		partial.alloc((size_t)MAX_DOTS*DOT_BLOCKS);
		dout.alloc(MAX_KRYLOV_DOTS);
	}
	void spmv(const double *x, double *y) override { launch_spmv(*A, x, y, stream); }
	void gemv3(double a, const double *x, double b, const double *y, double *z) override {
		launch_gemv3(*A, a, x, b, y, z, stream);
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
		return dout.p;
	}
	bool prec_reads_output() const override { return M && prec_sweeps_in_place(*M); }
	void check_prec() override { if(M) prec_check(*M); }
};

namespace {
/// device copy of a host array, or the caller's device pointer itself
template <typename T>
struct Staged {
	DevBuf<T> buf;
	const T *p = nullptr;
	Staged(const T *src, size_t n, bool on_device, cudaStream_t st) {
		if(!src) return;
		if(on_device) { p = src; return; }
		buf.alloc(std::max<size_t>(n, 1));
		if(n) B200_CUDA(cudaMemcpyAsync(buf.p, src, n*sizeof(T), cudaMemcpyHostToDevice, st));
		p = buf.p;
	}
};
}

}  // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return g_error.c_str(); }

int b200_device_count(void)
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int b200_set_device(int device) { return guarded([&] { B200_CUDA(cudaSetDevice(device)); }); }

long long b200_kernel_launches(void) { return g_launches.load(); }
void b200_reset_kernel_launches(void) { g_launches.store(0); }

void b200_profile_enable(int on) { g_prof.enabled = on != 0; }
void b200_profile_reset(void) { try { g_prof.reset(); } catch(...) {} }
int b200_profile_get(double ms[8], long long count[8])
{
	return guarded([&] {
		g_prof.collect();
		for(int i = 0; i < KC_BASIC; i++) { ms[i] = g_prof.ms[i]; count[i] = g_prof.count[i]; }
	});
}
int b200_profile_classes(void) { return KC_COUNT; }
int b200_profile_get_n(int n, double *ms, long long *count)
{
	return guarded([&] {
		if(n < 0 || n > KC_COUNT || !ms || !count) throw Error("profile_get_n: invalid arguments");
		g_prof.collect();
		for(int i = 0; i < n; i++) { ms[i] = g_prof.ms[i]; count[i] = g_prof.count[i]; }
	});
}

// ------------------------------------------------------------------ matrix

static int mat_create(int nbrows, int bs, int blockstorage, const int *browptr, const int *bcolind,
                      const double *vals, const int *diagind, bool from_host, b200_mat **out)
{
	return guarded([&] {
		require_device();
		if(!out) throw Error("null output handle");
		*out = nullptr;
		if(nbrows < 0 || !browptr) throw Error("invalid matrix arguments");
		if(!(bs == 1 || bs == 3 || bs == 4 || bs == 5 || bs == 7))
			throw Error("Block size " + std::to_string(bs) + " not supported");
		if(blockstorage != B200_COLMAJOR && blockstorage != B200_ROWMAJOR)
			throw Error("Block ordering must be either rowmajor or colmajor!");
		b200_mat *h = new b200_mat;
		try {
			Mat& A = h->m;
			A.nbrows = nbrows; A.bs = bs; A.blockstorage = blockstorage;
			cudaStream_t st = A.stream;
			const cudaMemcpyKind kind = from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
			A.browptr.alloc((size_t)nbrows + 1);
			B200_CUDA(cudaMemcpyAsync(A.browptr, browptr, ((size_t)nbrows+1)*sizeof(int), kind, st));
			int nnzb = 0;
			if(from_host) nnzb = browptr[nbrows];
			else {
				B200_CUDA(cudaMemcpyAsync(&nnzb, browptr + nbrows, sizeof(int), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
			}
			A.nnzb = nnzb;
			A.bcolind.alloc(std::max<long long>(A.nnzb, 1));
			A.vals.alloc(std::max<size_t>((size_t)A.nnzb*bs*bs, 1));
			if(A.nnzb) B200_CUDA(cudaMemcpyAsync(A.bcolind, bcolind, A.nnzb*sizeof(int), kind, st));
			upload_values(A, vals, from_host, st);
			finish_matrix(A, from_host ? diagind : nullptr, st);
			B200_CUDA(cudaStreamSynchronize(st));
		} catch(...) { delete h; throw; }
		*out = h;
	});
}

int b200_mat_create_host(int nbrows, int bs, int blockstorage, const int *browptr,
                         const int *bcolind, const double *vals, const int *diagind, b200_mat **out)
{
	return mat_create(nbrows, bs, blockstorage, browptr, bcolind, vals, diagind, true, out);
}

int b200_mat_create_device(int nbrows, int bs, int blockstorage, const int *d_browptr,
                           const int *d_bcolind, const double *d_vals, b200_mat **out)
{
	return mat_create(nbrows, bs, blockstorage, d_browptr, d_bcolind, d_vals, nullptr, false, out);
}

int b200_mat_update_values_host(b200_mat *m, const double *vals)
{
	return guarded([&] { upload_values(m->m, vals, true, m->m.stream); });
}

int b200_mat_update_values_device(b200_mat *m, const double *d_vals)
{
	return guarded([&] { upload_values(m->m, d_vals, false, m->m.stream); });
}

void b200_mat_destroy(b200_mat *m)
{
	if(!m) return;
	for(int i = 0; i < 2; i++) {
		if(m->m.ev_up[i]) cudaEventDestroy(m->m.ev_up[i]);
		if(m->m.ev_free[i]) cudaEventDestroy(m->m.ev_free[i]);
	}
	if(m->m.copy_stream) cudaStreamDestroy(m->m.copy_stream);
	delete m;
}
int b200_mat_dim(const b200_mat *m) { return m->m.dim(); }
int b200_mat_nbrows(const b200_mat *m) { return m->m.nbrows; }
long long b200_mat_nnzb(const b200_mat *m) { return m->m.nnzb; }
int b200_mat_set_stream(b200_mat *m, void *s) { m->m.stream = (cudaStream_t)s; return 0; }
int b200_mat_release_workspace(b200_mat *m)
{
	return guarded([&] {
		B200_CUDA(cudaStreamSynchronize(m->m.stream));
		m->m.krylov_ws.release();
	});
}

// ---- front end (frontend.cu): coordinate input, permutation, scaling


int b200_mat_create_coo(int nrows, long long nnz, const int *rowind, const int *colind,
                        const double *vals, int bs, int blockstorage, int on_device, b200_mat **out)
{
	return guarded([&] {
		require_device();
		if(!out) throw Error("null output handle");
		*out = nullptr;
		if(!(bs == 1 || bs == 3 || bs == 4 || bs == 5 || bs == 7))
			throw Error("Block size " + std::to_string(bs) + " not supported");
		if(blockstorage != B200_COLMAJOR && blockstorage != B200_ROWMAJOR)
			throw Error("Block ordering must be either rowmajor or colmajor!");
		if(nnz > 0 && (!rowind || !colind || !vals)) throw Error("null coordinate arrays");
		b200_mat *h = new b200_mat;
		try {
			Mat& A = h->m;
			A.bs = bs; A.blockstorage = blockstorage;
			cudaStream_t st = A.stream;
			Staged<int> r(rowind, nnz, on_device != 0, st), c(colind, nnz, on_device != 0, st);
			Staged<double> v(vals, nnz, on_device != 0, st);
			coo_to_mat(A, nrows, nnz, r.p, c.p, v.p, st);
		} catch(...) { delete h; throw; }
		*out = h;
	});
}

int b200_mat_get_host(const b200_mat *m, int *browptr, int *bcolind, int *diagind, double *vals)
{
	return guarded([&] {
		const Mat& A = m->m;
		cudaStream_t st = A.stream;
		if(browptr) B200_CUDA(cudaMemcpyAsync(browptr, A.browptr.p, ((size_t)A.nbrows+1)*sizeof(int), cudaMemcpyDeviceToHost, st));
		if(bcolind && A.nnzb) B200_CUDA(cudaMemcpyAsync(bcolind, A.bcolind.p, A.nnzb*sizeof(int), cudaMemcpyDeviceToHost, st));
		if(diagind && A.nbrows) B200_CUDA(cudaMemcpyAsync(diagind, A.diagind.p, A.nbrows*sizeof(int), cudaMemcpyDeviceToHost, st));
		if(vals && A.nnzb) {
			const size_t n = (size_t)A.nnzb*A.bs*A.bs;
			if(A.bs > 1 && (A.blockstorage == B200_ROWMAJOR) != device_rowmajor(A.bs)) {
				DevBuf<double> t;
				t.alloc(n);
				transpose_blocks(A.bs, A.nnzb, A.vals, t, st);
				B200_CUDA(cudaMemcpyAsync(vals, t.p, n*sizeof(double), cudaMemcpyDeviceToHost, st));
				B200_CUDA(cudaStreamSynchronize(st));
			} else
				B200_CUDA(cudaMemcpyAsync(vals, A.vals.p, n*sizeof(double), cudaMemcpyDeviceToHost, st));
		}
		B200_CUDA(cudaStreamSynchronize(st));
	});
}

int b200_mat_reorder(b200_mat *m, const int *rord, const int *cord, int inverse, int on_device)
{
	return guarded([&] {
		Mat& A = m->m;
		Staged<int> r(rord, A.nbrows, on_device != 0, A.stream), c(cord, A.nbrows, on_device != 0, A.stream);
		mat_reorder(A, r.p, c.p, inverse != 0, A.stream);
	});
}

int b200_mat_scale(b200_mat *m, const double *rowscale, const double *colscale, int inverse,
                   int on_device)
{
	return guarded([&] {
		Mat& A = m->m;
		Staged<double> r(rowscale, A.nbrows, on_device != 0, A.stream), c(colscale, A.nbrows, on_device != 0, A.stream);
		mat_scale(A, r.p, c.p, inverse != 0, A.stream);
		B200_CUDA(cudaStreamSynchronize(A.stream));
	});
}

int b200_vec_reorder(double *vec, long long n, int bs, const int *ord, int inverse, int on_device)
{
	return guarded([&] {
		require_device();
		if(n < 0 || bs < 1) throw Error("invalid vector size");
		if(!ord || n == 0) return;                  // no ordering set: nothing to do (:222-225)
		if(on_device) { vec_reorder(n, bs, ord, inverse != 0, vec, 0); return; }
		Staged<int> o(ord, n, false, 0);
		DevBuf<double> v;
		v.alloc((size_t)n*bs);
		B200_CUDA(cudaMemcpyAsync(v.p, vec, (size_t)n*bs*sizeof(double), cudaMemcpyHostToDevice, 0));
		vec_reorder(n, bs, o.p, inverse != 0, v, 0);
		B200_CUDA(cudaMemcpy(vec, v.p, (size_t)n*bs*sizeof(double), cudaMemcpyDeviceToHost));
	});
}

int b200_vec_scale(double *vec, long long n, int bs, const double *scale, int inverse, int on_device)
{
	return guarded([&] {
		require_device();
		if(n < 0 || bs < 1) throw Error("invalid vector size");
		if(!scale || n == 0) return;
		if(on_device) { vec_scale(n, bs, scale, inverse != 0, vec, 0); B200_CUDA(cudaStreamSynchronize(0)); return; }
		Staged<double> sc(scale, n, false, 0);
		DevBuf<double> v;
		v.alloc((size_t)n*bs);
		B200_CUDA(cudaMemcpyAsync(v.p, vec, (size_t)n*bs*sizeof(double), cudaMemcpyHostToDevice, 0));
		vec_scale(n, bs, sc.p, inverse != 0, v, 0);
		B200_CUDA(cudaMemcpy(vec, v.p, (size_t)n*bs*sizeof(double), cudaMemcpyDeviceToHost));
	});
}

int b200_mat_apply(const b200_mat *m, const double *d_x, double *d_y)
{
	return guarded([&] { launch_spmv(m->m, d_x, d_y, m->m.stream); });
}

int b200_mat_gemv3(const b200_mat *m, double a, const double *d_x, double b, const double *d_y,
                   double *d_z)
{
	return guarded([&] { launch_gemv3(m->m, a, d_x, b, d_y, d_z, m->m.stream); });
}

int b200_mat_apply_host(const b200_mat *m, const double *x, double *y)
{
	return guarded([&] {
		const Mat& A = m->m;
		const size_t n = A.dim();
		A.hx.alloc(n); A.hy.alloc(n);
		B200_CUDA(cudaMemcpyAsync(A.hx, x, n*sizeof(double), cudaMemcpyHostToDevice, A.stream));
		launch_spmv(A, A.hx, A.hy, A.stream);
		B200_CUDA(cudaMemcpyAsync(y, A.hy, n*sizeof(double), cudaMemcpyDeviceToHost, A.stream));
		B200_CUDA(cudaStreamSynchronize(A.stream));
	});
}

int b200_mat_gemv3_host(const b200_mat *m, double a, const double *x, double b, const double *y,
                        double *z)
{
	return guarded([&] {
		const Mat& A = m->m;
		const size_t n = A.dim();
		A.hx.alloc(n); A.hy.alloc(n); A.hz.alloc(n);
		B200_CUDA(cudaMemcpyAsync(A.hx, x, n*sizeof(double), cudaMemcpyHostToDevice, A.stream));
		B200_CUDA(cudaMemcpyAsync(A.hy, y, n*sizeof(double), cudaMemcpyHostToDevice, A.stream));
		launch_gemv3(A, a, A.hx, b, A.hy, A.hz, A.stream);
		B200_CUDA(cudaMemcpyAsync(z, A.hz, n*sizeof(double), cudaMemcpyDeviceToHost, A.stream));
		B200_CUDA(cudaStreamSynchronize(A.stream));
	});
}

// ------------------------------------------------------------------ preconditioner

int b200_prec_create(const b200_settings *s, b200_mat *m, b200_prec **out)
{
	return guarded([&] {
		require_device();
		if(!s || !m || !out) throw Error("null argument");
		*out = nullptr;
		const Mat& A = m->m;
		if(s->bs != A.bs) throw Error("settings block size differs from the matrix block size");
		// SRFactory::create_preconditioner, src/solverfactory.cpp:131-228
		if(s->bs != 1) {
			if(s->blockstorage == B200_ROWMAJOR) {
				if(s->bs != 4)
					throw Error("Block size " + std::to_string(s->bs) + " not supported for row major!");
			} else if(s->blockstorage == B200_COLMAJOR) {
				if(s->bs != 4 && s->bs != 5)
					throw Error("Block size " + std::to_string(s->bs) + " not supported for column major!");
			} else
				throw Error("Block ordering must be either rowmajor or colmajor!");
		}
		b200_prec *h = new b200_prec;
		Prec& P = h->p;
		P.s = *s;
		P.A = &m->m;
		P.stream = A.stream;
		switch(s->prectype) {
		case B200_JACOBI: case B200_GS: case B200_SGS:
			P.is_jacobi_family = true; break;
		case B200_LEVEL_SGS:
			P.is_jacobi_family = true; P.uses_levels = true; break;
		case B200_ILU0: P.is_ilu = true; break;
		case B200_SEQILU0: P.is_ilu = true; P.threadedfactor = false; P.threadedapply = false; break;
		case B200_SFILU0: P.is_ilu = true; P.threadedfactor = false; break;
		case B200_SAPILU0: P.is_ilu = true; P.threadedapply = false; break;
		case B200_ASYNC_LEVEL_ILU0: P.is_ilu = true; P.uses_levels = true; break;
		case B200_NO_PREC: break;
		default:
			delete h;
			throw Error("Invalid preconditioner!");       // solverfactory.cpp:122,190
		}
		if(P.s.level_mode != B200_LEVELS_DAG && P.s.level_mode != B200_LEVELS_CONTIGUOUS)
			P.s.level_mode = B200_LEVELS_DAG;
		*out = h;
	});
}

int b200_prec_compute(b200_prec *p, double precinfo[6])
{
	return guarded([&] { prec_compute(p->p, precinfo); });
}

int b200_prec_compute_host(b200_prec *p, const double *vals, double precinfo[6])
{
	return guarded([&] {
		Prec& P = p->p;
		Mat& A = *P.A;
		if(!vals) { prec_compute(P, precinfo); return; }
		if(P.stream != A.stream)
			throw Error("compute: the matrix and the preconditioner are on different streams");
		if(prec_init_is_chunkable(P)) {
			// values, layout conversion and the initial guess of the factor travel chunk by chunk:
			// only the last chunk's conversion + initialisation are not hidden behind a copy
			upload_values_pipelined(A, vals, [&](long long b0, long long nb) {
				launch_ilu0_init_range(A, P.pl, P.sf, b0, b0 + nb, A.stream);
			});
			prec_compute(P, precinfo, true);
		} else {
			upload_values_pipelined(A, vals, [](long long, long long) {});
			prec_compute(P, precinfo);
		}
	});
}

int b200_prec_apply(b200_prec *p, const double *d_r, double *d_z)
{
	return guarded([&] { prec_apply(p->p, d_r, d_z); });
}

int b200_prec_apply_host(b200_prec *p, const double *r, double *z)
{
	return guarded([&] {
		Prec& P = p->p;
		const size_t n = P.dim();
		P.hr.alloc(n); P.hz.alloc(n);
		// r travels on a copy stream of its own: an asynchronous compute() enqueued just before
		// (the usual PCSetUp -> PCApply order) keeps the SMs busy while the link carries r
		// (pinned r; a pageable one is staged by the driver and gains nothing)
		if(!P.copy_stream) {
			B200_CUDA(cudaStreamCreateWithFlags(&P.copy_stream, cudaStreamNonBlocking));
			B200_CUDA(cudaEventCreateWithFlags(&P.ev_copy, cudaEventDisableTiming));
		}
		B200_CUDA(cudaMemcpyAsync(P.hr, r, n*sizeof(double), cudaMemcpyHostToDevice, P.copy_stream));
		B200_CUDA(cudaEventRecord(P.ev_copy, P.copy_stream));
		B200_CUDA(cudaStreamWaitEvent(P.stream, P.ev_copy, 0));
		if(P.s.prectype == B200_GS || (P.s.prectype == B200_SGS && P.s.apply_inittype == B200_INIT_A_NONE))
			B200_CUDA(cudaMemcpyAsync(P.hz, z, n*sizeof(double), cudaMemcpyHostToDevice, P.stream));
		prec_apply(P, P.hr, P.hz);
		B200_CUDA(cudaMemcpyAsync(z, P.hz, n*sizeof(double), cudaMemcpyDeviceToHost, P.stream));
		B200_CUDA(cudaStreamSynchronize(P.stream));
		prec_check(P);
	});
}

int b200_prec_apply_relax(b200_prec *p, const double *d_b, double *d_x, int maxits)
{
	return guarded([&] { prec_apply_relax(p->p, d_b, d_x, maxits > 0 ? maxits : p->p.maxits); });
}

int b200_prec_apply_relax_host(b200_prec *p, const double *b, double *x, int maxits)
{
	return guarded([&] {
		Prec& P = p->p;
		const size_t n = P.dim();
		P.hr.alloc(n); P.hz.alloc(n);
		B200_CUDA(cudaMemcpyAsync(P.hr, b, n*sizeof(double), cudaMemcpyHostToDevice, P.stream));
		B200_CUDA(cudaMemcpyAsync(P.hz, x, n*sizeof(double), cudaMemcpyHostToDevice, P.stream));
		prec_apply_relax(P, P.hr, P.hz, maxits > 0 ? maxits : P.maxits);
		B200_CUDA(cudaMemcpyAsync(x, P.hz, n*sizeof(double), cudaMemcpyDeviceToHost, P.stream));
		B200_CUDA(cudaStreamSynchronize(P.stream));
	});
}

int b200_prec_set_apply_params(b200_prec *p, double rtol, double atol, double dtol, int ctol, int maxits)
{
	Prec& P = p->p;
	P.rtol = rtol; P.atol = atol; P.dtol = dtol; P.ctol = ctol != 0; P.maxits = maxits;
	return 0;
}

int b200_prec_check(b200_prec *p) { return guarded([&] { prec_check(p->p); }); }

int b200_prec_dim(const b200_prec *p) { return p->p.dim(); }

int b200_prec_relaxation_available(const b200_prec *p)
{
	// include/solverops_jacobi.hpp:28,67; solverops_sgs.hpp:44,91; solverops_levels_sgs.hpp:22,57;
	// false for ILU0 (solverops_ilu0.hpp:51,125) and NoPreconditioner (solverops_base.hpp:91)
	return p->p.is_jacobi_family ? 1 : 0;
}

void b200_prec_destroy(b200_prec *p)
{
	if(!p) return;
	if(p->p.ev0) cudaEventDestroy(p->p.ev0);
	if(p->p.ev1) cudaEventDestroy(p->p.ev1);
	if(p->p.evc0) cudaEventDestroy(p->p.evc0);
	if(p->p.evc1) cudaEventDestroy(p->p.evc1);
	if(p->p.ev_copy) cudaEventDestroy(p->p.ev_copy);
	if(p->p.copy_stream) cudaStreamDestroy(p->p.copy_stream);
	for(void *g : p->p.level_graph) if(g) cudaGraphExecDestroy((cudaGraphExec_t)g);
	if(p->p.cap_stream) cudaStreamDestroy(p->p.cap_stream);
	delete p;
}

int b200_prec_set_stream(b200_prec *p, void *s) { p->p.stream = (cudaStream_t)s; return 0; }

int b200_prec_set_sweeps(b200_prec *p, int nbuild, int napply)
{
	p->p.s.nbuildsweeps = nbuild;
	p->p.s.napplysweeps = napply;
	return 0;
}

int b200_prec_positions_size(b200_prec *p, long long *npos)
{
	return guarded([&] {
		if(!p->p.pl.built) throw Error("ILU positions not built (not an ILU0 type, or compute() not called)");
		*npos = p->p.pl.npos;
	});
}

int b200_prec_pattern_stats(b200_prec *p, long long stats[5])
{
	return guarded([&] {
		if(!p->p.pl.built) throw Error("ILU positions not built (not an ILU0 type, or compute() not called)");
		pattern_stats(p->p.pl, stats, p->p.stream);
	});
}

int b200_prec_get_positions(b200_prec *p, int *posptr, int *lowerp, int *upperp)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.pl.built) throw Error("ILU positions not built");
		B200_CUDA(cudaMemcpy(posptr, P.pl.posptr, (P.A->nnzb+1)*sizeof(int), cudaMemcpyDeviceToHost));
		if(P.pl.npos) {
			B200_CUDA(cudaMemcpy(lowerp, P.pl.lowerp, P.pl.npos*sizeof(int), cudaMemcpyDeviceToHost));
			B200_CUDA(cudaMemcpy(upperp, P.pl.upperp, P.pl.npos*sizeof(int), cudaMemcpyDeviceToHost));
		}
	});
}

int b200_prec_levels_size(b200_prec *p, int *nlevels)
{
	return guarded([&] {
		if(!p->p.levels.built) throw Error("levels not built");
		*nlevels = p->p.levels.nlevels;
	});
}

int b200_prec_get_levels(b200_prec *p, int *level_ptr, int *level_rows)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.levels.built) throw Error("levels not built");
		std::memcpy(level_ptr, P.levels.level_ptr.data(), (P.levels.nlevels+1)*sizeof(int));
		if(level_rows) {
			if(P.levels.mode == B200_LEVELS_DAG)
				B200_CUDA(cudaMemcpy(level_rows, P.levels.level_rows, P.A->nbrows*sizeof(int),
				                     cudaMemcpyDeviceToHost));
			else
				for(int i = 0; i < P.A->nbrows; i++) level_rows[i] = i;
		}
	});
}

int b200_prec_get_factor(b200_prec *p, double *iluvals)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.is_ilu || !P.computed) throw Error("no ILU factor available");
		B200_CUDA(cudaStreamSynchronize(P.stream));
		const Mat& A = *P.A;
		if(A.bs == 1) {
			DevBuf<double> tmp;
			tmp.alloc(std::max<long long>(A.nnzb, 1));
			scalar_ilu0_gather(A, P.pl, P.sf, tmp, P.stream);
			download_blocks(A, A.nnzb, tmp, iluvals, P.stream);
			return;
		}
		// reference layout: matrix order, diagonal blocks hold their inverses
		// (async_blockilu_factor.cpp:144-146)
		DevBuf<double> tmp;
		tmp.alloc(std::max<size_t>((size_t)A.nnzb*A.bs*A.bs, 1));
		block_factor_assemble(A, P.pl, P.sf, P.dinv, tmp, P.stream);
		download_blocks(A, A.nnzb, tmp, iluvals, P.stream);
	});
}

int b200_prec_get_dblocks(b200_prec *p, double *dblocks)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.is_jacobi_family || !P.computed) throw Error("no inverted diagonal available");
		download_blocks(*P.A, P.A->nbrows, P.dinv, dblocks, P.stream);
	});
}

int b200_prec_get_scale(b200_prec *p, double *scale)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.scale.p || !P.computed) throw Error("no scaling vector available");
		B200_CUDA(cudaMemcpyAsync(scale, P.scale, P.dim()*sizeof(double), cudaMemcpyDeviceToHost, P.stream));
		B200_CUDA(cudaStreamSynchronize(P.stream));
	});
}

int b200_prec_ilu_residual(b200_prec *p, double *res)
{
	return guarded([&] {
		Prec& P = p->p;
		if(!P.is_ilu || !P.computed) throw Error("no ILU factor available");
		const Mat& A = *P.A;
		const double *scale = P.s.scale ? P.scale.p : nullptr;
		// the residual is defined with UN-inverted diagonal blocks (async_blockilu_factor.cpp:257-297
		// runs before :144-146), which is how the device keeps the factor
		*res = (A.bs == 1) ? scalar_ilu0_residual(A, P.pl, scale, P.sf, P.scratch, P.stream)
		                   : ilu0_residual(A, P.pl, scale, P.sf, P.scratch, P.stream);
	});
}

int b200_prec_last_times(b200_prec *p, double *compute_ms, double *apply_ms)
{
	return guarded([&] {
		Prec& P = p->p;
		if(compute_ms) {
			if(P.compute_timed && P.evc1 && cudaEventSynchronize(P.evc1) == cudaSuccess) {
				float ms = 0;
				if(cudaEventElapsedTime(&ms, P.evc0, P.evc1) == cudaSuccess) P.compute_ms = ms;
			}
			cudaGetLastError();
			*compute_ms = P.compute_ms;
		}
		if(apply_ms) {
			*apply_ms = 0;
			if(P.ev1 && cudaEventSynchronize(P.ev1) == cudaSuccess) {
				float ms = 0;
				if(cudaEventElapsedTime(&ms, P.ev0, P.ev1) == cudaSuccess) *apply_ms = ms;
			}
			cudaGetLastError();
		}
	});
}

// ------------------------------------------------------------------ Krylov drivers

int b200_solve(const char *solver, const b200_mat *A, b200_prec *M, const double *d_b, double *d_x,
               double tol, int maxiter, int restart, b200_solve_info *info)
{
	return guarded([&] {
		LocalOps ops(&A->m, M ? &M->p : nullptr);
		krylov_solve(solver, ops, d_b, d_x, tol, maxiter, restart, info);
	});
}

int b200_solve_host(const char *solver, const b200_mat *A, b200_prec *M, const double *b, double *x,
                    double tol, int maxiter, int restart, b200_solve_info *info)
{
	return guarded([&] {
		const size_t n = A->m.dim();
		DevBuf<double> db, dx;
		db.alloc(n); dx.alloc(n);
		cudaStream_t st = A->m.stream;
		B200_CUDA(cudaMemcpyAsync(db, b, n*sizeof(double), cudaMemcpyHostToDevice, st));
		B200_CUDA(cudaMemcpyAsync(dx, x, n*sizeof(double), cudaMemcpyHostToDevice, st));
		LocalOps ops(&A->m, M ? &M->p : nullptr);
		krylov_solve(solver, ops, db, dx, tol, maxiter, restart, info);
		B200_CUDA(cudaMemcpyAsync(x, dx, n*sizeof(double), cudaMemcpyDeviceToHost, st));
		B200_CUDA(cudaStreamSynchronize(st));
	});
}

}  // extern "C"
