/** \file common.cuh
 * \brief Internal definitions shared by the CUDA translation units of libblasted_b200.so.
 *
 * Device storage (the re-design of include/device_container.hpp, srmatrixdefs.hpp of the reference):
 * CSR/BSR arrays resident in HBM, int32 indices, fp64 values; the block layout on the device is
 * fixed per block size (blockops.cuh: bs=4 row-major and 128-byte aligned, others column-major) and
 * the caller's layout is converted at the boundary.
 */
#ifndef B200_COMMON_CUH
#define B200_COMMON_CUH

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <stdexcept>
#include <vector>
#include <memory>
#include <atomic>

#include "../../include/blasted_b200.h"

namespace b200 {

// ---------------------------------------------------------------- errors

void set_error(const std::string& msg);

struct Error : public std::runtime_error {
	explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define B200_CUDA(call)                                                                      \
	do {                                                                                     \
		cudaError_t e__ = (call);                                                            \
		if(e__ != cudaSuccess)                                                               \
			throw b200::Error(std::string(#call) + " failed: " + cudaGetErrorString(e__) +   \
			                  " (" __FILE__ ":" + std::to_string(__LINE__) + ")");            \
	} while(0)

extern std::atomic<long long> g_launches;

/// Call after every kernel launch: counts it and surfaces launch-configuration errors
#define B200_LAUNCHED()                                                                      \
	do {                                                                                     \
		b200::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
		B200_CUDA(cudaGetLastError());                                                       \
	} while(0)

// ---------------------------------------------------------------- per-kernel-class timing
//
// Optional (off by default): CUDA events around the launches of each kernel class on the launching
// stream, accumulated on request.  This is what bench.py uses for the live roofline figure.

enum KClass { KC_FACTOR_LOWER = 0, KC_FACTOR_UPPER, KC_FACTOR_INIT, KC_DIAG_INVERT, KC_TRI_LOWER,
              KC_TRI_UPPER, KC_SPMV, KC_OTHER,
              // classes of the Krylov drivers and of the partitioned layer (b200_profile_get_n)
              KC_BLAS1, KC_HALO_PACK, KC_HALO_WAIT, KC_ALLREDUCE, KC_COUNT };
constexpr int KC_BASIC = 8;             ///< what b200_profile_get's fixed arrays hold

struct Profiler {
	bool enabled = false;
	struct Rec { cudaEvent_t a, b; int kc; };
	std::vector<Rec> pending;
	std::vector<cudaEvent_t> pool;
	double ms[KC_COUNT] = {0};
	long long count[KC_COUNT] = {0};
	cudaEvent_t get();
	void begin(int kc, cudaStream_t st);
	void end(cudaStream_t st);
	void collect();                       ///< synchronises the pending events and accumulates
	void reset();
	~Profiler();
};
extern Profiler g_prof;

struct ProfScope {
	cudaStream_t st; bool on;
	ProfScope(int kc, cudaStream_t s) : st(s), on(g_prof.enabled) { if(on) g_prof.begin(kc, st); }
	~ProfScope() { if(on) g_prof.end(st); }
};

// ---------------------------------------------------------------- device buffers

template <typename T>
struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	DevBuf() {}
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	~DevBuf() { release(); }
	void release() { if(p) cudaFree(p); p = nullptr; n = 0; }
	void alloc(size_t count) {
		if(count == n && p) return;
		release();
		if(count == 0) return;
		// 64 bytes of slack: bulk async copies round their spans up to 16 bytes
		B200_CUDA(cudaMalloc((void**)&p, count*sizeof(T) + 64));
		n = count;
	}
	operator T*() const { return p; }
};

inline int div_up(long long a, long long b) { return (int)((a + b - 1)/b); }

/// Vectors for the restarted Krylov drivers, allocated a few at a time as the basis grows (a
/// solve that converges in 4 iterations should not pay for 62 vectors) and kept between solves.
struct KrylovStore {
	static constexpr int PER_CHUNK = 8;
	long long stride = 0;                                   ///< doubles per vector (even: 16-byte aligned)
	std::vector<std::unique_ptr<DevBuf<double>>> chunks;
	void release() { chunks.clear(); stride = 0; }
	/// New chunks are zero-filled on `st` (the reference's work vectors are value-initialised
	/// std::vectors, tests/solvers.cpp:264-270, and some preconditioners sweep in place on them)
	double *vec(long long n, int i, cudaStream_t st) {
		const long long need = (n + 1) & ~1LL;
		if(need != stride) { release(); stride = need; }
		const size_t c = (size_t)i / PER_CHUNK;
		while(chunks.size() <= c) {
			chunks.emplace_back(new DevBuf<double>());
			const size_t cnt = (size_t)std::max<long long>(stride, 2)*PER_CHUNK;
			chunks.back()->alloc(cnt);
			B200_CUDA(cudaMemsetAsync(chunks.back()->p, 0, cnt*sizeof(double), st));
		}
		return chunks[c]->p + (size_t)(i % PER_CHUNK)*stride;
	}
};

// ---------------------------------------------------------------- the matrix

/// Device-resident sparse (block-)row matrix.  Blocks are column-major on the device.
struct Mat {
	int nbrows = 0;
	int bs = 1;
	int blockstorage = B200_COLMAJOR;     ///< layout of the CALLER's blocks (for copy-in/out)
	long long nnzb = 0;
	DevBuf<int> browptr, bcolind, diagind, browind;   ///< browind[jj] = block-row of entry jj
	DevBuf<double> vals;
	bool has_diag = false;                ///< every row stores its diagonal (block)
	int max_row_len = 0;
	double avg_row_len = 0;
	cudaStream_t stream = 0;
	mutable DevBuf<double> hx, hy, hz;    ///< staging for the *_host entry points
	DevBuf<double> stage;                 ///< fixed staging buffer for layout-converting uploads
	cudaStream_t copy_stream = nullptr;   ///< host uploads of the values run beside the conversion / initialisation
	cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
	mutable KrylovStore krylov_ws;        ///< Krylov basis storage, kept between solves (grow-only)

	int dim() const { return nbrows*bs; }
};

// ---------------------------------------------------------------- kernels' host launchers
// (defined in the .cu files; every launcher is asynchronous on `st`)

// spmv.cu
void launch_spmv(const Mat& A, const double *x, double *y, cudaStream_t st);
void launch_gemv3(const Mat& A, double a, const double *x, double b, const double *y, double *z,
                  cudaStream_t st);

/// z(rows) += a (A x)(rows) over a device list of block rows; nonempty_rows builds the list of the
/// rows of A that hold entries (returns its length)
void launch_gemv_add_rows(const Mat& A, int nlist, const int *d_rows, double a, const double *x,
                          double *z, cudaStream_t st);
int nonempty_rows(const Mat& A, DevBuf<int>& rows, cudaStream_t st);

// storage.cu
void transpose_blocks(int bs, long long nblocks, const double *in, double *out, cudaStream_t st);
void build_browind(const Mat& A, cudaStream_t st);
/// locate diagonals; returns number of rows without one
int find_diagonals(Mat& A, cudaStream_t st);
int max_row_length(const Mat& A, cudaStream_t st);
/// derived arrays of a matrix whose browptr/bcolind are set: browind, diagind, row statistics (api.cu)
void finish_matrix(Mat& A, const int *h_diagind, cudaStream_t st);

// frontend.cu: coordinate triplets -> CSR/BSR, permutation and scaling of a resident matrix
void coo_to_mat(Mat& A, int nrows, long long nnz, const int *d_row, const int *d_col,
                const double *d_val, cudaStream_t st);
void mat_reorder(Mat& A, const int *d_rord, const int *d_cord, bool inverse, cudaStream_t st);
void vec_reorder(long long n, int bs, const int *d_ord, bool inverse, double *d_vec, cudaStream_t st);
void mat_scale(Mat& A, const double *d_rowscale, const double *d_colscale, bool inverse, cudaStream_t st);
void vec_scale(long long n, int bs, const double *d_scale, bool inverse, double *d_vec, cudaStream_t st);

// pattern.cu
struct IluPattern {
	long long npos = 0;
	DevBuf<int> posptr, lowerp, upperp;  ///< the reference's ILUPositions arrays (bit-identical)
	// Split form, built from the above (device-side design): strict lower part L, strict upper part
	// U (CSR each) and the diagonal as separate streams, so that every sweep reads contiguous
	// arrays, with packed per-phase work lists indexed into them
	long long nlower = 0, nupper = 0;
	long long nuwork = 0;                ///< upper entries that change between sweeps (have products
	                                     ///< or are diagonal); the others are U_ij = A_ij
	long long nstrict = 0;               ///< strict upper entries (= nupper - nbrows)
	int max_lower_len = 0, max_upper_len = 0;   ///< longest strict lower / strict upper row part
	DevBuf<int> lptr, lcol, uptr, ucol;
	DevBuf<int4> slmeta;                 ///< per lower entry: {entry, column, pos begin, pos end}
	DevBuf<int4> suall, suwork;          ///< per upper entry: {entry, pos begin, pos end, dest};
	                                     ///< dest = index into the U values (blocks: into the
	                                     ///< column-ordered copy UT), or ~row for a diagonal
	DevBuf<int2> spairs;                 ///< products re-indexed into the split L / U (blocks: UT) arrays
	DevBuf<int> ut_order, utpos;         ///< blocks: UT[d] = U[ut_order[d]], utpos = inverse
	DevBuf<int4> dmeta, dmeta_u;         ///< bs = 5 staged upper launch, per row: {A entry of the diagonal,
	                                     ///< first L index, products, 0} and the U indices of the products
	bool diag_runs_ok = false;           ///< lists valid: diagonals only, <= 3 products, L partners a run
	bool diag_u_runs = false;            ///< ... and the U partners a run as well (column-ordered copy)
	DevBuf<int4> lrow_meta, lrow_cols;   ///< bs = 5 staged lower launch, per row: {first A entry, first L
	                                     ///< index, lower entries, 0} and the columns of the lower entries
	bool lower_rows_ok = false;          ///< lists valid: no lower entry has products, <= 3 per row
	DevBuf<int> lentry, uentry;          ///< split-only build (SGS): position in A of every L / U entry
	bool built = false, split_built = false;
};
void build_ilu_pattern(const Mat& A, IluPattern& pl, cudaStream_t st);
void pattern_stats(const IluPattern& pl, long long out[5], cudaStream_t st);
/// {longest strict lower row part, longest strict upper row part} of a matrix with diagonals
void part_max_lengths(const Mat& A, int out[2], cudaStream_t st);
/// Scalar matrices: only the split CSR structure (lptr/lcol/uptr/ucol, lentry/uentry, part lengths),
/// for sweeps over A itself (SGS); no ILU position lists.
void build_split_csr(const Mat& A, IluPattern& pl, cudaStream_t st);
/// lval[t] = vals[lentry[t]], uval[t] = vals[uentry[t]]  (scalars, or bs x bs blocks)
void gather_split_values(const IluPattern& pl, const double *vals, double *lval, double *uval,
                         cudaStream_t st, int bs = 1);

/// Values of the ILU(0) factor in split form (see IluPattern): scalars, or bs x bs blocks
struct ScalarFactor {
	DevBuf<double> lval, uval, udiag;   ///< strict lower part, strict upper part (row order), diagonal
	DevBuf<double> alow, aupw;   ///< scalar: (scaled) entries of A in lower-list / upper-work-list order
	DevBuf<double> ut;           ///< blocks: strict upper part in column order (factor.cu)
};

// scalar_ilu.cu
void scalar_ilu0_init(const Mat& A, const IluPattern& pl, const double *scale, int fact_init,
                      ScalarFactor& F, cudaStream_t st);
void scalar_ilu0_sweep(const Mat& A, const IluPattern& pl, const double *scale, ScalarFactor& F,
                       int *d_changed, bool all_upper, cudaStream_t st);
/// exact factorisation in one launch over the level-sorted rows; rowdone: nbrows ints of scratch,
/// flags: {ticket, error}
void scalar_ilu0_exact(const Mat& A, const IluPattern& pl, const int *level_rows, const double *scale,
                       ScalarFactor& F, int *rowdone, int *flags, cudaStream_t st);
double scalar_ilu0_residual(const Mat& A, const IluPattern& pl, const double *scale,
                            const ScalarFactor& F, double *d_scratch, cudaStream_t st);
/// out[entry] for every stored entry, in the reference's iluvals ordering
void scalar_ilu0_gather(const Mat& A, const IluPattern& pl, const ScalarFactor& F, double *out,
                        cudaStream_t st);

// csrstream.cu
enum StreamKind { STREAM_SPMV, STREAM_GEMV3, STREAM_TRI_LOWER, STREAM_TRI_UPPER,
                  STREAM_SGS_FWD, STREAM_SGS_BWD };
struct StreamArgs {
	const int *ptr = nullptr, *col = nullptr;     ///< CSR part (rows ptr[i]..ptr[i+1])
	const double *val = nullptr;
	const double *x = nullptr;                    ///< gathered vector
	double *out = nullptr;
	const double *rhs = nullptr, *rscale = nullptr, *diag = nullptr, *yin = nullptr;
	double alpha = 1, beta = 0;
	int row_begin = 0, row_end = 0;
	int descending = 0;
	const double *xscale = nullptr;   ///< sweeps: the gathered vector is x .* xscale
	int chain = 0;           ///< asynchronous sweeps: rows of a warp solved exactly along i -> i±1
};
bool stream_supported(int max_row_len);
void launch_csr_stream(StreamKind kind, const StreamArgs& a, int max_len, cudaStream_t st);

struct Levels {
	int mode = B200_LEVELS_DAG;
	int nlevels = 0;
	std::vector<int> level_ptr;          ///< host copy of the level boundaries (nlevels+1)
	DevBuf<int> d_level_ptr;
	DevBuf<int> level_rows;              ///< DAG mode: rows ordered by level; contiguous: identity (unused)
	bool built = false;
};
void build_levels(const Mat& A, Levels& lv, int mode, cudaStream_t st);

// factor.cu
void launch_scaling_vector(const Mat& A, double *scale, cudaStream_t st);
/// Block factors live in split form too (factor.cu, "storage").
void block_factor_alloc(const Mat& A, const IluPattern& pl, ScalarFactor& F);
/// Writes the initial guess (INIT_F_ORIGINAL / _SGS / _ZERO) into the split arrays.
void launch_ilu0_init(const Mat& A, const IluPattern& pl, const double *scale, int fact_init,
                      ScalarFactor& F, cudaStream_t st);
/// INIT_F_ORIGINAL (unscaled) on the block entries [e0, e1) only
void launch_ilu0_init_range(const Mat& A, const IluPattern& pl, ScalarFactor& F, long long e0,
                            long long e1, cudaStream_t st);
/// One asynchronous sweep (lower launch, then upper launch).  If d_changed is non-null it is set to
/// 1 when any entry's value changed bitwise (used to iterate to the exact fixed point).
/// `all_upper`: also recompute the upper entries without products (needed once when the initial
/// guess did not already set them to the scaled A, i.e. for INIT_F_ZERO / INIT_F_NONE).
void launch_ilu0_sweep(const Mat& A, const IluPattern& pl, const double *scale, ScalarFactor& F,
                       double *dinv, int *d_changed, bool all_upper, cudaStream_t st);
/// Exact block factorisation in one launch over level-sorted, warp-padded row slots (factor.cu);
/// rowdone: nbrows ints of scratch, flags: {ticket, error}
int block_exact_slots(const Mat& A, const Levels& lv, DevBuf<int>& slots, cudaStream_t st);
void launch_ilu0_exact(const Mat& A, const IluPattern& pl, const int *slots, int nslots,
                       const double *scale, ScalarFactor& F, double *dinv, int *rowdone, int *flags,
                       cudaStream_t st);
/// After the last sweep: the row-order copy of the strict upper part catches up with the
/// column-order one (entries of the upper work list, or all)
void launch_sync_upper(const Mat& A, const IluPattern& pl, ScalarFactor& F, bool all, cudaStream_t st);
/// Factor in the reference's matrix order; diagonal blocks taken from diag_src (U_ii or inverses)
void block_factor_assemble(const Mat& A, const IluPattern& pl, const ScalarFactor& F,
                           const double *diag_src, double *out, cudaStream_t st);
/// dst[positions[i]] <- src[i] for nbrows blocks (copies the compact inverted diagonal blocks into
/// the factor, the reference's in-place inversion, async_blockilu_factor.cpp:144-146)
void launch_scatter_blocks(const Mat& A, const double *src_compact, const int *positions, double *dst,
                           cudaStream_t st);
void launch_invert_diag_blocks(const Mat& A, const double *src_vals, const int *positions_or_null,
                               double *dst, bool dst_is_compact, cudaStream_t st);
double ilu0_residual(const Mat& A, const IluPattern& pl, const double *scale, const ScalarFactor& F,
                     double *d_scratch, cudaStream_t st);
void diag_dominance(const Mat& A, const double *vals, double out[4], double *d_scratch,
                    cudaStream_t st);

// apply.cu
enum TriKind { TRI_ILU_LOWER, TRI_ILU_UPPER, TRI_SGS_FWD, TRI_SGS_BWD, TRI_RELAX };
struct TriArgs {
	const double *vals = nullptr;    ///< factor (ILU) or matrix values (SGS/relax)
	const double *dinv = nullptr;    ///< compact inverted diagonal blocks (SGS/relax), or nullptr
	const double *rhs = nullptr;     ///< right-hand side vector of this sweep
	const double *rscale = nullptr;  ///< optional scaling applied to rhs on the fly (or nullptr)
	double *x = nullptr;             ///< vector updated by the sweep
	const double *xsrc = nullptr;    ///< gather source; nullptr = x (in-place, chaotic)
	const int *rows = nullptr;       ///< optional explicit row list (level scheduling)
	int row_begin = 0, row_end = 0;  ///< range of rows (or of positions in `rows`)
	bool descending = false;         ///< map CTAs to rows in descending order
	int max_part_len = 0;            ///< longest row part of this sweep (0 = unknown): selects the staged bs = 5 kernel
	// scalar split form (bs == 1 ILU): the part to sweep as its own CSR arrays; vals indexes it
	const int *part_ptr = nullptr, *part_col = nullptr;
	const double *part_diag = nullptr;
};
void launch_tri_sweep(const Mat& A, TriKind kind, const TriArgs& a, cudaStream_t st);
/// exact substitution over the level-sorted row list `a.rows` (all rows) in one launch; a.x must not
/// alias a.rhs; *err is raised if a dependency never arrives (invalid ordering)
void launch_tri_syncfree(const Mat& A, TriKind kind, const TriArgs& a, int *ticket, int *err,
                         cudaStream_t st);
void launch_jacobi_apply(const Mat& A, const double *dinv, const double *r, double *z,
                         cudaStream_t st);
void launch_vec_scale_copy(long long n, const double *scale, const double *in, double *out,
                           cudaStream_t st);   ///< out = scale ? scale*in : in
void launch_vec_fill(long long n, double v, double *out, cudaStream_t st);

// blas1.cu
void launch_axpby(long long n, double p, double *z, double q, const double *x, cudaStream_t st);
void launch_axpbypcz(long long n, double p, double *z, double q, const double *x, double r,
                     const double *y, cudaStream_t st);
/// out[0..nd) = dots of pairs (a[i], b[i]); result left on the device in d_out
constexpr int MAX_DOTS = 16;
constexpr int MAX_KRYLOV_DOTS = 64;  ///< products per KrylovOps::dots call (chunks of MAX_DOTS, one sync)
constexpr int DOT_BLOCKS = 592;      // 148 SMs x 4
/// d_out[i] = a[i] . b[i] for i < nd <= MAX_DOTS, deterministic two-stage reduction;
/// d_partial must hold MAX_DOTS*DOT_BLOCKS doubles.  Result stays on the device.
void launch_multi_dot(long long n, int nd, const double *const *a, const double *const *b,
                      double *d_partial, double *d_out, cudaStream_t st);
/// y += sum_l coef[l] * v[l]  for l < nv (coefficients read from device memory), nv <= 32 per call
void launch_multi_axpy(long long n, int nv, const double *const *v, const double *d_coef,
                       double *y, cudaStream_t st, double sign = 1.0);
/// out = (y + sign * sum_l coef[l] v[l]) * (*d_scale), nv <= 32; y is not modified
void launch_multi_axpy_scaled(long long n, int nv, const double *const *v, const double *d_coef,
                              const double *y, double *out, const double *d_scale, cudaStream_t st,
                              double sign = 1.0);
/// FGMRES iteration j, device side: Hessenberg column + reciprocal norm from the fused multi-dot
void launch_fgmres_column(int j, const double *d_dots, double *d_hcol, int hn_slot, double *d_inv_hn,
                          cudaStream_t st);
/// x = xtemp, returning sum (xtemp - x)^2 in d_out[0] (Jacobi relaxation with tolerance checks)
void launch_update_diffnorm(long long n, const double *xtemp, double *x, double *d_out, cudaStream_t st);
/// out = alpha * in
void launch_vec_scal(long long n, double alpha, const double *in, double *out, cudaStream_t st);

// ---------------------------------------------------------------- preconditioner object

/// Device counterpart of the SRPreconditioner subclasses (include/solverops_*.hpp of the reference)
struct Prec {
	b200_settings s;
	Mat *A = nullptr;
	cudaStream_t stream = 0;
	bool threadedfactor = true, threadedapply = true;   ///< false: exact ("sequential") variant
	bool is_ilu = false, is_jacobi_family = false, uses_levels = false;

	IluPattern pl;
	Levels levels;
	DevBuf<double> scale, ytemp, dinv, xtemp, scratch, dot_partial;
	ScalarFactor sf;                    ///< the ILU(0) factor, split form (scalars and blocks)
	DevBuf<int> flag;
	DevBuf<double> hr, hz;              ///< staging for *_host entry points
	bool computed = false;
	int factor_sweeps_done = 0;         ///< sweeps used by the last compute (exact variants iterate)
	// SolveParams of the reference (include/solverops_base.hpp:18-26), set by setApplyParams
	double rtol = 0, atol = 0, dtol = 1e30;
	bool ctol = false;
	int maxits = 1;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;      ///< around the last apply()
	cudaEvent_t evc0 = nullptr, evc1 = nullptr;    ///< around the last compute() (read lazily)
	bool compute_timed = false;
	cudaStream_t copy_stream = nullptr;            ///< host-pointer apply: r is uploaded beside a running compute()
	cudaEvent_t ev_copy = nullptr;
	double compute_ms = 0, apply_ms = 0;
	// level-scheduled applies replayed as CUDA graphs (precond.cu::run_level_graph)
	DevBuf<double> lev_r, lev_z;
	cudaStream_t cap_stream = nullptr;
	void *level_graph[2] = {nullptr, nullptr};        ///< cudaGraphExec_t
	DevBuf<int> sync_flags;                           ///< {ticket, error} of the one-launch exact solves
	DevBuf<int> rowdone;                              ///< per-row flags of the one-launch exact factorisation
	/// asynchronous scalar sweeps solve the rows of a warp exactly along i -> i±1 (csrstream.cu)
	bool chain_sweeps = getenv("B200_NO_CHAIN") == nullptr;          // A/B switch (development)
	int a_max_lower = 0, a_max_upper = 0;             ///< longest row parts of A (block SGS sweeps)
	DevBuf<int> exact_slots;                          ///< blocks: level-sorted rows, levels padded to whole warps
	int n_exact_slots = 0;

	int dim() const { return A->nbrows*A->bs; }
};

/// apply() sweeps in place on what z holds on entry (ChaoticRelaxation::apply; SGS with INIT_A_NONE)
inline bool prec_sweeps_in_place(const Prec& P)
{
	return P.s.prectype == B200_GS || (P.s.prectype == B200_SGS && P.s.apply_inittype == B200_INIT_A_NONE);
}

}  // namespace b200
// the opaque handles of the C ABI
struct b200_mat { b200::Mat m; };
struct b200_prec { b200::Prec p; };
namespace b200 {

/// init_done: the initial guess (INIT_F_ORIGINAL) has already been written for the current values
/// (chunk-wise, by b200_prec_compute_host)
void prec_compute(Prec& P, double precinfo[6], bool init_done = false);
/// whether prec_compute's initialisation can run chunk by chunk behind an upload of the values
bool prec_init_is_chunkable(const Prec& P);
void prec_apply(Prec& P, const double *d_r, double *d_z);
void prec_apply_relax(Prec& P, const double *d_b, double *d_x, int maxits);
/// Synchronises the handle's stream and raises (once) if a one-launch exact substitution recorded a
/// dependency that never arrived since the last check
void prec_check(Prec& P);

// ---------------------------------------------------------------- Krylov drivers (krylov.cu)

/// Operations a Krylov driver needs; the single-GPU implementation forwards to Mat/Prec, the
/// multi-GPU one (dist.cu) adds the halo exchange and the all-reduce.
struct KrylovOps {
	long long n = 0;                                         ///< local vector length
	cudaStream_t stream = 0;
	virtual ~KrylovOps() {}
	virtual void spmv(const double *x, double *y) = 0;
	virtual void gemv3(double a, const double *x, double b, const double *y, double *z) = 0;
	virtual void prec(const double *r, double *z) = 0;
	/// out[i] = a[i].b[i], i < nd, summed over all ranks, returned on the host; the return value
	/// points to the same results in device memory (valid until the next call)
	virtual const double *dots(int nd, const double *const *a, const double *const *b, double *out) = 0;
	/// the same without the hand-over to the host: launches (and all-reduces) only, result on the
	/// device, valid until the next dots call; no host synchronisation
	virtual const double *dots_device(int nd, const double *const *a, const double *const *b) = 0;
	/// Basis storage for the restarted solvers.  Taken from a buffer that outlives the solve (the
	/// operator's) when there is one: allocating and freeing gigabytes inside every solve costs
	/// tens of milliseconds of idle GPU.
	KrylovStore *ws = nullptr;
	KrylovStore own_ws;
	double *work_vec(int i) { return (ws ? *ws : own_ws).vec(n, i, stream); }
	/// true if prec() sweeps in place on what its output vector holds on entry (chaotic GS, SGS
	/// without an initial guess): the drivers then zero that vector before its first use in a solve
	virtual bool prec_reads_output() const { return false; }
	/// raises if the preconditioner recorded a failed exact substitution since the last check
	virtual void check_prec() {}
};

void krylov_solve(const std::string& solver, KrylovOps& ops, const double *d_b, double *d_x,
                  double tol, int maxiter, int restart, b200_solve_info *info);

}  // namespace b200
#endif
