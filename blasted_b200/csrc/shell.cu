/** \file shell.cu
 * \brief PETSc-free core of the PCSHELL glue (include/blasted_b200_shell.h): options -> settings,
 * first-call construction, value refresh + compute, apply, relaxation, time accounting and the
 * context list.  Replaces the non-PETSc half of src/blasted_petsc.cpp of the reference.
 */
#include "common.cuh"
#include "../../include/blasted_b200_shell.h"
#include <cstring>
#include <ctime>
#include <chrono>
#include <array>

namespace {

thread_local std::string g_shell_error;

struct InfoList { std::vector<std::array<double,6>> v; };

struct Stopwatch {
	std::chrono::steady_clock::time_point w0;
	std::clock_t c0;
	Stopwatch() : w0(std::chrono::steady_clock::now()), c0(std::clock()) { }
	void add(double& wall, double& cpu) const {
		wall += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
		cpu += (double)(std::clock() - c0)/CLOCKS_PER_SEC;
	}
};

struct TypeName { const char *name; int type; };
// include/solverfactory.hpp:22-42
const TypeName kTypes[] = {
	{"none", B200_NO_PREC}, {"jacobi", B200_JACOBI}, {"gs", B200_GS}, {"sgs", B200_SGS},
	{"ilu0", B200_ILU0}, {"seqilu0", B200_SEQILU0}, {"sfilu0", B200_SFILU0},
	{"sapilu0", B200_SAPILU0}, {"cscbgs", B200_CSC_BGS}, {"level_sgs", B200_LEVEL_SGS},
	{"async_level_ilu0", B200_ASYNC_LEVEL_ILU0}};

int type_from_string(const char *s)
{
	for(const TypeName& t : kTypes)
		if(std::strcmp(s, t.name) == 0) return t.type;
	throw std::invalid_argument("BLASTed: Preconditioner type not available!");
}

int fact_init_from_string(const char *s)
{
	// include/async_initialization_decl.hpp:37-48
	if(!std::strcmp(s, "init_zero")) return B200_INIT_F_ZERO;
	if(!std::strcmp(s, "init_original")) return B200_INIT_F_ORIGINAL;
	if(!std::strcmp(s, "init_sgs")) return B200_INIT_F_SGS;
	if(!std::strcmp(s, "init_none")) return B200_INIT_F_NONE;
	throw std::invalid_argument("Factor initialization not recongnized!");
}

int apply_init_from_string(const char *s)
{
	// include/async_initialization_decl.hpp:51-60
	if(!std::strcmp(s, "init_zero")) return B200_INIT_A_ZERO;
	if(!std::strcmp(s, "init_jacobi")) return B200_INIT_A_JACOBI;
	if(!std::strcmp(s, "init_none")) return B200_INIT_A_NONE;
	throw std::invalid_argument("Apply initialization not recongnized!");
}

bool uses_sweeps(int t) { return t != B200_JACOBI && t != B200_LEVEL_SGS && t != B200_NO_PREC; }
bool uses_factor_init(int t) { return t == B200_ILU0 || t == B200_SAPILU0 || t == B200_ASYNC_LEVEL_ILU0; }

template <typename F>
int shell_guarded(F&& f)
{
	try { f(); return 0; }
	catch(std::exception& e) { b200::set_error(e.what()); return 1; }
	catch(...) { b200::set_error("unknown error"); return 1; }
}

void need(int rc) { if(rc) throw std::runtime_error(b200_last_error()); }

void copy_str(char (&dst)[B200_OPT_STRLEN], const char *src)
{
	std::snprintf(dst, B200_OPT_STRLEN, "%s", src);
}

/// setSweeps_checkSeq, src/blasted_petsc.cpp:94-134
void sweeps_and_sequential(const b200_shell_node *ctx, b200_settings& s)
{
	s.nbuildsweeps = ctx->nbuildsweeps;
	s.napplysweeps = ctx->napplysweeps;
	if(s.prectype == B200_SEQILU0) return;
	const bool seqb = ctx->nbuildsweeps == B200_SEQUENTIAL_SYMBOL;
	const bool seqa = ctx->napplysweeps == B200_SEQUENTIAL_SYMBOL;
	if((seqa && seqb) || (seqa && s.prectype == B200_SFILU0) || (seqb && s.prectype == B200_SAPILU0)) {
		s.prectype = B200_SEQILU0;
		s.nbuildsweeps = 1; s.napplysweeps = 1;
		return;
	}
	if(seqa) {
		if(s.prectype != B200_ILU0 && s.prectype != B200_SAPILU0)
			throw std::runtime_error(" Seq. appl. only supported with async ILU factorization!");
		s.napplysweeps = 1;
		s.prectype = B200_SAPILU0;
	}
	if(seqb) {
		if(s.prectype != B200_ILU0 && s.prectype != B200_SFILU0)
			throw std::runtime_error(" Seq. fact. only supported with async triangular application!");
		s.nbuildsweeps = 1;
		s.prectype = B200_SFILU0;
	}
}

void settings_from_node(const b200_shell_node *ctx, b200_settings& s)
{
	std::memset(&s, 0, sizeof(s));
	s.prectype = type_from_string(ctx->prectypestr);
	s.bs = ctx->bs;
	s.blockstorage = B200_COLMAJOR;                  // "required for PETSc", :256
	s.scale = ctx->scale;
	sweeps_and_sequential(ctx, s);
	s.thread_chunk_size = ctx->threadchunksize;
	s.compute_precinfo = ctx->compute_precinfo;
	s.fact_inittype = B200_INIT_F_NONE;
	s.apply_inittype = B200_INIT_A_NONE;
	if(uses_sweeps(s.prectype)) {
		if(uses_factor_init(s.prectype)) s.fact_inittype = fact_init_from_string(ctx->factinittype);
		s.apply_inittype = apply_init_from_string(ctx->applyinittype);
	}
	s.relax = 0;
	s.level_mode = B200_LEVELS_DAG;
}

void relax_common(b200_shell_node *node, double rtol, double abstol, double dtol, int it)
{
	if(!node || !node->bprec) throw std::runtime_error("shell: relax before setup");
	need(b200_prec_set_apply_params(node->bprec, rtol, abstol, dtol, 0, it));
}

}  // namespace

extern "C" {

b200_shell_list b200_shell_list_new(void)
{
	b200_shell_list l;
	l.ctxlist = nullptr; l.size = 0;
	l.factorcputime = l.factorwalltime = l.applycputime = l.applywalltime = 0.0;
	return l;
}

b200_shell_node b200_shell_node_new(void)
{
	b200_shell_node n;
	std::memset(&n, 0, sizeof(n));
	return n;
}

void b200_shell_list_append(b200_shell_list *list, b200_shell_node node)
{
	// the new node becomes the head of the list, :378-388
	b200_shell_node *n = new b200_shell_node(node);
	n->next = list->ctxlist;
	list->ctxlist = n;
	list->size++;
}

void b200_shell_total_times(b200_shell_list *l)
{
	l->factorcputime = l->factorwalltime = l->applycputime = l->applywalltime = 0.0;
	for(b200_shell_node *n = l->ctxlist; n; n = n->next) {
		l->factorwalltime += n->factorwalltime; l->applywalltime += n->applywalltime;
		l->factorcputime += n->factorcputime; l->applycputime += n->applycputime;
	}
}

int b200_shell_cleanup(b200_shell_node *node)
{
	if(!node) return 0;
	if(node->bprec) { b200_prec_destroy(node->bprec); node->bprec = nullptr; }
	if(node->bmat) { b200_mat_destroy(node->bmat); node->bmat = nullptr; }
	return 0;
}

int b200_shell_list_destroy(b200_shell_list *l)
{
	while(l->ctxlist) {
		b200_shell_node *n = l->ctxlist;
		l->ctxlist = n->next;
		b200_shell_cleanup(n);
		delete static_cast<InfoList*>(n->infolist);
		delete n;
		l->size--;
	}
	if(l->size != 0) { b200::set_error("Could not delete Blasted_data_list properly!"); return 1; }
	return 0;
}

int b200_shell_set_options(b200_shell_node *node, const b200_shell_options *o)
{
	return shell_guarded([&] {
		if(!node || !o) throw std::runtime_error("null argument");
		copy_str(node->prectypestr, o->pc_type);
		const int ptype = type_from_string(node->prectypestr);
		int sweeps[2] = {1, 1};
		if(uses_sweeps(ptype)) {
			sweeps[0] = o->async_sweeps[0]; sweeps[1] = o->async_sweeps[1];
			if(uses_factor_init(ptype)) {
				node->scale = o->use_symmetric_scaling ? 1 : 0;
				copy_str(node->factinittype, o->fact_init_type);
			} else {
				node->scale = 0;
				copy_str(node->factinittype, "NA");
			}
			copy_str(node->applyinittype, o->apply_init_type);
			node->threadchunksize = o->thread_chunk_size;
		}
		node->compute_precinfo = o->compute_preconditioner_info ? 1 : 0;
		node->prectype = ptype;
		node->nbuildsweeps = sweeps[0];
		node->napplysweeps = sweeps[1];
		node->first_setup_done = 1;
		node->cputime = node->walltime = node->factorcputime = node->factorwalltime =
			node->applycputime = node->applywalltime = 0;
	});
}

int b200_shell_settings(const b200_shell_node *node, b200_settings *out)
{
	return shell_guarded([&] {
		if(!node || !out) throw std::runtime_error("null argument");
		settings_from_node(node, *out);
	});
}

int b200_shell_setup(b200_shell_node *node, int bs, int nbrows, const int *ia, const int *ja,
                     const double *a, const int *diag)
{
	return shell_guarded([&] {
		if(!node) throw std::runtime_error("null argument");
		if(!node->first_setup_done)
			throw std::runtime_error("shell: b200_shell_set_options must come before the first setup");
		bool created = false;
		if(!node->bprec) {
			created = true;
			// createNewPreconditioner, :216-311
			if(bs <= 0 || bs > 5 || bs == 2)
				throw std::runtime_error("BLASTed: Block size " + std::to_string(bs) + " is not supported!");
			if(diag)
				for(int i = 0; i < nbrows; i++)
					if(diag[i] < ia[i] || diag[i] >= ia[i+1] || ja[diag[i]] != i)
						throw std::runtime_error("! Zero diagonal in (block-)row " + std::to_string(i) + "!");
			node->bs = bs;
			b200_settings s;
			settings_from_node(node, s);
			need(b200_mat_create_host(nbrows, bs, B200_COLMAJOR, ia, ja, a, diag, &node->bmat));
			if(b200_prec_create(&s, node->bmat, &node->bprec)) {
				const std::string msg = b200_last_error();
				b200_mat_destroy(node->bmat); node->bmat = nullptr;
				throw std::runtime_error(msg);
			}
			delete static_cast<InfoList*>(node->infolist);
			node->infolist = node->compute_precinfo ? new InfoList : nullptr;
			if(node->infolist) static_cast<InfoList*>(node->infolist)->v.reserve(250);
		}
		else if(bs != node->bs || nbrows != b200_mat_nbrows(node->bmat))
			throw std::runtime_error("shell: the local matrix changed size between setups");
		Stopwatch sw;
		// updatePreconditioner, :314-327: PETSc has rewritten `a` in place; same pattern
		// (creation uploaded the values; later set-ups upload and factorise in one pipelined call)
		double info[6] = {0, 0, 0, 0, 0, 0};
		need(b200_prec_compute_host(node->bprec, (!created && a) ? a : nullptr, info));
		// compute() only enqueues; the set-up time of the reference (blasted_petsc.cpp:321-326) is
		// that of the finished factorisation, so wait for it before the stopwatch is read
		{ double ms = 0; need(b200_prec_last_times(node->bprec, &ms, nullptr)); }
		if(node->infolist) {
			std::array<double,6> rec;
			for(int i = 0; i < 6; i++) rec[i] = info[i];
			static_cast<InfoList*>(node->infolist)->v.push_back(rec);
		}
		sw.add(node->factorwalltime, node->factorcputime);
	});
}

int b200_shell_apply(b200_shell_node *node, const double *r, double *z)
{
	return shell_guarded([&] {
		if(!node || !node->bprec) throw std::runtime_error("shell: apply before setup");
		Stopwatch sw;
		need(b200_prec_apply_host(node->bprec, r, z));
		sw.add(node->applywalltime, node->applycputime);
	});
}

int b200_shell_apply_device(b200_shell_node *node, const double *d_r, double *d_z)
{
	return shell_guarded([&] {
		if(!node || !node->bprec) throw std::runtime_error("shell: apply before setup");
		Stopwatch sw;
		need(b200_prec_apply(node->bprec, d_r, d_z));
		sw.add(node->applywalltime, node->applycputime);      // enqueue time only: the call is asynchronous
	});
}

int b200_shell_relax(b200_shell_node *node, const double *rhs, double *x, double rtol, double abstol,
                     double dtol, int it, int guesszero, int *outits, int *reason)
{
	return shell_guarded([&] {
		relax_common(node, rtol, abstol, dtol, it);
		const int n = b200_prec_dim(node->bprec);
		if(guesszero) std::memset(x, 0, (size_t)n*sizeof(double));
		Stopwatch sw;
		need(b200_prec_apply_relax_host(node->bprec, rhs, x, it));
		sw.add(node->applywalltime, node->applycputime);
		if(reason) *reason = 4;                          // PCRICHARDSON_CONVERGED_ITS
		if(outits) *outits = it;
	});
}

int b200_shell_relax_device(b200_shell_node *node, const double *d_rhs, double *d_x, double rtol,
                            double abstol, double dtol, int it, int guesszero, int *outits, int *reason)
{
	return shell_guarded([&] {
		relax_common(node, rtol, abstol, dtol, it);
		const int n = b200_prec_dim(node->bprec);
		if(guesszero) B200_CUDA(cudaMemset(d_x, 0, (size_t)n*sizeof(double)));
		Stopwatch sw;
		need(b200_prec_apply_relax(node->bprec, d_rhs, d_x, it));
		sw.add(node->applywalltime, node->applycputime);
		if(reason) *reason = 4;
		if(outits) *outits = it;
	});
}

int b200_shell_offers_relaxation(const b200_shell_node *node)
{
	if(!node) return 0;
	return node->prectype != B200_ILU0 && node->prectype != B200_CSC_BGS && node->prectype != B200_NO_PREC;
}

int b200_shell_type_offers_relaxation(const char *pc_type)
{
	if(!pc_type) return 1;
	int t;
	try { t = type_from_string(pc_type); } catch(...) { return 1; }       // unknown: fails at the first set-up
	return t != B200_ILU0 && t != B200_CSC_BGS && t != B200_NO_PREC;
}

int b200_shell_info_count(const b200_shell_node *node)
{
	if(!node || !node->infolist) return 0;
	return (int)static_cast<const InfoList*>(node->infolist)->v.size();
}

int b200_shell_info_get(const b200_shell_node *node, int i, double precinfo[6])
{
	if(i < 0 || i >= b200_shell_info_count(node)) { b200::set_error("info index out of range"); return 1; }
	const auto& rec = static_cast<const InfoList*>(node->infolist)->v[i];
	for(int k = 0; k < 6; k++) precinfo[k] = rec[k];
	return 0;
}

}  // extern "C"
