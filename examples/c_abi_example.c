/* Plain-C use of libblasted_b200.so (include/blasted_b200.h + blasted_b200_shell.h): a 1-D Poisson
 * matrix with 4x4 diagonal blocks, asynchronous block-ILU(0) through the ABI and through the
 * PETSc-free PCSHELL core, checked against each other.
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Lblasted_b200 -lblasted_b200 \
 *       -Wl,-rpath,$PWD/blasted_b200 -lm -o c_abi_example && ./c_abi_example
 * Exit code 0 = results agree; 77 = no CUDA device (nothing to run). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "blasted_b200.h"
#include "blasted_b200_shell.h"

#define NB 64        /* block rows */
#define BS 4

static int fail(const char *what) { fprintf(stderr, "%s: %s\n", what, b200_last_error()); return 1; }

int main(void)
{
	if(b200_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 77; }

	/* block tridiagonal: diagonal blocks 4 I + coupling, off-diagonal blocks -I (column-major) */
	int *ia = malloc((NB + 1)*sizeof(int)), *ja = malloc(3*NB*sizeof(int)), *diag = malloc(NB*sizeof(int));
	double *a = calloc((size_t)3*NB*BS*BS, sizeof(double));
	int nnzb = 0;
	for(int i = 0; i < NB; i++) {
		ia[i] = nnzb;
		for(int j = i - 1; j <= i + 1; j++) {
			if(j < 0 || j >= NB) continue;
			ja[nnzb] = j;
			if(j == i) diag[i] = nnzb;
			for(int c = 0; c < BS; c++)
				for(int r = 0; r < BS; r++)
					a[(size_t)nnzb*BS*BS + c*BS + r] = (j == i) ? (r == c ? 4.0 : 0.1/(1 + r + c)) : (r == c ? -1.0 : 0.0);
			nnzb++;
		}
	}
	ia[NB] = nnzb;
	const int n = NB*BS;
	double *r = malloc(n*sizeof(double)), *z1 = malloc(n*sizeof(double)), *z2 = malloc(n*sizeof(double));
	for(int i = 0; i < n; i++) r[i] = sin(0.1*i);

	/* 1. the ABI itself: exact ILU(0) ("seqilu0") */
	b200_settings s;
	memset(&s, 0, sizeof(s));
	s.prectype = B200_SEQILU0; s.bs = BS; s.blockstorage = B200_COLMAJOR;
	s.nbuildsweeps = 1; s.napplysweeps = 1;
	s.fact_inittype = B200_INIT_F_ORIGINAL; s.apply_inittype = B200_INIT_A_ZERO;
	b200_mat *A = NULL; b200_prec *M = NULL; double info[6];
	if(b200_mat_create_host(NB, BS, B200_COLMAJOR, ia, ja, a, diag, &A)) return fail("mat_create");
	if(b200_prec_create(&s, A, &M)) return fail("prec_create");
	if(b200_prec_compute(M, info)) return fail("compute");
	if(b200_prec_apply_host(M, r, z1)) return fail("apply");

	/* 2. the same through the PCSHELL core: "-blasted_pc_type ilu0 -blasted_async_sweeps -1,-1" */
	b200_shell_list list = b200_shell_list_new();
	b200_shell_list_append(&list, b200_shell_node_new());
	b200_shell_options o;
	memset(&o, 0, sizeof(o));
	strcpy(o.pc_type, "ilu0");
	o.async_sweeps[0] = B200_SEQUENTIAL_SYMBOL; o.async_sweeps[1] = B200_SEQUENTIAL_SYMBOL;
	strcpy(o.fact_init_type, "init_original"); strcpy(o.apply_init_type, "init_zero");
	o.thread_chunk_size = 128;
	if(b200_shell_set_options(list.ctxlist, &o)) return fail("set_options");
	if(b200_shell_setup(list.ctxlist, BS, NB, ia, ja, a, diag)) return fail("shell_setup");
	if(b200_shell_apply(list.ctxlist, r, z2)) return fail("shell_apply");
	b200_shell_total_times(&list);

	double d = 0, nz = 0;
	for(int i = 0; i < n; i++) { d = fmax(d, fabs(z1[i] - z2[i])); nz = fmax(nz, fabs(z1[i])); }
	printf("n = %d, |z| = %.6e, |z_abi - z_shell| = %.3e, factor %.3f ms, apply %.3f ms\n", n, nz, d,
	       1e3*list.factorwalltime, 1e3*list.applywalltime);

	b200_shell_list_destroy(&list);
	b200_prec_destroy(M); b200_mat_destroy(A);
	free(ia); free(ja); free(diag); free(a); free(r); free(z1); free(z2);
	return (d == 0.0 && nz > 0.0) ? 0 : 1;
}
