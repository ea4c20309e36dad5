/* DECLARATION-ONLY stand-in for the PETSc headers used by examples/petsc_glue/blasted_b200_petsc.c.
 *
 * PETSc is not in this image.  This header declares - with PETSc's own published signatures (3.8 to
 * 3.19, the range the reference supports: README.md:16 "PETSc 3.8 or above") - exactly the types,
 * functions and macros the glue uses, so that the glue is type-checked and compiled here
 * (tests/test_petsc_glue_cpu.py) instead of living as text in a document.  Nothing is implemented:
 * on a box with PETSc, compile the same source against the real <petscksp.h> and the private
 * headers named below.
 */
#ifndef B200_PETSC_STUB_H
#define B200_PETSC_STUB_H
#include <stddef.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_ERR_SUP 56
#define PETSC_ERR_LIB 76
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_WRONGSTATE 73
typedef const char *MatType;
typedef const char *PCType;
#define PCBJACOBI "bjacobi"
#define PCASM "asm"
#define PCSHELL "shell"
#define PCMG "mg"
#define PCGAMG "gamg"
#define PCKSP "ksp"
#define MATSEQAIJ "seqaij"
#define MATSEQBAIJ "seqbaij"
#define MATBAIJ "baij"
#define MATMPIBAIJ "mpibaij"
#define MATSEQAIJCUSPARSE "seqaijcusparse"

typedef struct _p_PetscObject *PetscObject;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_Vec *Vec;
/* <petsc/private/matimpl.h>: the glue reads Mat::data (src/blasted_petsc.cpp:278-279) */
struct _p_Mat { void *data; };
typedef struct _p_Mat *Mat;
/* <../src/mat/impls/aij/seq/aij.h>, <../src/mat/impls/baij/seq/baij.h>: i, j, a, diag (PETSc-side
 * facts relied upon: SURVEY.md appendix C) */
typedef struct { PetscInt *i, *j, *diag; PetscScalar *a; } Mat_SeqAIJ;
typedef struct { PetscInt *i, *j, *diag; PetscScalar *a; } Mat_SeqBAIJ;
typedef enum { PCRICHARDSON_CONVERGED_RTOL = 2, PCRICHARDSON_CONVERGED_ATOL = 3,
               PCRICHARDSON_CONVERGED_ITS = 4, PCRICHARDSON_DIVERGED_DTOL = -4 } PCRichardsonConvergedReason;

PetscErrorCode PetscError(MPI_Comm, int, const char*, const char*, PetscErrorCode, int, const char*, ...);
#define CHKERRQ(ierr) do { if(ierr) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, ierr, 1, " "); } while(0)
#define SETERRQ(comm, code, msg) return PetscError(comm, __LINE__, __func__, __FILE__, code, 0, msg)
#define SETERRQ1(comm, code, msg, a) return PetscError(comm, __LINE__, __func__, __FILE__, code, 0, msg, a)

PetscErrorCode PetscObjectTypeCompare(PetscObject, const char[], PetscBool*);
PetscErrorCode PetscOptionsGetString(void*, const char[], const char[], char[], size_t, PetscBool*);
PetscErrorCode PetscOptionsGetIntArray(void*, const char[], const char[], PetscInt[], PetscInt*, PetscBool*);
PetscErrorCode PetscOptionsGetInt(void*, const char[], const char[], PetscInt*, PetscBool*);
PetscErrorCode PetscOptionsGetBool(void*, const char[], const char[], PetscBool*, PetscBool*);

PetscErrorCode KSPGetPC(KSP, PC*);
PetscErrorCode KSPSetUp(KSP);
PetscErrorCode KSPGetOperators(KSP, Mat*, Mat*);
PetscErrorCode PCSetUp(PC);
PetscErrorCode PCGetOperators(PC, Mat*, Mat*);
PetscErrorCode PCBJacobiGetSubKSP(PC, PetscInt*, PetscInt*, KSP*[]);
PetscErrorCode PCASMGetSubKSP(PC, PetscInt*, PetscInt*, KSP*[]);
PetscErrorCode PCMGGetLevels(PC, PetscInt*);
PetscErrorCode PCMGGetSmoother(PC, PetscInt, KSP*);
PetscErrorCode PCMGGetCoarseSolve(PC, KSP*);
PetscErrorCode PCKSPGetKSP(PC, KSP*);
PetscErrorCode PCShellGetContext(PC, void**);
PetscErrorCode PCShellSetContext(PC, void*);
PetscErrorCode PCShellSetName(PC, const char[]);
PetscErrorCode PCShellSetSetUp(PC, PetscErrorCode (*)(PC));
PetscErrorCode PCShellSetApply(PC, PetscErrorCode (*)(PC, Vec, Vec));
PetscErrorCode PCShellSetDestroy(PC, PetscErrorCode (*)(PC));
PetscErrorCode PCShellSetApplyRichardson(PC, PetscErrorCode (*)(PC, Vec, Vec, Vec, PetscReal, PetscReal,
                                         PetscReal, PetscInt, PetscBool, PetscInt*,
                                         PCRichardsonConvergedReason*));

PetscErrorCode MatGetBlockSize(Mat, PetscInt*);
PetscErrorCode MatGetType(Mat, MatType*);
PetscErrorCode MatGetLocalSize(Mat, PetscInt*, PetscInt*);
PetscErrorCode MatMissingDiagonal(Mat, PetscBool*, PetscInt*);

PetscErrorCode VecGetType(Vec, const char**);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar**);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar**);
PetscErrorCode VecGetArray(Vec, PetscScalar**);
PetscErrorCode VecRestoreArray(Vec, PetscScalar**);
/* device vectors (PETSc built --with-cuda): zero-copy access to the device arrays */
PetscErrorCode VecCUDAGetArrayRead(Vec, const PetscScalar**);
PetscErrorCode VecCUDARestoreArrayRead(Vec, const PetscScalar**);
PetscErrorCode VecCUDAGetArray(Vec, PetscScalar**);
PetscErrorCode VecCUDARestoreArray(Vec, PetscScalar**);
PetscErrorCode VecCUDAGetArrayWrite(Vec, PetscScalar**);
PetscErrorCode VecCUDARestoreArrayWrite(Vec, PetscScalar**);
#endif
