/* Minimal DECLARATIONS-ONLY stand-in for R's Rinternals.h, written for this repository so that
 * tests/test_abi_host.py can type-check r/mclcar_b200_glue.c against include/mclcar_b200.h with
 * `gcc -fsyntax-only` (R itself is not installed in the build image).  Only what the glue uses;
 * signatures as documented in "Writing R Extensions".  Nothing here is linked or executed. */
#ifndef MCLB_R_STUB_RINTERNALS_H
#define MCLB_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
extern SEXP R_NilValue;
double* REAL(SEXP x);
int* INTEGER(SEXP x);
int LENGTH(SEXP x);
R_xlen_t XLENGTH(SEXP x);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_install(const char*);
SEXP Rf_setAttrib(SEXP vec, SEXP name, SEXP val);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
char* R_alloc(size_t n, int size);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#endif
