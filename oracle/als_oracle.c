/*
 * als_oracle.c — CPU restatement of the reference worker's per-portion ALS update
 * and RMSE accumulation.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as
 * the checker / the timed CPU arm — never by the product path.
 *
 * PARITY UNPINNED: the reference has no tests, golden vectors or fixtures
 * (package.json:30) and its arithmetic lives in the un-vendored third-party
 * modules vectorious-plus ^4.3.16 -> nblas-plus (package.json:26, README.md:15),
 * i.e. CBLAS sgemm/sgemv/sdot + LAPACK sgesv.  This file restates the published
 * semantics of those routines at the reference's own call sites:
 *
 *   portion header parse ........ lib/emf/EmfWorker.js:176-199, 214-219
 *   row gather .................. lib/emf/EmfBase.js:537-555  (BLAS.BufCopy per rating)
 *   A = Y^T Y  (sgemm T,N) ...... lib/emf/EmfWorker.js:231-232
 *   A += (lambda*n) I ........... lib/emf/EmfWorker.js:233-235
 *   b = Y^T r ................... lib/emf/EmfWorker.js:238-245
 *   solve A x = b (gesv) ........ lib/emf/EmfWorker.js:246
 *   S[rowId,:] = x .............. lib/emf/EmfWorker.js:221-224, 247; EmfBase.js:518-532
 *   RMSE portion ................ lib/emf/EmfWorker.js:266-315; predict EmfBase.js:815-827
 *
 * Two arithmetic modes, both compiled for float (O32) and double (O64):
 *   - portable loops (default): sequential accumulation in the element type, LU with
 *     partial pivoting (the textbook sgesv algorithm);
 *   - OpenBLAS (oracle_set_blas): the same routine sequence the reference issues,
 *     through the cblas/LAPACK symbols of a dlopen'ed library — this is the
 *     multi-threaded CPU arm bench.py times.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- optional BLAS/LAPACK backend ------------------------------------------- */
enum { CblasRowMajor = 101, CblasNoTrans = 111, CblasTrans = 112 };
typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemm_fn)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sgemv_fn)(int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemv_fn)(int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef float (*sdot_fn)(int, const float*, int, const float*, int);
typedef double (*ddot_fn)(int, const double*, int, const double*, int);
typedef void (*sgesv_fn)(const int*, const int*, float*, const int*, int*, float*, const int*, int*);
typedef void (*dgesv_fn)(const int*, const int*, double*, const int*, int*, double*, const int*, int*);
typedef void (*setthreads_fn)(int);

static struct {
  void* handle;
  sgemm_fn sgemm; dgemm_fn dgemm; sgemv_fn sgemv; dgemv_fn dgemv;
  sdot_fn sdot; ddot_fn ddot; sgesv_fn sgesv; dgesv_fn dgesv; setthreads_fn set_threads;
} g_blas;

static void* sym2(void* h, const char* a, const char* b) {
  void* p = dlsym(h, a);
  return p ? p : dlsym(h, b);
}

/* path: a shared library exporting cblas_* / *gesv_ (optionally with the scipy_ prefix).
 * Returns 0 on success.  threads <= 0 keeps the library default. */
int oracle_set_blas(const char* path, int threads) {
  memset(&g_blas, 0, sizeof(g_blas));
  if (!path || !*path) return 0;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  g_blas.sgemm = (sgemm_fn)sym2(h, "scipy_cblas_sgemm", "cblas_sgemm");
  g_blas.dgemm = (dgemm_fn)sym2(h, "scipy_cblas_dgemm", "cblas_dgemm");
  g_blas.sgemv = (sgemv_fn)sym2(h, "scipy_cblas_sgemv", "cblas_sgemv");
  g_blas.dgemv = (dgemv_fn)sym2(h, "scipy_cblas_dgemv", "cblas_dgemv");
  g_blas.sdot = (sdot_fn)sym2(h, "scipy_cblas_sdot", "cblas_sdot");
  g_blas.ddot = (ddot_fn)sym2(h, "scipy_cblas_ddot", "cblas_ddot");
  g_blas.sgesv = (sgesv_fn)sym2(h, "scipy_sgesv_", "sgesv_");
  g_blas.dgesv = (dgesv_fn)sym2(h, "scipy_dgesv_", "dgesv_");
  g_blas.set_threads = (setthreads_fn)sym2(h, "scipy_openblas_set_num_threads", "openblas_set_num_threads");
  if (!g_blas.sgemm || !g_blas.dgemm || !g_blas.sgemv || !g_blas.dgemv || !g_blas.sdot ||
      !g_blas.ddot || !g_blas.sgesv || !g_blas.dgesv) {
    memset(&g_blas, 0, sizeof(g_blas));
    return 2;
  }
  g_blas.handle = h;
  if (threads > 0 && g_blas.set_threads) g_blas.set_threads(threads);
  return 0;
}
int oracle_has_blas(void) { return g_blas.handle != NULL; }

/* ---- generic body, instantiated for float and double ------------------------- */
#define T float
#define SUF f32
#define GEMM g_blas.sgemm
#define GEMV g_blas.sgemv
#define DOT g_blas.sdot
#define GESV g_blas.sgesv
#include "als_oracle_body.inc"
#undef T
#undef SUF
#undef GEMM
#undef GEMV
#undef DOT
#undef GESV

#define T double
#define SUF f64
#define GEMM g_blas.dgemm
#define GEMV g_blas.dgemv
#define DOT g_blas.ddot
#define GESV g_blas.dgesv
#include "als_oracle_body.inc"
