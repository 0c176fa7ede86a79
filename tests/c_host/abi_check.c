/* A non-Python host of libtheatergen.so: plain C99, no CUDA headers.  Checks that include/theatergen.h is valid C, that
 * the library loads with dlopen, that every entry point named on the command line resolves, and that a call with bad
 * arguments fails with a negative code and a message instead of crashing (tests/test_abi_cpu.py builds and runs it).
 *
 *     gcc -std=c99 -Wall -Wextra -pedantic -I include tests/c_host/abi_check.c -ldl -o abi_check
 *     ./abi_check theatergen_b200/libtheatergen.so tg_gemm tg_conv3x3 ...
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "theatergen.h"

/* function-pointer types straight from the header's declarations (no link-time reference to the library) */
typedef int (*version_fn)(void);
typedef const char* (*error_fn)(void);
typedef int (*gemm_fn)(const tg_half*, int, int, const tg_half*, int, const tg_half*, int, int, int, const tg_half*,
                       const tg_half*, int, int, const tg_half*, int, int, float, tg_half*, int, void*);
typedef int (*cross_scores_fn)(const tg_half*, int, const tg_half*, int, int, int, int, int, int, const float*, tg_half*,
                               int, void*);

/* compile-time check that the typedefs above match the header: assigning the declared functions' addresses would need the
 * library at link time, so compare the types instead (GNU C) */
#if defined(__GNUC__)
static void type_checks(void) {
  gemm_fn g = (__typeof__(&tg_gemm))0;
  cross_scores_fn c = (__typeof__(&tg_cross_scores))0;
  version_fn v = (__typeof__(&tg_version))0;
  error_fn e = (__typeof__(&tg_last_error))0;
  (void)g; (void)c; (void)v; (void)e;
}
#endif

static void* sym(void* h, const char* name) {
  void* p = dlsym(h, name);
  if (!p) fprintf(stderr, "missing symbol %s\n", name);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s libtheatergen.so [symbol ...]\n", argv[0]);
    return 2;
  }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 1;
  }
#if defined(__GNUC__)
  type_checks();
#endif
  int missing = 0;
  for (int i = 2; i < argc; ++i)
    if (!sym(h, argv[i])) ++missing;
  if (missing) return 1;
  version_fn version;
  error_fn last_error;
  gemm_fn gemm;
  *(void**)(&version) = sym(h, "tg_version");
  *(void**)(&last_error) = sym(h, "tg_last_error");
  *(void**)(&gemm) = sym(h, "tg_gemm");
  if (!version || !last_error || !gemm) return 1;
  if (version() < 100) {
    fprintf(stderr, "unexpected version %d\n", version());
    return 1;
  }
  /* null operands: must be refused before anything touches the device */
  int rc = gemm(NULL, 64, 64, NULL, 0, NULL, 128, 64, 64, NULL, NULL, 1, 0, NULL, 0, TG_ACT_NONE, 1.0f, NULL, 64, NULL);
  if (rc >= 0 || strlen(last_error()) == 0) {
    fprintf(stderr, "tg_gemm(NULL, ...) returned %d, message '%s'\n", rc, last_error());
    return 1;
  }
  printf("ok version=%d symbols=%d refused: %s\n", version(), argc - 2, last_error());
  dlclose(h);
  return 0;
}
