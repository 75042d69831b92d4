/* The C ABI used from plain C (C99, -pedantic): includes include/gcnb200.h, links libgcnb200.so and makes the
 * host-only calls a C / cgo / JNI binding would make first -- version, error string, the pre-split tap image of a
 * weight matrix and the operator image of a small path graph.  No GPU is touched.  tests/test_abi.py compiles and
 * runs it and compares the numbers with the ctypes binding. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gcnb200.h"

int main(void) {
  enum { M = 8, NNZ = 14 };
  int32_t rowptr[M + 1], col[NNZ];
  float val[NNZ];
  int n = 0, i;
  /* path graph 0-1-...-7, off-diagonal weights -0.5 (a rescaled normalised Laplacian has a zero diagonal) */
  for (i = 0; i < M; ++i) {
    rowptr[i] = n;
    if (i > 0) { col[n] = i - 1; val[n++] = -0.5f; }
    if (i < M - 1) { col[n] = i + 1; val[n++] = -0.5f; }
  }
  rowptr[M] = n;
  if (n != NNZ) return 2;
  printf("version %d\n", gcnb_version());
  printf("tap_bytes %lu\n", (unsigned long)gcnb_cheb_tap_image_bytes(32, 32, 5));
  printf("tap_bytes_unsupported %lu\n", (unsigned long)gcnb_cheb_tap_image_bytes(4, 32, 5));
  {
    size_t bytes = gcnb_cheb_image_bytes(rowptr, col, 4, M, NNZ, 16, 8, 3, 2, 0);
    unsigned char* img = (unsigned char*)calloc(bytes ? bytes : 1, 1);
    unsigned long sum = 0;
    size_t k;
    int rc = gcnb_cheb_image_build(rowptr, col, val, 4, M, NNZ, 16, 8, 3, 2, 0, img, bytes);
    for (k = 0; k < bytes; ++k) sum += img[k];
    printf("image_bytes %lu rc %d checksum %lu\n", (unsigned long)bytes, rc, sum);
    rc = gcnb_cheb_image_build(rowptr, col, val, 4, M, NNZ, 16, 8, 3, 2, 0, img, bytes + 16);
    printf("wrong_size rc %d error_set %d\n", rc, (int)(strlen(gcnb_last_error_string()) > 0));
    free(img);
  }
  return 0;
}
