/* A plain C99 client of include/fcn8s_b200.h: what a maintainer binding the library from C / cgo / JNI compiles.
 * Built and run by tests/test_host_cpu.py::test_header_is_valid_c_and_a_c_client_links (no GPU needed: only host
 * entry points and argument validation are exercised). */
#include <stdio.h>
#include <string.h>

#include "fcn8s_b200.h"

int main(void) {
  if (fcn8_version() != 100) return 1;
  /* CRC-32C check value of "123456789" (RFC 3720 appendix B.4) */
  if (fcn8_crc32c("123456789", 9, 0) != 0xE3069283u) return 2;
  /* label feed: four pixels, three classes */
  uint8_t onehot[4 * 3] = {0, 1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0}, ids[4];
  if (fcn8_pack_labels(onehot, 4, 3, ids, 1) != 0 || ids[0] != 1 || ids[1] != 0 || ids[2] != 2 || ids[3] != 1) return 3;
  onehot[0] = 1; /* two classes on pixel 0: not one-hot any more */
  if (fcn8_pack_labels(onehot, 4, 3, ids, 1) != 1) return 4;
  /* error convention: negative status + thread-local message */
  if (fcn8_pack_labels(NULL, 4, 3, ids, 1) >= 0 || !strlen(fcn8_last_error())) return 5;
  Fcn8ConvParams conv;
  memset(&conv, 0, sizeof conv);
  if (fcn8_conv_gemm(&conv, NULL, 0, NULL) >= 0) return 6; /* rejected before any CUDA call */
  Fcn8DeconvParams dec;
  memset(&dec, 0, sizeof dec);
  if (fcn8_deconv_loss(&dec, NULL) >= 0) return 7;
  printf("abi ok\n");
  return 0;
}
