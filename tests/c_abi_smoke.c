/* The header must be valid C and the library linkable from C: no C++ or torch types in the ABI.
 * Host-only calls (no GPU needed): version, workspace queries, sparse index work, error codes. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "b200_lanczos.h"

int main(void) {
  if (strstr(bl_version(), "sm_100a") == NULL) return 1;
  if (bl_arnoldi_workspace_bytes(1000, 10, BL_F32) == 0) return 2;
  if (bl_lanczos3_workspace_bytes(1000, 10, BL_F64) == 0) return 3;
  int32_t row[5] = {0, 1, 2, 2, 0}, col[5] = {0, 1, 2, 0, 2};
  bl_operator_t* op = NULL;
  if (bl_op_sparse_create(3, 3, 5, row, col, &op) != BL_OK) return 4;
  int32_t row_ptr[4], col_idx[5], perm[5];
  if (bl_op_sparse_export_csr(op, row_ptr, col_idx, perm) != BL_OK) return 5;
  /* rows: {0:(0,2)}, {1:(1)}, {2:(0,2)} ; perm = COO positions sorted by (row, col) */
  const int32_t want_ptr[4] = {0, 2, 3, 5}, want_col[5] = {0, 2, 1, 0, 2}, want_perm[5] = {0, 4, 1, 3, 2};
  if (memcmp(row_ptr, want_ptr, sizeof want_ptr) || memcmp(col_idx, want_col, sizeof want_col) ||
      memcmp(perm, want_perm, sizeof want_perm))
    return 6;
  int np = 0;
  int64_t numel = 0, n = 0;
  if (bl_op_num_params(op, &np) != BL_OK || np != 1) return 7;
  if (bl_op_param_size(op, 0, &numel) != BL_OK || numel != 5) return 8;
  if (bl_op_size(op, &n) != BL_OK || n != 3) return 9;
  if (bl_op_destroy(op) != BL_OK) return 10;
  int32_t bad_row[1] = {7}, bad_col[1] = {0};
  if (bl_op_sparse_create(3, 3, 1, bad_row, bad_col, &op) != BL_EINVAL) return 11;
  if (strstr(bl_last_error(), "out of range") == NULL) return 12;
  int count = -1;
  if (bl_device_count(&count) != BL_OK || count < 0) return 13;
  printf("c abi ok (%s, %d device(s))\n", bl_version(), count);
  return 0;
}
