/* A plain C consumer of the drop-in boundary: links libprl.so through its header only (no Python, no torch).
 * Built and run by tests/test_boundary.py::test_plain_c_client_links_and_calls on the CPU box: it checks the version, the
 * workspace-size queries and that a bad call comes back as a negative status with a message instead of crashing. */
#include <stdio.h>
#include <string.h>
#include "prl.h"

int main(void) {
  if (prl_version() < 100) { printf("bad version %d\n", (int)prl_version()); return 1; }
  if (prl_adamw_workspace_bytes() == 0 || prl_rowops_workspace_bytes(4096) == 0) { printf("bad workspace sizes\n"); return 2; }
  /* K = 20 is not a multiple of 8: rejected before any CUDA call; dummy non-NULL pointers are never dereferenced */
  char dummy[64];
  int rc = prl_gemm_ex(dummy, 20, 0, dummy, 24, 0, 16, 16, 20, dummy, 16, 0, 0, NULL, NULL, 0, 1.0f, NULL);
  if (rc >= 0) { printf("expected an error status, got %d\n", rc); return 3; }
  const char* msg = prl_last_error();
  if (!msg || !strstr(msg, "multiples of 8")) { printf("unexpected message: %s\n", msg ? msg : "(null)"); return 4; }
  printf("ok version=%d rc=%d msg=%s\n", (int)prl_version(), rc, msg);
  return 0;
}
