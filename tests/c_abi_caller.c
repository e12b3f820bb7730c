/* A plain-C caller of libg4r.so: what a C / cgo / JNI binding would do on the host side.
 * Built and run by tests/test_host_logic.py::test_c_caller_links_against_the_abi with gcc -std=c99; it only touches entry
 * points that need no GPU (version, workspace size, the C++ schedule builder = gru4rec.py:594-651, ownership arithmetic),
 * then checks that creating a handle without a device fails with a message instead of crashing. */
#include <stdio.h>
#include <string.h>
#include "g4r.h"

int main(void) {
  g4r_config cfg;
  size_t bytes = 0;
  /* four sessions of lengths 3, 2, 4, 2 */
  const int64_t items[11] = {5, 6, 7, 1, 2, 9, 8, 7, 6, 3, 4};
  const int32_t offs[5] = {0, 3, 5, 9, 11};
  g4r_schedule* sched = NULL;
  g4r_handle* h = NULL;
  int32_t X[64], Y[64], M[16], slots[64];
  uint8_t F[64];
  int64_t n, k;
  if (g4r_version() < 100) return 1;
  memset(&cfg, 0, sizeof cfg);
  cfg.n_items = 1000; cfg.n_layers = 1; cfg.layers[0] = 100; cfg.batch_size = 32; cfg.loss = G4R_LOSS_BPR_MAX;
  cfg.final_act = G4R_ACT_ELU; cfg.final_act_p1 = 0.5f; cfg.hidden_act = G4R_ACT_TANH; cfg.learning_rate = 0.1f;
  cfg.n_sample = 2048; cfg.sample_alpha = 0.75f; cfg.bpreg = 1.0f; cfg.adapt = G4R_ADAPT_ADAGRAD; cfg.sample_store = 1 << 20;
  cfg.world_size = 1;
  if (g4r_workspace_bytes(&cfg, &bytes) != G4R_OK || bytes == 0) return 2;
  if (g4r_schedule_build(items, 11, offs, 4, NULL, 2, 0, 0, &sched) != G4R_OK) return 3;
  n = g4r_schedule_steps(sched);
  if (n <= 0 || n > 16 || g4r_schedule_events(sched) <= 0) return 4;
  if (g4r_schedule_export(sched, X, Y, F, M, slots) != G4R_OK) return 5;
  if (X[0] != 5 || Y[0] != 6 || X[1] != 1 || Y[1] != 2 || M[0] != 2) return 6;      /* first mini-batch: sessions 0 and 1 */
  for (k = 0; k < n; k++) printf("step %d: M=%d X=[%d,%d] Y=[%d,%d] reset=[%d,%d]\n", (int)k, (int)M[k], (int)X[2 * k], (int)X[2 * k + 1],
                                  (int)Y[2 * k], (int)Y[2 * k + 1], F[2 * k] & 1, F[2 * k + 1] & 1);
  g4r_schedule_free(sched);
  if (g4r_schedule_build(items, 11, offs, 4, NULL, 8, 0, 0, &sched) != G4R_ERR_INDEX) return 7;   /* fewer sessions than lanes */
  if (g4r_mg_owner(37482, 8) != 37482 % 8 || g4r_mg_local_row(37482, 8) != 37482 / 8) return 8;
  /* no device in this process: a loud failure with a message, never a CPU fallback */
  if (g4r_create(&cfg, NULL, 0, &h) == G4R_OK) { printf("device present: handle created\n"); g4r_destroy(h); }
  else printf("g4r_create: %s\n", g4r_last_error(NULL));
  printf("c caller ok: workspace %zu bytes, %d steps\n", bytes, (int)n);
  return 0;
}
