// abi_on_oracle.cpp — TEST INFRASTRUCTURE ONLY: the C-ABI entry points the C++ shim calls, answered by the CPU oracle.
//
// This is NOT a CPU fallback of the product (the product library, pl-svo_b200/csrc/libplsvo_b200.so, has none and is not
// involved here).  It exists so that the shim's packing / unpacking of the reference's own Frame / Feature / SE3 objects
// (pl-svo_b200/host/plsvo_shim.cpp compiled with -DPLSVO_SHIM_WITH_REFERENCE_HEADERS) can be checked on a machine
// without a GPU: `make -C oracle shimref-cpu` links shimref_harness.cpp + the shim + this file + plsvo_oracle.cpp into
// oracle/_ref/libplsvo_shimref_cpu.so, and tests/test_reference_tu_cpu.py requires that reference objects pushed through the
// shim come back exactly as the reference's own sparse_img_align.cpp / pose_optimizer.cpp leave them.
#include <cstdlib>
#include <cstring>

#include "../include/plsvo_b200.h"

extern "C" {
int plsvo_oracle_align_batch(const plsvo_align_batch*, const plsvo_align_params*, const plsvo_align_result*, int, int);
int plsvo_oracle_poseopt_batch(const plsvo_poseopt_batch*, const plsvo_poseopt_params*, const plsvo_poseopt_result*, int);
int plsvo_oracle_structopt_batch(const plsvo_structopt_batch*, const plsvo_structopt_result*, int);
int plsvo_oracle_match_direct_batch(const plsvo_match_batch*, const plsvo_match_result*, int);
int plsvo_oracle_seed_update_batch(const plsvo_seed_batch*, const plsvo_seed_result*, int);
int plsvo_oracle_line_seed_update_batch(const plsvo_line_seed_batch*, const plsvo_line_seed_result*, int);

struct plsvo_ctx {
  int unused;
};
static plsvo_ctx g_fake_ctx;
int plsvo_ctx_create(int, void*, plsvo_ctx** out) {
  *out = &g_fake_ctx;
  return PLSVO_OK;
}
void plsvo_ctx_destroy(plsvo_ctx*) {}
int plsvo_host_alloc(void** ptr, size_t bytes) {  // plain memory: nothing is copied to a device here
  *ptr = malloc(bytes);
  return *ptr ? PLSVO_OK : PLSVO_ERR_INVALID;
}
int plsvo_host_free(void* ptr) {
  free(ptr);
  return PLSVO_OK;
}
const char* plsvo_last_error(const plsvo_ctx*) { return "oracle-backed test adapter"; }
int plsvo_align_batch_run(plsvo_ctx*, const plsvo_align_batch* b, const plsvo_align_params* p, const plsvo_align_result* o) {
  return plsvo_oracle_align_batch(b, p, o, 1, 0);
}
int plsvo_poseopt_batch_run(plsvo_ctx*, const plsvo_poseopt_batch* b, const plsvo_poseopt_params* p, const plsvo_poseopt_result* o) {
  return plsvo_oracle_poseopt_batch(b, p, o, 1);
}
int plsvo_structopt_batch_run(plsvo_ctx*, const plsvo_structopt_batch* b, const plsvo_structopt_result* o) {
  return plsvo_oracle_structopt_batch(b, o, 1);
}
int plsvo_match_direct_batch_run(plsvo_ctx*, const plsvo_match_batch* b, const plsvo_match_result* o) {
  return plsvo_oracle_match_direct_batch(b, o, 4);
}
int plsvo_seed_update_batch_run(plsvo_ctx*, const plsvo_seed_batch* b, const plsvo_seed_result* o) {
  return plsvo_oracle_seed_update_batch(b, o, 4);
}
int plsvo_line_seed_update_batch_run(plsvo_ctx*, const plsvo_line_seed_batch* b, const plsvo_line_seed_result* o) {
  return plsvo_oracle_line_seed_update_batch(b, o, 4);
}
}
