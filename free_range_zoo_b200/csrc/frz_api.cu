// Library-wide entry points of libfrz.so: version, thread-local error string, control-block initialisation.
#include <cstdarg>
#include <cstdio>

#include "frz_common.cuh"

namespace frz {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list args;
  va_start(args, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, args);
  va_end(args);
}

int check_launch(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) return FRZ_OK;
  set_error("%s: %s", what, cudaGetErrorString(err));
  return FRZ_ERR_CUDA;
}

namespace {
__global__ void control_init_kernel(FrzControl* control, uint64_t seed) {
  control->seed = seed;
  control->step = 0;
  control->ctas_done = 0;
  control->alive_acc = 0;
  control->alive = 3u;
  control->error_word = 0;
  control->agents_with_tasks_acc = 0;
  control->agents_with_tasks = 0;
}
}  // namespace

}  // namespace frz

extern "C" {

int frz_version(void) { return FRZ_ABI_VERSION; }

const char* frz_last_error(void) { return frz::g_error; }

int frz_control_init(FrzControl* control, uint64_t seed, void* stream) {
  if (control == nullptr) {
    frz::set_error("frz_control_init: control is NULL");
    return FRZ_ERR_NULL;
  }
  frz::control_init_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(control, seed);
  return frz::check_launch("control_init_kernel");
}

}  // extern "C"
