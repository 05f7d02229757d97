// Library-wide entry points of libfrz.so: version, thread-local error string, control-block initialisation.
#include <cstdarg>
#include <cstdio>

#include "frz_common.cuh"
#include "frz_host.cuh"

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

// every chunk block starts the step as a copy of the main block (seed, step counter, published flags)
__global__ void control_broadcast_kernel(const FrzControl* main_block, FrzControl* chunk_blocks, int count) {
  const int i = threadIdx.x;
  if (i >= count) return;
  FrzControl block = *main_block;
  block.ctas_done = 0;
  block.alive_acc = 0;
  block.agents_with_tasks_acc = 0;
  block.error_word = 0;
  chunk_blocks[i] = block;
}

// the main block ends the step with what one launch over the whole batch would have published: the flags are ORs over
// the slices, the step counter is the slices' common one
__global__ void control_merge_kernel(FrzControl* main_block, const FrzControl* chunk_blocks, int count) {
  unsigned alive = 0, agents = 0, faults = 0;
  for (int i = 0; i < count; ++i) {
    alive |= chunk_blocks[i].alive;
    agents |= chunk_blocks[i].agents_with_tasks;
    faults |= chunk_blocks[i].error_word;
  }
  if (chunk_blocks[0].step != main_block->step) {  // the slices stepped (no "every environment is done" early-out)
    main_block->alive = alive;
    main_block->agents_with_tasks = agents;
    main_block->step = chunk_blocks[0].step;
  }
  main_block->error_word |= faults;
}
}  // namespace

int control_broadcast(FrzControl* main_block, FrzControl* chunk_blocks, int count, cudaStream_t stream) {
  control_broadcast_kernel<<<1, 32, 0, stream>>>(main_block, chunk_blocks, count);
  return check_launch("control_broadcast_kernel");
}

int control_merge(FrzControl* main_block, FrzControl* chunk_blocks, int count, cudaStream_t stream) {
  control_merge_kernel<<<1, 1, 0, stream>>>(main_block, chunk_blocks, count);
  return check_launch("control_merge_kernel");
}

cudaEvent_t* pipeline_events() {
  static thread_local cudaEvent_t events[1 + 2 * FRZ_MAX_CHUNKS];
  static thread_local int device_of_events = -1;
  int device = 0;
  cudaGetDevice(&device);
  if (device_of_events != device) {  // (events belong to the device they were created on)
    for (auto& event : events)
      if (cudaEventCreateWithFlags(&event, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    device_of_events = device;
  }
  return events;
}

}  // namespace frz

extern "C" {

int frz_version(void) { return FRZ_ABI_VERSION; }

const char* frz_last_error(void) { return frz::g_error; }

int frz_host_slices(int32_t parallel_envs, int32_t chunks, int32_t* bounds) {
  if (bounds == nullptr) {
    frz::set_error("frz_host_slices: bounds is NULL");
    return -FRZ_ERR_NULL;
  }
  if (parallel_envs <= 0 || chunks < 1 || chunks > FRZ_MAX_CHUNKS) {
    frz::set_error("frz_host_slices: parallel_envs=%d chunks=%d", parallel_envs, chunks);
    return -FRZ_ERR_SHAPE;
  }
  return frz::slice_bounds(parallel_envs, chunks, bounds);
}

int frz_control_init(FrzControl* control, uint64_t seed, void* stream) {
  if (control == nullptr) {
    frz::set_error("frz_control_init: control is NULL");
    return FRZ_ERR_NULL;
  }
  frz::control_init_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(control, seed);
  return frz::check_launch("control_init_kernel");
}

}  // extern "C"
