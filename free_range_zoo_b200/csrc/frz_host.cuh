// Pipelined host-buffer step shared by the three domains (frz_<domain>_step_host, include/frz.h).
//
// The caller's actions are in page-locked host memory and it wants rewards / done flags back in host memory.  Done
// naively that is upload -> kernel -> download, each waiting for the previous one, and on a PCIe-attached B200 the two
// copies cost more than the fused step itself.  Here the batch is cut into contiguous slices of environments; slice i
// runs [upload, step kernel, download] on its own stream, so the kernel of slice i overlaps the upload of slice i+1
// and the download of slice i-1 (the link is full duplex), and the kernels of neighbouring slices fill each other's
// tails.  Every slice has its own control block (seed / step counter / published flags): they are copies of the main
// block when the step starts and are folded back into it when the last slice has finished, so the flags the next
// launch reads ("every environment is done", "agent a has a task somewhere") stay batch-wide like the reference's.
#pragma once

#include "frz_common.cuh"

namespace frz {

// slices start on multiples of this many environments: keeps every per-environment array slice 16-byte aligned (the
// cybersecurity step moves its tiles with bulk copies) and the slices' warps fully populated
constexpr int kSliceAlignment = 1024;

struct HostArrays {  // device side of what the pipeline moves: [B, A, 2] actions in, [B, A] rewards and [B] flags out
  const int32_t* actions;
  const float* rewards;
  const uint8_t* terminated;
  const uint8_t* truncated;
  FrzControl* control;
  int agents;
};

// launch_slice(first_env, env_count, control_block, stream) enqueues the domain's step kernel for one slice
template <class LaunchSlice>
int run_host_pipeline(const char* what, const FrzHostStep* host, const HostArrays& device, int B, cudaStream_t main_stream,
                      LaunchSlice&& launch_slice) {
  if (host == nullptr || host->actions == nullptr || host->rewards == nullptr || host->terminated == nullptr ||
      host->truncated == nullptr || host->chunk_controls == nullptr || host->streams == nullptr) {
    set_error("%s: NULL host buffers / control blocks / streams", what);
    return FRZ_ERR_NULL;
  }
  if (host->chunks < 1 || host->chunks > FRZ_MAX_CHUNKS) {
    set_error("%s: chunks=%d outside [1, %d]", what, host->chunks, FRZ_MAX_CHUNKS);
    return FRZ_ERR_SHAPE;
  }
  cudaEvent_t* const events = pipeline_events();
  if (events == nullptr) return check_launch("pipeline events");

  int per_slice = (B + host->chunks - 1) / host->chunks;
  per_slice = (per_slice + kSliceAlignment - 1) / kSliceAlignment * kSliceAlignment;
  const int slices = (B + per_slice - 1) / per_slice;
  const size_t A = size_t(device.agents);

  int status = control_broadcast(device.control, host->chunk_controls, slices, main_stream);
  if (status != FRZ_OK) return status;
  cudaEventRecord(events[0], main_stream);
  for (int i = 0; i < slices; ++i) {
    const int first = i * per_slice, count = (B - first < per_slice) ? B - first : per_slice;
    cudaStream_t stream = static_cast<cudaStream_t>(host->streams[i]);
    cudaStreamWaitEvent(stream, events[0], 0);
    cudaMemcpyAsync(const_cast<int32_t*>(device.actions) + size_t(first) * A * 2, host->actions + size_t(first) * A * 2,
                    size_t(count) * A * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
    status = launch_slice(first, count, host->chunk_controls + i, stream);
    if (status != FRZ_OK) return status;
    cudaMemcpyAsync(host->rewards + size_t(first) * A, device.rewards + size_t(first) * A, size_t(count) * A * sizeof(float),
                    cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(host->terminated + first, device.terminated + first, size_t(count), cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(host->truncated + first, device.truncated + first, size_t(count), cudaMemcpyDeviceToHost, stream);
    cudaEventRecord(events[1 + i], stream);
  }
  for (int i = 0; i < slices; ++i) cudaStreamWaitEvent(main_stream, events[1 + i], 0);
  status = control_merge(device.control, host->chunk_controls, slices, main_stream);
  if (status != FRZ_OK) return status;
  return check_launch(what);
}

}  // namespace frz
