// Pipelined host-buffer step shared by the three domains (frz_<domain>_step_host, include/frz.h).
//
// The caller's actions are in page-locked host memory and it wants rewards / done flags back in host memory.  Done
// naively that is upload -> kernel -> download, each waiting for the previous one, and on a PCIe-attached B200 the two
// copies cost more than the fused step itself (65 536 wildfire environments: 105 us up, 105 us kernel, 60 us down).
// Here the batch is cut into contiguous slices of environments:
//   streams[0]      uploads the slices' actions back to back (one DMA queue, no gaps) and signals each one,
//   streams[i]      steps slice i as soon as its actions have arrived and downloads the slice's rewards,
//   the main stream steps slice 0, waits for every slice, downloads the done flags and folds the control blocks.
// So the kernel of slice i overlaps the upload of slice i+1 and the download of slice i-1 (the link is full duplex),
// and the kernels of neighbouring slices fill each other's tails.  The first slice is half as long as the others: the
// kernels cannot start before it has arrived.  Every slice has its own control block (seed / step counter / published
// flags): they are copies of the main block when the step starts and are folded back into it at the end, so the flags
// the next launch reads ("every environment is done", "agent a has a task somewhere") stay batch-wide like the
// reference's.
//
// Measured alternative, not kept: letting the kernel read the actions from / write the results to the page-locked host
// buffers in place (mapped memory) -- the 4- and 1-byte accesses become tiny PCIe transactions and the step gets slower
// than with the DMA copies (profiles/README.md).
#pragma once

#include "frz_common.cuh"

namespace frz {

// slices start on multiples of this many environments: keeps every per-environment array slice 16-byte aligned (the
// cybersecurity step moves its tiles with bulk copies) and the slices' warps fully populated
constexpr int kSliceAlignment = 1024;

// Slice boundaries of a batch of B environments cut into at most `chunks` slices: bounds[0] = 0 < ... < bounds[n] = B,
// every inner boundary a multiple of kSliceAlignment, the first slice half as long as the others.  Returns n.
inline int slice_bounds(int B, int chunks, int* bounds) {
  int slices = 0;
  const long long halves = 2LL * chunks - 1;  // the batch in units of half a regular slice
  bounds[0] = 0;
  for (int i = 1; i <= chunks; ++i) {
    long long end = (long long)B * (2LL * i - 1) / halves;
    end = (end + kSliceAlignment - 1) / kSliceAlignment * kSliceAlignment;
    if (end > B || i == chunks) end = B;
    if (end > bounds[slices]) bounds[++slices] = int(end);
  }
  return slices;
}

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
  cudaEvent_t const started = events[0];
  cudaEvent_t* const uploaded = events + 1;                   // [slice]
  cudaEvent_t* const finished = events + 1 + FRZ_MAX_CHUNKS;  // [slice]

  int bounds[FRZ_MAX_CHUNKS + 1];
  const int slices = slice_bounds(B, host->chunks, bounds);
  const size_t A = size_t(device.agents);

  if (slices == 1) {  // nothing to overlap: upload, step and download on the caller's stream, on the main control block
    cudaMemcpyAsync(const_cast<int32_t*>(device.actions), host->actions, size_t(B) * A * 2 * sizeof(int32_t),
                    cudaMemcpyHostToDevice, main_stream);
    const int launched = launch_slice(0, B, device.control, main_stream);
    if (launched != FRZ_OK) return launched;
    cudaMemcpyAsync(host->rewards, device.rewards, size_t(B) * A * sizeof(float), cudaMemcpyDeviceToHost, main_stream);
    cudaMemcpyAsync(host->terminated, device.terminated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
    cudaMemcpyAsync(host->truncated, device.truncated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
    return check_launch(what);
  }

  int status = control_broadcast(device.control, host->chunk_controls, slices, main_stream);
  if (status != FRZ_OK) return status;
  cudaEventRecord(started, main_stream);
  cudaStream_t upload_stream = static_cast<cudaStream_t>(host->streams[0]);
  cudaStreamWaitEvent(upload_stream, started, 0);
  for (int i = 0; i < slices; ++i) {
    const int first = bounds[i], count = bounds[i + 1] - first;
    cudaMemcpyAsync(const_cast<int32_t*>(device.actions) + size_t(first) * A * 2, host->actions + size_t(first) * A * 2,
                    size_t(count) * A * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, upload_stream);
    cudaEventRecord(uploaded[i], upload_stream);
  }
  for (int i = 0; i < slices; ++i) {
    const int first = bounds[i], count = bounds[i + 1] - first;
    // (slice 0 runs on the main stream: streams[0] is busy uploading the later slices)
    cudaStream_t stream = (i == 0) ? main_stream : static_cast<cudaStream_t>(host->streams[i]);
    cudaStreamWaitEvent(stream, uploaded[i], 0);
    status = launch_slice(first, count, host->chunk_controls + i, stream);
    if (status != FRZ_OK) return status;
    cudaMemcpyAsync(host->rewards + size_t(first) * A, device.rewards + size_t(first) * A, size_t(count) * A * sizeof(float),
                    cudaMemcpyDeviceToHost, stream);
    if (i > 0) cudaEventRecord(finished[i], stream);
  }
  for (int i = 1; i < slices; ++i) cudaStreamWaitEvent(main_stream, finished[i], 0);
  // the done flags of the whole batch: one download when the two arrays are adjacent on both sides
  if (device.truncated == device.terminated + B && host->truncated == host->terminated + B) {
    cudaMemcpyAsync(host->terminated, device.terminated, 2 * size_t(B), cudaMemcpyDeviceToHost, main_stream);
  } else {
    cudaMemcpyAsync(host->terminated, device.terminated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
    cudaMemcpyAsync(host->truncated, device.truncated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
  }
  status = control_merge(device.control, host->chunk_controls, slices, main_stream);
  if (status != FRZ_OK) return status;
  return check_launch(what);
}

}  // namespace frz
