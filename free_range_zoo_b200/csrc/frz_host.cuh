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
  // the caller's Frz<Domain>Params / Frz<Domain>Buffers (bytes): part of what a cached pipeline graph depends on
  const void* params;
  size_t params_bytes;
  const void* io;
  size_t io_bytes;
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
  const bool packed = host->action_format == FRZ_HOST_ACTIONS_I16 || host->action_format == FRZ_HOST_ACTIONS_I8;
  const size_t packed_bytes = host->action_format == FRZ_HOST_ACTIONS_I8 ? 1 : 2;  // per action word
  if (host->action_format != FRZ_HOST_ACTIONS_I32 && !packed) {
    set_error("%s: unknown action_format %d", what, host->action_format);
    return FRZ_ERR_UNSUPPORTED;
  }
  if (packed && host->packed_actions == nullptr) {
    set_error("%s: FRZ_HOST_ACTIONS_I16 / _I8 need the packed_actions device scratch", what);
    return FRZ_ERR_NULL;
  }
  cudaEvent_t* const events = pipeline_events(host->pipeline);
  if (events == nullptr) return FRZ_ERR_CUDA;
  cudaEvent_t const started = events[0];
  cudaEvent_t* const uploaded = events + 1;                   // [slice]
  cudaEvent_t* const finished = events + 1 + FRZ_MAX_CHUNKS;  // [slice]

  int bounds[FRZ_MAX_CHUNKS + 1];
  const int slices = slice_bounds(B, host->chunks, bounds);
  const size_t A = size_t(device.agents);
  // upload the actions of environments [first, first + count) on `stream`; int16 / int8 pairs are widened on the device
  const auto upload = [&](int first, int count, cudaStream_t stream, bool widen_here) -> int {
    const size_t at = size_t(first) * A * 2, words = size_t(count) * A * 2;
    int32_t* const actions = const_cast<int32_t*>(device.actions) + at;
    if (!packed) {
      cudaMemcpyAsync(actions, static_cast<const int32_t*>(host->actions) + at, words * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
      return FRZ_OK;
    }
    char* const staged = reinterpret_cast<char*>(host->packed_actions) + at * packed_bytes;
    cudaMemcpyAsync(staged, static_cast<const char*>(host->actions) + at * packed_bytes, words * packed_bytes,
                    cudaMemcpyHostToDevice, stream);
    // (the pipelined path widens on the slice's own stream instead, once the slice has arrived)
    return widen_here ? widen_actions(staged, int(packed_bytes), actions, words, stream) : FRZ_OK;
  };

  if (slices == 1) {  // nothing to overlap: upload, step and download on the caller's stream, on the main control block
    const int uploaded_ok = upload(0, B, main_stream, true);
    if (uploaded_ok != FRZ_OK) return uploaded_ok;
    const int launched = launch_slice(0, B, device.control, main_stream);
    if (launched != FRZ_OK) return launched;
    cudaMemcpyAsync(host->rewards, device.rewards, size_t(B) * A * sizeof(float), cudaMemcpyDeviceToHost, main_stream);
    cudaMemcpyAsync(host->terminated, device.terminated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
    cudaMemcpyAsync(host->truncated, device.truncated, size_t(B), cudaMemcpyDeviceToHost, main_stream);
    return check_launch(what);
  }

  // ---- the pipelined step, enqueued on the caller's streams (directly, or once into a graph: see below)
  const auto enqueue = [&]() -> int {
    int status = control_broadcast(device.control, host->chunk_controls, slices, main_stream);
    if (status != FRZ_OK) return status;
    cudaEventRecord(started, main_stream);
    cudaStream_t upload_stream = static_cast<cudaStream_t>(host->streams[0]);
    cudaStreamWaitEvent(upload_stream, started, 0);
    for (int i = 0; i < slices; ++i) {
      upload(bounds[i], bounds[i + 1] - bounds[i], upload_stream, false);
      cudaEventRecord(uploaded[i], upload_stream);
    }
    int launched = 0;  // slices whose stream has work enqueued (all of them are joined below, also after an error)
    for (int i = 0; i < slices && status == FRZ_OK; ++i) {
      const int first = bounds[i], count = bounds[i + 1] - first;
      // (slice 0 runs on the main stream: streams[0] is busy uploading the later slices)
      cudaStream_t stream = (i == 0) ? main_stream : static_cast<cudaStream_t>(host->streams[i]);
      cudaStreamWaitEvent(stream, uploaded[i], 0);
      launched = i + 1;
      if (packed) status = widen_actions(reinterpret_cast<const char*>(host->packed_actions) + size_t(first) * A * 2 * packed_bytes,
                                         int(packed_bytes), const_cast<int32_t*>(device.actions) + size_t(first) * A * 2,
                                         size_t(count) * A * 2, stream);
      if (status == FRZ_OK) status = launch_slice(first, count, host->chunk_controls + i, stream);
      if (status == FRZ_OK)
        cudaMemcpyAsync(host->rewards + size_t(first) * A, device.rewards + size_t(first) * A, size_t(count) * A * sizeof(float),
                        cudaMemcpyDeviceToHost, stream);
      if (i > 0) cudaEventRecord(finished[i], stream);
    }
    // join: the main stream continues only when every slice stream -- and the upload stream -- has drained, so the
    // caller's synchronisation covers everything that was enqueued even when a launch failed half way
    for (int i = 1; i < launched; ++i) cudaStreamWaitEvent(main_stream, finished[i], 0);
    cudaStreamWaitEvent(main_stream, uploaded[slices - 1], 0);
    if (status != FRZ_OK) return status;
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
  };

  // With a pipeline handle the whole multi-stream step is captured once and replayed: one cudaGraphLaunch per step
  // instead of ~40 stream calls.  The graph depends on the device buffers, the host result buffers, the streams and the
  // parameter block -- any change re-captures -- and on the host action buffer, whose address is patched into the
  // upload nodes when the caller passes a different page-locked buffer.  A caller who is capturing the step into a
  // graph of its own gets the plain enqueue.
  FrzHostPipeline* const cache = host->pipeline;
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(main_stream, &capturing);
  if (cache == nullptr || capturing != cudaStreamCaptureStatusNone) return enqueue();

  std::vector<unsigned char> signature;
  const auto absorb = [&signature](const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    signature.insert(signature.end(), p, p + bytes);
  };
  absorb(device.params, device.params_bytes);
  absorb(device.io, device.io_bytes);
  absorb(&B, sizeof(B));
  absorb(&main_stream, sizeof(main_stream));
  absorb(&host->rewards, sizeof(host->rewards));
  absorb(&host->terminated, sizeof(host->terminated));
  absorb(&host->truncated, sizeof(host->truncated));
  absorb(&host->chunk_controls, sizeof(host->chunk_controls));
  absorb(&host->chunks, sizeof(host->chunks));
  absorb(&host->action_format, sizeof(host->action_format));
  absorb(&host->packed_actions, sizeof(host->packed_actions));
  absorb(host->streams, sizeof(void*) * size_t(host->chunks));

  if (cache->exec == nullptr || cache->signature != signature) {
    if (cache->exec != nullptr) cudaGraphExecDestroy(cache->exec);
    if (cache->graph != nullptr) cudaGraphDestroy(cache->graph);
    cache->exec = nullptr;
    cache->graph = nullptr;
    cache->action_copies.clear();
    if (cudaStreamBeginCapture(main_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return enqueue();  // (capture unavailable on this stream: run uncached)
    }
    const int enqueued = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t ended = cudaStreamEndCapture(main_stream, &graph);
    if (enqueued != FRZ_OK || ended != cudaSuccess || graph == nullptr) {
      if (graph != nullptr) cudaGraphDestroy(graph);
      if (enqueued != FRZ_OK) return enqueued;
      return check_launch("host pipeline capture");
    }
    // the upload nodes: memcpy nodes whose source lies in the host action buffer
    size_t node_count = 0;
    cudaGraphGetNodes(graph, nullptr, &node_count);
    std::vector<cudaGraphNode_t> nodes(node_count);
    cudaGraphGetNodes(graph, nodes.data(), &node_count);
    const size_t element = packed ? packed_bytes : sizeof(int32_t);
    const char* const base = static_cast<const char*>(host->actions);
    const size_t span = size_t(B) * A * 2 * element;
    for (cudaGraphNode_t node : nodes) {
      cudaGraphNodeType type;
      if (cudaGraphNodeGetType(node, &type) != cudaSuccess || type != cudaGraphNodeTypeMemcpy) continue;
      cudaMemcpy3DParms copy = {};
      if (cudaGraphMemcpyNodeGetParams(node, &copy) != cudaSuccess) continue;
      const char* const source = static_cast<const char*>(copy.srcPtr.ptr);
      if (source >= base && source < base + span)
        cache->action_copies.push_back({node, copy.dstPtr.ptr, size_t(source - base), copy.extent.width});
    }
    if (cudaGraphInstantiate(&cache->exec, graph, 0) != cudaSuccess) {
      cudaGraphDestroy(graph);
      cache->exec = nullptr;
      return check_launch("host pipeline instantiate");
    }
    cache->graph = graph;
    cache->signature = signature;
    cache->captured_actions = host->actions;
  } else if (cache->captured_actions != host->actions) {
    for (const FrzHostPipeline::ActionCopy& copy : cache->action_copies) {
      if (cudaGraphExecMemcpyNodeSetParams1D(cache->exec, copy.node, copy.dst,
                                             static_cast<const char*>(host->actions) + copy.offset, copy.bytes,
                                             cudaMemcpyHostToDevice) != cudaSuccess)
        return check_launch("host pipeline update");
    }
    cache->captured_actions = host->actions;
  }
  if (cudaGraphLaunch(cache->exec, main_stream) != cudaSuccess) return check_launch("host pipeline launch");
  return FRZ_OK;
}

}  // namespace frz
