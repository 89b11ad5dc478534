// Multi-GPU helpers of the C-ABI (SURVEY §8b last bullet, §8e): symmetric buffers that the
// GPUs of one box can address directly over NVLink (CUDA peer access inside a process, CUDA IPC
// between the one-process-per-GPU ranks), the 1-row flow-map halo exchange and the gather of
// finished output rows, both as peer DMA writes followed by a system-scope release of a
// generation flag in the receiver's memory, and the matching device-side wait.
//
// The reference's only parallel construct is a process pool whose results `p.map` collects
// into one array (R:src/dynlab/diagnostics/_base_classes.py:139-148, :178-187); the row gather
// is the device-side equivalent of that collection step, the halo exchange exists because the
// stencil of R:lagrangian.py:50-55 (np.gradient) reads the neighbouring slab's boundary row.
//
// No collective library is involved: a copy engine moves the rows while the SMs keep
// integrating the next row chunk (the integrator never touches memory), and one warp publishes
// / polls the flags.
#include "common.cuh"

namespace {

struct FlagPtrs {
  unsigned long long* p[DLB_MAX_PEERS];
};

// flag value `v` becomes visible to the receiver only after everything this GPU wrote before
// (the peer copies are earlier in the same stream, i.e. complete)
__global__ void signal_flags_byval_kernel(const FlagPtrs f, int n, unsigned long long v) {
  const int i = threadIdx.x;
  if (i < n && f.p[i]) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[i]), "l"(v) : "memory");
  }
}

// One warp; lane i polls flag i until it reaches `min_value` (acquire at system scope: the rows
// the sender copied before its release are visible to every later kernel of this stream).
// Bounded: after `timeout_ns` the kernel gives up and reports through *timed_out, so a dead
// peer turns into an error on the host instead of a hung GPU.
__global__ void wait_flags_kernel(const unsigned long long* flags, int n,
                                  unsigned long long min_value, unsigned long long timeout_ns,
                                  int* timed_out) {
  const int i = threadIdx.x;
  bool late = false;
  if (i < n) {
    unsigned long long t0, now, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned ns = 64;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
      if (v >= min_value) break;
      __nanosleep(ns);
      if (ns < 2048) ns <<= 1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > timeout_ns) {
        late = true;
        break;
      }
    }
  }
  if (late && timed_out) atomicExch(timed_out, 1 + i);
  __syncwarp();
  __threadfence_system();
}

int push_flags(uint64_t* const* d_flags, int n, uint64_t value, cudaStream_t stream) {
  FlagPtrs f;
  memset(&f, 0, sizeof(f));
  int any = 0;
  for (int i = 0; i < n; ++i) {
    f.p[i] = (unsigned long long*)d_flags[i];
    any |= d_flags[i] != nullptr;
  }
  if (!any) return DLB_OK;
  signal_flags_byval_kernel<<<1, 32, 0, stream>>>(f, n, (unsigned long long)value);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}

}  // namespace

// used by dlb_ftle_stencil_gather (stencil.cu): the release that follows the fused kernels
int dlb_push_flags_after(uint64_t* const* d_flags, int n, uint64_t value, cudaStream_t stream) {
  return push_flags(d_flags, n, value, stream);
}

extern "C" int dlb_device_count(int* n) {
  if (!n) {
    dlb_set_error("dlb_device_count: null pointer");
    return DLB_ERR_ARG;
  }
  DLB_CUDA_CHECK(cudaGetDeviceCount(n));
  return DLB_OK;
}

extern "C" int dlb_peer_enable(int device, int peer) {
  if (device == peer) return DLB_OK;
  int can = 0, cur = 0;
  DLB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device, peer));
  if (!can) {
    dlb_set_error("device %d cannot address device %d's memory (no NVLink / PCIe peer path)",
                  device, peer);
    return DLB_ERR_CUDA;
  }
  DLB_CUDA_CHECK(cudaGetDevice(&cur));
  DLB_CUDA_CHECK(cudaSetDevice(device));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    e = cudaSuccess;
  }
  cudaSetDevice(cur);
  DLB_CUDA_CHECK(e);
  return DLB_OK;
}

extern "C" int dlb_peer_alloc(size_t bytes, void** d_ptr) {
  if (!d_ptr) {
    dlb_set_error("dlb_peer_alloc: null pointer");
    return DLB_ERR_ARG;
  }
  // plain cudaMalloc: exportable with cudaIpcGetMemHandle (pool / VMM allocations are not)
  DLB_CUDA_CHECK(cudaMalloc(d_ptr, bytes ? bytes : 1));
  DLB_CUDA_CHECK(cudaMemset(*d_ptr, 0, bytes ? bytes : 1));
  return DLB_OK;
}

extern "C" int dlb_peer_free(void* d_ptr) {
  if (d_ptr) DLB_CUDA_CHECK(cudaFree(d_ptr));
  return DLB_OK;
}

extern "C" int dlb_ipc_export(const void* d_ptr, unsigned char handle[DLB_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == DLB_IPC_HANDLE_BYTES, "IPC handle size");
  if (!d_ptr || !handle) {
    dlb_set_error("dlb_ipc_export: null pointer");
    return DLB_ERR_ARG;
  }
  cudaIpcMemHandle_t h;
  DLB_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
  memcpy(handle, &h, sizeof(h));
  return DLB_OK;
}

extern "C" int dlb_ipc_open(const unsigned char handle[DLB_IPC_HANDLE_BYTES], void** d_ptr) {
  if (!handle || !d_ptr) {
    dlb_set_error("dlb_ipc_open: null pointer");
    return DLB_ERR_ARG;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  DLB_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DLB_OK;
}

extern "C" int dlb_ipc_close(void* d_ptr) {
  if (d_ptr) DLB_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
  return DLB_OK;
}

extern "C" int dlb_peer_copy(void* d_dst, const void* d_src, size_t bytes, void* stream_) {
  if (bytes == 0) return DLB_OK;
  if (!d_dst || !d_src) {
    dlb_set_error("dlb_peer_copy: null pointer");
    return DLB_ERR_ARG;
  }
  DLB_CUDA_CHECK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream_));
  return DLB_OK;
}

extern "C" int dlb_gather_rows(const void* d_src, size_t bytes, void* const* d_dst, int n_dst,
                               uint64_t* const* d_flags, uint64_t flag_value, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_dst < 0 || n_dst > DLB_MAX_PEERS || (n_dst && (!d_dst || !d_flags)) || (bytes && !d_src)) {
    dlb_set_error("dlb_gather_rows: bad argument");
    return DLB_ERR_ARG;
  }
  if (bytes)
    for (int i = 0; i < n_dst; ++i)
      if (d_dst[i] && d_dst[i] != d_src)
        DLB_CUDA_CHECK(cudaMemcpyAsync(d_dst[i], d_src, bytes, cudaMemcpyDefault, stream));
  return push_flags(d_flags, n_dst, flag_value, stream);
}

extern "C" int dlb_halo_exchange(const void* d_local, size_t row_bytes, int64_t first_own,
                                 int64_t last_own, void* d_up_slot, void* d_down_slot,
                                 uint64_t* d_up_flag, uint64_t* d_down_flag, uint64_t flag_value,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!d_local || row_bytes == 0 || first_own < 0 || last_own < first_own) {
    dlb_set_error("dlb_halo_exchange: bad argument");
    return DLB_ERR_ARG;
  }
  const char* base = (const char*)d_local;
  if (d_up_slot)
    DLB_CUDA_CHECK(cudaMemcpyAsync(d_up_slot, base + (size_t)first_own * row_bytes, row_bytes,
                                   cudaMemcpyDefault, stream));
  if (d_down_slot)
    DLB_CUDA_CHECK(cudaMemcpyAsync(d_down_slot, base + (size_t)last_own * row_bytes, row_bytes,
                                   cudaMemcpyDefault, stream));
  uint64_t* flags[2] = {d_up_slot ? d_up_flag : nullptr, d_down_slot ? d_down_flag : nullptr};
  return push_flags(flags, 2, flag_value, stream);
}

extern "C" int dlb_wait_flags(const uint64_t* d_flags, int n, uint64_t min_value,
                              double timeout_s, int* d_timed_out, void* stream_) {
  if (n == 0) return DLB_OK;
  if (!d_flags || n < 0 || n > 32 || !(timeout_s > 0)) {
    dlb_set_error("dlb_wait_flags: bad argument");
    return DLB_ERR_ARG;
  }
  wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(
      (const unsigned long long*)d_flags, n, (unsigned long long)min_value,
      (unsigned long long)(timeout_s * 1e9), d_timed_out);
  DLB_CUDA_CHECK(cudaGetLastError());
  return DLB_OK;
}
