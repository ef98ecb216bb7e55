// common.cuh : host-side plumbing shared by the C-ABI translation units - error capture,
// the library stream, launch accounting and a small RAII device buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

namespace dla {

struct Runtime {
  int device = -1;
  bool ready = false;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of the next batch's spectra while the current batch computes
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  double last_kernel_ms = 0.0;
  long long launches = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
};

Runtime& runtime();
void set_error(const std::string& msg);
int fail(const std::string& msg);

#define DLA_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t err__ = (call);                                                                \
    if (err__ != cudaSuccess) {                                                                \
      char buf__[512];                                                                         \
      snprintf(buf__, sizeof(buf__), "CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(err__), \
               __FILE__, __LINE__, #call);                                                     \
      return ::dla::fail(buf__);                                                               \
    }                                                                                          \
  } while (0)

#define DLA_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::dla::fail(msg); \
  } while (0)

#define DLA_CHECK_READY() \
  do {                    \
    int rc__ = ::dla::ensure_ready(); \
    if (rc__) return rc__; \
  } while (0)

int ensure_ready();

// count a kernel launch and surface launch errors
#define DLA_LAUNCHED()                 \
  do {                                 \
    ::dla::runtime().launches += 1;    \
    DLA_CUDA(cudaGetLastError());      \
  } while (0)

// Free list of large device allocations.  The class API creates one dla_spectrum per set_data call (three per
// spectrum: null, subDLA, DLA model) and each grows a profile cache, product buffers and a Gram basis of tens to
// hundreds of MB; cudaMalloc + cudaFree of those (cudaFree synchronises the device) dominated the per-spectrum
// latency (59 ms per spectrum; 15.6 ms with the large buffers pooled).  Buffers of at least POOL_MIN_BYTES go back to this list when their owner dies and are handed to the next
// request they fit (first fit, at most 2x oversize); the list is bounded and flushed when dla_init changes device.
struct BufferPool {
  static constexpr size_t POOL_MIN_BYTES = (size_t)4 << 10;   // a set_data call makes ~20 allocations of 5 KB - 1 MB as well
  static constexpr size_t POOL_MAX_ENTRIES = 512;
  static constexpr size_t POOL_MAX_ENTRY_BYTES = (size_t)1 << 30;   // catalogue-sized buffers are simply freed
  static constexpr size_t POOL_MAX_TOTAL_BYTES = (size_t)4 << 30;
  struct Entry { void* p; size_t bytes; };
  std::vector<Entry> free_list;
  size_t total_bytes = 0;
  void* take(size_t bytes, size_t* got) {
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].bytes >= bytes && free_list[i].bytes <= 2 * bytes) {
        void* p = free_list[i].p;
        *got = free_list[i].bytes;
        total_bytes -= free_list[i].bytes;
        free_list.erase(free_list.begin() + i);
        return p;
      }
    return nullptr;
  }
  bool give(void* p, size_t bytes) {
    if (bytes < POOL_MIN_BYTES || bytes > POOL_MAX_ENTRY_BYTES || free_list.size() >= POOL_MAX_ENTRIES ||
        total_bytes + bytes > POOL_MAX_TOTAL_BYTES)
      return false;
    free_list.push_back({p, bytes});
    total_bytes += bytes;
    return true;
  }
  void flush() {
    for (Entry& e : free_list) cudaFree(e.p);
    free_list.clear();
    total_bytes = 0;
  }
};
BufferPool& buffer_pool();

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && !buffer_pool().give(p, n * sizeof(T))) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    if (count * sizeof(T) >= BufferPool::POOL_MIN_BYTES) {
      size_t got = 0;
      if (void* q = buffer_pool().take(count * sizeof(T), &got)) {
        p = static_cast<T*>(q);
        n = got / sizeof(T);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
  cudaError_t upload(const T* host, size_t count, cudaStream_t s) {
    return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  cudaError_t download(T* host, size_t count, cudaStream_t s) const {
    return cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s);
  }
};

// page-locked host staging buffer (descriptor uploads and result read-backs that must not block the host)
template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { release(); }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t ensure(size_t count) {
    if (count <= n) return cudaSuccess;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocHost(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
};

// bracket the kernels of one API call with events on the library stream
struct KernelTimer {
  bool active = false;
  cudaError_t begin();
  cudaError_t end();  // synchronises the stream and stores runtime().last_kernel_ms
};

}  // namespace dla
