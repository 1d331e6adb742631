// Field-buffer pool behind gx_malloc / gx_free.
//
// The reference copies fields freely (fields::Field copy constructor, multilevelfield.hpp:40-60, applyFilterInWindow
// field.hpp:752-772, addFieldFromDifferentGridWithFilter :819-832): one covector metric sweep of a three-level zoom makes and
// drops ~20 field-sized temporaries. With cudaMalloc / cudaFree every drop is a device-wide synchronisation plus an unmap and
// every copy a fresh mapping: config 3 spent most of its 5.6 s there (profiles/r1s17_zoom256.json, 280 us per launch).
// Freed blocks up to GX_POOL_MAX_BLOCK bytes are kept (at most GX_POOL_BYTES in total) and handed back by exact size.
//
// Ordering: a block is freed with work possibly still in flight. The free records an event on the legacy default stream
// (which also orders after every blocking stream); the block is only handed out again once that event has completed,
// otherwise gx_malloc takes another block or maps a new one -- never blocks the host. Callers that free a buffer used on a
// NON-blocking stream synchronise that stream first (slab.py does; so did the implicit synchronisation of cudaFree).
// Blocks stay plain cudaMalloc allocations: the CUDA IPC entry points keep working on them.
#include "gx_common.cuh"
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>

namespace gx {
namespace {
struct FreeBlock { void *ptr; cudaEvent_t ev; int device; };
struct Pool {
  std::mutex mu;
  std::unordered_map<void *, size_t> live;      // poolable blocks handed out: ptr -> bytes
  std::multimap<size_t, FreeBlock> idle;        // by exact size
  size_t idle_bytes = 0;
  size_t cap = 0, max_block = 0;
  bool configured = false;
  void configure() {
    if (configured) return;
    configured = true;
    const char *on = getenv("GX_POOL");
    const bool enabled = !(on && atoi(on) == 0);
    const char *c = getenv("GX_POOL_BYTES"), *m = getenv("GX_POOL_MAX_BLOCK");
    cap = enabled ? (c ? strtoull(c, nullptr, 10) : (size_t) 8 << 30) : 0;
    max_block = enabled ? (m ? strtoull(m, nullptr, 10) : (size_t) 1536 << 20) : 0;
  }
};
Pool &pool() { static Pool p; return p; }

void drop_idle_locked(Pool &p) {
  for (auto &kv : p.idle) {
    cudaEventDestroy(kv.second.ev);
    cudaFree(kv.second.ptr);
  }
  p.idle.clear();
  p.idle_bytes = 0;
}
}  // namespace

int pool_malloc(void **ptr, size_t bytes) {
  Pool &p = pool();
  std::lock_guard<std::mutex> lock(p.mu);
  p.configure();
  const bool poolable = bytes > 0 && bytes <= p.max_block;
  int dev = 0;
  if (poolable) {
    GX_CUDA_CHECK(cudaGetDevice(&dev));
    auto range = p.idle.equal_range(bytes);
    for (auto it = range.first; it != range.second; ++it) {
      if (it->second.device != dev) continue;
      if (cudaEventQuery(it->second.ev) != cudaSuccess) continue;  // its last user is still running: leave it
      *ptr = it->second.ptr;
      cudaEventDestroy(it->second.ev);
      p.idle.erase(it);
      p.idle_bytes -= bytes;
      p.live[*ptr] = bytes;
      return 0;
    }
  }
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaErrorMemoryAllocation && !p.idle.empty()) {  // give the pooled blocks back and try again
    cudaGetLastError();
    cudaDeviceSynchronize();
    drop_idle_locked(p);
    e = cudaMalloc(ptr, bytes);
  }
  GX_CUDA_CHECK(e);
  if (poolable) p.live[*ptr] = bytes;
  return 0;
}

int pool_free(void *ptr) {
  if (!ptr) return 0;
  Pool &p = pool();
  std::lock_guard<std::mutex> lock(p.mu);
  auto it = p.live.find(ptr);
  if (it == p.live.end()) {
    GX_CUDA_CHECK(cudaFree(ptr));
    return 0;
  }
  const size_t bytes = it->second;
  p.live.erase(it);
  if (p.idle_bytes + bytes > p.cap) {
    GX_CUDA_CHECK(cudaFree(ptr));
    return 0;
  }
  FreeBlock b{ptr, nullptr, 0};
  GX_CUDA_CHECK(cudaGetDevice(&b.device));
  GX_CUDA_CHECK(cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming));
  GX_CUDA_CHECK(cudaEventRecord(b.ev, nullptr));
  p.idle.emplace(bytes, b);
  p.idle_bytes += bytes;
  return 0;
}

void pool_release() {
  Pool &p = pool();
  std::lock_guard<std::mutex> lock(p.mu);
  if (p.idle.empty()) return;
  cudaDeviceSynchronize();
  drop_idle_locked(p);
}

size_t pool_idle_bytes() {
  Pool &p = pool();
  std::lock_guard<std::mutex> lock(p.mu);
  return p.idle_bytes;
}
}  // namespace gx
