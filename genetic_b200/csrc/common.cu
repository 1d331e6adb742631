#include "gx_common.cuh"
#include <map>
#include <vector>
#include <cstring>

namespace gx {
static thread_local std::string g_last_error;
unsigned long long g_launch_count = 0;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(const std::string &msg) { g_last_error = msg; return 1; }
const char *last_error_cstr() { return g_last_error.c_str(); }

// ---- timing -------------------------------------------------------------------------------------------
struct TimingRec { std::string name; cudaEvent_t e0, e1; };
static bool g_timing = false;
static std::vector<TimingRec> g_recs;

TimedScope::TimedScope(const char *name, cudaStream_t s) : idx(-1), st(s) {
  if (!g_timing) return;
  TimingRec r;
  r.name = name;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
  idx = (int) g_recs.size() - 1;
}
TimedScope::~TimedScope() {
  if (idx >= 0) cudaEventRecord(g_recs[idx].e1, st);
}
}  // namespace gx

extern "C" {
int gx_timing_enable(int on) {
  for (auto &r : gx::g_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  gx::g_recs.clear();
  gx::g_timing = on != 0;
  return 0;
}
// Writes "name count total_ms" lines into buf (NUL terminated); returns number of classes.
int gx_timing_report(char *buf, int buflen) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int, double>> acc;
  for (auto &r : gx::g_recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { acc[r.name].first++; acc[r.name].second += ms; }
  }
  std::string out;
  for (auto &kv : acc) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  if ((int) out.size() + 1 > buflen) out.resize(buflen > 0 ? buflen - 1 : 0);
  if (buflen > 0) { memcpy(buf, out.c_str(), out.size()); buf[out.size()] = 0; }
  return (int) acc.size();
}

// ---- peer memory between the ranks of one node (one process per GPU): CUDA IPC over a cudaMalloc'ed allocation ----------
// handle_out: 64 bytes (cudaIpcMemHandle_t). ptr must be the base of an allocation made by gx_malloc.
int gx_ipc_get_handle(void *ptr, void *handle_out) {
  cudaIpcMemHandle_t h;
  GX_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}
int gx_ipc_open(const void *handle, void **ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  GX_CUDA_CHECK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int gx_ipc_close(void *ptr) {
  GX_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return 0;
}
}
