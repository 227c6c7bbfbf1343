// common.h -- shared host-side helpers for the kkx CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <stdexcept>

namespace kkx {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
struct ArgError : std::runtime_error {
  explicit ArgError(const std::string& s) : std::runtime_error(s) {}
};
struct IoError : std::runtime_error {
  explicit IoError(const std::string& s) : std::runtime_error(s) {}
};

#define KKX_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
               cudaGetErrorString(_e));                                                     \
      throw ::kkx::CudaError(_b);                                                           \
    }                                                                                       \
  } while (0)

// Launch-count bookkeeping ("gpu_launches" in bench.py) + optional per-launch error checks.
struct LaunchStats {
  int64_t launches = 0;
  bool check_each = false;  // KKX_DEBUG_SYNC=1: synchronize + check after every launch
  // per-kernel timing (kkx_profile_enable): one event after every launch on the in-order stream;
  // the gap between consecutive events is attributed to the launch in between
  bool profile = false;
  bool detail = false;                   // per-shape kernel names (KKX_PROFILE_DETAIL=1)
  std::vector<cudaEvent_t> events;       // pool, events[0] = start of run
  std::vector<std::string> names;        // names[i] = kernel launched before events[i+1]
  size_t n_events = 0;
  double conv_flops = 0;                 // algorithmic FLOPs issued through the shifted-GEMM kernels (all of them)
  double arb_flops = 0, arb_bytes = 0;   // share of the fused res-block conv (kernels_arb.cu) + its algorithmic HBM bytes
  cudaEvent_t next_event() {
    if (n_events == events.size()) {
      cudaEvent_t e; cudaEventCreate(&e); events.push_back(e);
    }
    return events[n_events++];
  }
};
extern thread_local LaunchStats* g_launch_stats;
extern thread_local bool g_dry_run;  // true while sizing arenas: launchers return immediately

inline void post_launch(const char* name, cudaStream_t st) {
  if (g_launch_stats) {
    g_launch_stats->launches++;
    if (g_launch_stats->profile) {
      cudaEventRecord(g_launch_stats->next_event(), st);
      g_launch_stats->names.push_back(name);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && g_launch_stats && g_launch_stats->check_each) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    char b[512];
    snprintf(b, sizeof b, "kernel %s failed: %s", name, cudaGetErrorString(e));
    throw CudaError(b);
  }
}

// Bump allocator over one device allocation; reset per phase.  All returned pointers are
// 256-byte aligned.  Grow-only: a too-small arena is re-allocated (caller must have synced).
class Arena {
 public:
  ~Arena() { release(); }
  void release() {
    if (base_) cudaFree(base_);
    base_ = nullptr; cap_ = 0; used_ = 0;
  }
  void reserve(size_t bytes) {
    if (bytes <= cap_) return;
    release();
    KKX_CUDA(cudaMalloc(&base_, bytes));
    cap_ = bytes;
  }
  void reset() { used_ = 0; }
  // virtual mode: allocations only advance the counter (used to size the arena by a dry run)
  void set_virtual(bool v) { virtual_ = v; }
  size_t used() const { return used_; }
  size_t capacity() const { return cap_; }
  void* alloc_bytes(size_t bytes) {
    size_t a = (used_ + 255) & ~size_t(255);
    if (!virtual_ && a + bytes > cap_) {
      char b[256];
      snprintf(b, sizeof b, "arena overflow: need %zu + %zu > %zu", a, bytes, cap_);
      throw CudaError(b);
    }
    used_ = a + bytes;
    return static_cast<char*>(base_) + a;
  }
  template <class T> T* alloc(size_t n) { return static_cast<T*>(alloc_bytes(n * sizeof(T))); }

 private:
  void* base_ = nullptr;
  size_t cap_ = 0, used_ = 0;
  bool virtual_ = false;
};

// A ragged "level": B items packed along the row axis with zero gaps between items.
// Item b owns rows [off[b], off[b] + len[b]).
struct Level {
  int B = 0;
  std::vector<int> off, len;
  int rows = 0;      // total rows incl. gaps
  int max_len = 0;
  long long sum_len = 0;
  int* d_off = nullptr;  // device copies
  int* d_len = nullptr;
  // prefix sums (B+1) of the per-item tile counts for 128- and 256-row tiles (persistent kernels)
  int* d_tiles128 = nullptr; int* d_tiles256 = nullptr; int* d_tiles512 = nullptr;
  int ntiles128 = 0, ntiles256 = 0, ntiles512 = 0;
};

constexpr int kGapRows = 32;  // >= largest conv halo (k=11, dil=5 -> 25)

}  // namespace kkx
