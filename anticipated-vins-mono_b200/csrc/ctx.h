// ctx.h -- host-side context shared by the C-ABI translation units (ba_api.cu, sel_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/bvio.h"

struct ncclComm;

namespace bvio {

// one device slab + one pinned host slab carved into aligned sub-arrays
struct Slab {
  char* d = nullptr;
  char* h = nullptr;       // pinned mirror of the first h_bytes (inputs + outputs, not scratch)
  size_t d_bytes = 0, h_bytes = 0;
  void release() {
    if (d) cudaFree(d);
    if (h) cudaFreeHost(h);
    d = h = nullptr; d_bytes = h_bytes = 0;
  }
};

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
};

}  // namespace bvio

namespace bvio {
// Persistent host workers for packing / validating / scattering large batches: a batch call used to spawn (and join) a
// fresh set of threads three times per sub-batch, ~0.4 ms each time.  run(n, f) executes f(0) .. f(n-1), f(0) on the
// calling thread, and returns when all are done.  One pool per context; calls on one context are not concurrent.
class HostPool {
 public:
  ~HostPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }
  void run(int n, const std::function<void(int)>& f) {
    if (n <= 1) { if (n == 1) f(0); return; }
    while ((int)th_.size() < n - 1) { const int id = (int)th_.size() + 1; th_.emplace_back([this, id]() { worker(id); }); }
    { std::lock_guard<std::mutex> l(m_); fn_ = &f; ntask_ = n; pending_ = n - 1; gen_++; }
    cv_work_.notify_all();
    f(0);
    std::unique_lock<std::mutex> l(m_);
    cv_done_.wait(l, [this]() { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void worker(int id) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* f = nullptr;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_work_.wait(l, [&]() { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (id < ntask_) f = fn_;
      }
      if (f) {
        (*f)(id);
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) cv_done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  int ntask_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};
}  // namespace bvio

struct bvio_ctx {
  bvio::HostPool pool;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // H2D of the next sub-batch while the previous one computes (pipelined batches)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  int sm_count = 0;
  // grow-only slab caches so that one-shot calls (bvio_optimize / bvio_select at frame rate) do not
  // pay cudaMalloc + cudaMallocHost every call
  bvio::Slab ba_cache, sel_cache;
  bool ba_cache_busy = false, sel_cache_busy = false;
  // sub-batch slabs of the pipelined bvio_optimize_batch (host packing of sub-batch i+1 overlaps the solve of i)
  static constexpr int PIPE = 8;
  bvio::Slab ba_pipe[PIPE];
  bool ba_pipe_busy[PIPE] = {false, false, false, false, false, false, false, false};
  char* marg_scratch = nullptr;   // grow-only device scratch of bvio_marginalize
  size_t marg_bytes = 0;
  char* marg_host = nullptr;      // pinned staging of bvio_marginalize_begin / _end
  size_t marg_host_bytes = 0;
  bool marg_inflight = false;
  // multi-GPU selector
  ncclComm* comm = nullptr;
  int rank = 0, world = 1;
  // fused selector exchange: this rank's mailbox and every rank's mailbox mapped through CUDA IPC
  void* mbox_local = nullptr;
  void* mbox_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool p2p_ready = false;
  unsigned long long sel_epoch = 0;
};

namespace bvio {

// take a slab of at least (d_bytes, h_bytes): from the cache when it is free and large enough
inline cudaError_t slab_acquire(Slab& cache, bool& busy, size_t d_bytes, size_t h_bytes, Slab& out, bool& from_cache) {
  if (!busy) {
    if (cache.d_bytes < d_bytes || cache.h_bytes < h_bytes) {
      size_t nd = d_bytes > cache.d_bytes ? d_bytes + d_bytes / 4 : cache.d_bytes;
      size_t nh = h_bytes > cache.h_bytes ? h_bytes + h_bytes / 4 : cache.h_bytes;
      cache.release();
      cudaError_t e = cudaMalloc((void**)&cache.d, nd);
      if (e != cudaSuccess) return e;
      e = cudaMallocHost((void**)&cache.h, nh);
      if (e != cudaSuccess) { cache.release(); return e; }
      cache.d_bytes = nd; cache.h_bytes = nh;
    }
    out = cache; busy = true; from_cache = true;
    return cudaSuccess;
  }
  from_cache = false;
  out = Slab();
  cudaError_t e = cudaMalloc((void**)&out.d, d_bytes);
  if (e != cudaSuccess) return e;
  e = cudaMallocHost((void**)&out.h, h_bytes);
  if (e != cudaSuccess) { out.release(); return e; }
  out.d_bytes = d_bytes; out.h_bytes = h_bytes;
  return cudaSuccess;
}
inline void slab_release(Slab& s, bool& busy, bool from_cache) {
  if (from_cache) busy = false;
  else s.release();
  s = Slab();
}

inline int fail(bvio_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define BVIO_CUDA_OK(ctx, expr)                                                                     \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return bvio::fail(ctx, BVIO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
  } while (0)

}  // namespace bvio
