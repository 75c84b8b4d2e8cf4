// stage.cu — pinned, chunked, double-buffered host<->device staging for the host-pointer entry points.
//
// The reference's call surface hands over std::vector storage: pageable memory.  cudaMemcpyAsync
// from pageable memory is staged by the runtime through one internal buffer, one thread, ~8 GB/s,
// which made the host-pointer ABI 60x (PICP) to 1000x (triangulation) slower than the kernels it
// feeds.  Here the library owns a ring of pinned chunks per device; a small
// pool of worker threads copies the caller's memory into / out of the ring in parallel while the
// DMA engine moves the previous chunk, so a transfer runs at min(host memcpy bandwidth of the pool,
// PCIe) and the caller's buffer is free again when the call returns.
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vo {

namespace {

// bytes per pinned chunk (VO_STAGE_CHUNK_MB overrides, 1..64; read once)
static const size_t STAGE_CHUNK = []() {
  size_t mb = 16;  // measured (tools/stage_probe.py, 409 MB through vo_triangulate): 2 MB 28.8 ms, 4 MB 18.2, 8 MB 16.3, 16 MB 14.8
  if (const char* e = getenv("VO_STAGE_CHUNK_MB")) {
    const long v = atol(e);
    if (v >= 1 && v <= 64) mb = (size_t)v;
  }
  return mb << 20;
}();
constexpr int STAGE_SLOTS = 4;                   // chunks in the ring
constexpr int STAGE_MAX_WORKERS = 7;             // + the calling thread
constexpr size_t STAGE_MIN = (size_t)1 << 20;    // below: one plain cudaMemcpyAsync

// parallel memcpy: the caller and `n` persistent workers each copy one slice
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) {
      std::thread t([this, i]() { loop(i); });
      t.detach();  // parked on the condition variable for the life of the process
      ++n_;
    }
  }
  void copy(void* dst, const void* src, size_t bytes) {
    if (n_ == 0 || bytes < ((size_t)256 << 10)) {
      memcpy(dst, src, bytes);
      return;
    }
    const int parts = n_ + 1;
    // 4 KB-aligned slices so that no two threads share a page
    const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      slice_ = slice;
      pending_ = n_;
      ++gen_;
    }
    cv_.notify_all();
    if (slice < bytes || parts == 1) memcpy(dst, src, slice < bytes ? slice : bytes);
    else memcpy(dst, src, bytes);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this]() { return pending_ == 0; });
  }

 private:
  void loop(int id) {
    unsigned long long seen = 0;
    for (;;) {
      char* d;
      const char* s;
      size_t off, len;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
        off = (size_t)(id + 1) * slice_;
        len = off < bytes_ ? (bytes_ - off < slice_ ? bytes_ - off : slice_) : 0;
        d = dst_ + off;
        s = src_ + off;
      }
      if (len) memcpy(d, s, len);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  int n_ = 0;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
};

CopyPool* copy_pool() {
  static CopyPool* pool = []() {
    int n = 0;
    if (const char* e = getenv("VO_STAGE_THREADS")) n = atoi(e) - 1;
    else {
      const unsigned hw = std::thread::hardware_concurrency();
      n = (int)(hw / 2);
      if (n > STAGE_MAX_WORKERS) n = STAGE_MAX_WORKERS;
    }
    if (n < 0) n = 0;
    return new CopyPool(n);  // never destroyed: no static-destruction order to get wrong
  }();
  return pool;
}

struct Ring {
  std::mutex mu;  // one transfer at a time per device
  bool ready = false;
  void* pin[STAGE_SLOTS] = {};
  cudaEvent_t ev[STAGE_SLOTS] = {};
  bool used[STAGE_SLOTS] = {};
  int next = 0;
  int init() {
    if (ready) return VO_OK;
    for (int i = 0; i < STAGE_SLOTS; ++i) {
      VO_CUDA(cudaHostAlloc(&pin[i], STAGE_CHUNK, cudaHostAllocDefault));
      VO_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    ready = true;
    return VO_OK;
  }
  // the slot's previous DMA (either direction) has finished
  int acquire(int* slot) {
    const int s = next;
    next = (next + 1) % STAGE_SLOTS;
    if (used[s]) VO_CUDA(cudaEventSynchronize(ev[s]));
    used[s] = true;
    *slot = s;
    return VO_OK;
  }
};

Ring* ring_for(int device) {
  static Ring rings[64];
  return (device >= 0 && device < 64) ? &rings[device] : nullptr;
}

}  // namespace

// host -> device on stream `s` of the current device.  When the call returns the caller's buffer
// is no longer referenced (the last chunks may still be in flight out of the pinned ring).
int stage_h2d(int device, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return VO_OK;
  Ring* r = ring_for(device);
  if (bytes < STAGE_MIN || r == nullptr || !host_source_is_pageable(src_host)) {
    VO_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, s));
    // a pinned / managed source is read asynchronously: it must outlive the copy
    if (host_source_still_in_use(src_host)) VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
  }
  std::lock_guard<std::mutex> lk(r->mu);
  int rc = r->init();
  if (rc) return rc;
  CopyPool* pool = copy_pool();
  const char* src = static_cast<const char*>(src_host);
  char* dst = static_cast<char*>(dst_dev);
  for (size_t off = 0; off < bytes; off += STAGE_CHUNK) {
    const size_t len = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
    int slot;
    if ((rc = r->acquire(&slot))) return rc;
    pool->copy(r->pin[slot], src + off, len);  // overlaps the DMA of the previous chunk
    VO_CUDA(cudaMemcpyAsync(dst + off, r->pin[slot], len, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaEventRecord(r->ev[slot], s));
  }
  return VO_OK;
}

// device -> host on stream `s`; the data is in dst_host when the call returns.
int stage_d2h(int device, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return VO_OK;
  Ring* r = ring_for(device);
  if (bytes < STAGE_MIN || r == nullptr || !host_source_is_pageable(dst_host)) {
    VO_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
  }
  std::lock_guard<std::mutex> lk(r->mu);
  int rc = r->init();
  if (rc) return rc;
  CopyPool* pool = copy_pool();
  char* dst = static_cast<char*>(dst_host);
  const char* src = static_cast<const char*>(src_dev);
  const size_t n_chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
  int slots[STAGE_SLOTS];
  auto issue = [&](size_t c) -> int {
    const size_t off = c * STAGE_CHUNK;
    const size_t len = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
    int slot;
    int e = r->acquire(&slot);
    if (e) return e;
    slots[c % STAGE_SLOTS] = slot;
    VO_CUDA(cudaMemcpyAsync(r->pin[slot], src + off, len, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaEventRecord(r->ev[slot], s));
    return VO_OK;
  };
  size_t issued = 0;
  for (; issued < n_chunks && issued < (size_t)STAGE_SLOTS - 1; ++issued)
    if ((rc = issue(issued))) return rc;
  for (size_t c = 0; c < n_chunks; ++c) {
    if (issued < n_chunks) {  // keep the DMA engine busy while this chunk is copied out
      if ((rc = issue(issued))) return rc;
      ++issued;
    }
    const size_t off = c * STAGE_CHUNK;
    const size_t len = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
    const int slot = slots[c % STAGE_SLOTS];
    VO_CUDA(cudaEventSynchronize(r->ev[slot]));
    pool->copy(dst + off, r->pin[slot], len);
  }
  return VO_OK;
}

}  // namespace vo
