// mock_cuda_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see mock_cuda_runtime.h).
#include "mock_cuda_runtime.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <thread>

namespace {
using Clock = std::chrono::steady_clock;

struct Alloc { size_t bytes; cudaMemoryType type; int device; };
struct Globals {
  std::mutex mu;
  std::map<uintptr_t, Alloc> allocs;                 // base address -> allocation
  std::map<std::string, int> fail_nth;               // api name -> calls left until the injected failure (1 = the next call)
  std::deque<MockStream*> streams;
  int devices = 2;
  std::atomic<int> jitter_us{0}, expect_zero_free{0};
  std::atomic<int> kernel_delay_us[16];             // extra time of every kernel task on a device (a slow GPU / a slow link)
  std::atomic<long> violations{0}, memcpy_bytes{0}, kernel_tasks{0};
};
Globals& G() { static Globals* g = [] { Globals* x = new Globals; for (auto& d : x->kernel_delay_us) d.store(0); return x; }(); return *g; }     // never destroyed: stream threads outlive static destructors
thread_local int tl_device = 0;
thread_local cudaError_t tl_last = cudaSuccess;

cudaError_t ret(cudaError_t e) { if (e != cudaSuccess) tl_last = e; return e; }
bool inject(const char* api) {
  std::lock_guard<std::mutex> l(G().mu);
  auto it = G().fail_nth.find(api);
  if (it == G().fail_nth.end() || it->second <= 0) return false;
  if (--it->second == 0) { G().fail_nth.erase(it); return true; }
  return false;
}
// true if [p, p+bytes) lies inside one registered allocation or touches none
bool range_ok(const void* p, size_t bytes) {
  if (!bytes) return true;
  std::lock_guard<std::mutex> l(G().mu);
  const uintptr_t a = (uintptr_t)p;
  auto it = G().allocs.upper_bound(a);
  if (it != G().allocs.begin()) {
    auto b = std::prev(it);
    if (a < b->first + b->second.bytes) return a + bytes <= b->first + b->second.bytes;     // starts inside: must end inside
  }
  return it == G().allocs.end() || a + bytes <= it->first;                                   // starts outside: must not run into one
}
}  // namespace

struct MockStream {
  std::mutex mu; std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool busy = false, stop = false;
  std::thread th;
  std::minstd_rand rng{12345};
  int device = tl_device;
  MockStream() { th = std::thread([this] { run(); }); }
  void run() {
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return stop || !q.empty(); });
        if (q.empty()) return;
        fn = std::move(q.front()); q.pop_front(); busy = true;
      }
      const int j = G().jitter_us.load();
      if (j > 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % (unsigned)j));
      fn();
      { std::lock_guard<std::mutex> l(mu); busy = false; }
      cv.notify_all();
    }
  }
  void push(std::function<void()> fn) { { std::lock_guard<std::mutex> l(mu); q.push_back(std::move(fn)); } cv.notify_all(); }
  void sync() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return q.empty() && !busy; }); }
};
struct MockEvent {
  std::mutex mu; std::condition_variable cv;
  long recorded = 0, completed = 0;
  Clock::time_point t;
};

void mock_stream_enqueue(cudaStream_t s, std::function<void()> fn) {
  G().kernel_tasks++;
  const int us = G().kernel_delay_us[s->device & 15].load();
  if (us > 0) s->push([fn, us] { std::this_thread::sleep_for(std::chrono::microseconds(us)); fn(); }); else s->push(std::move(fn));
}

const char* cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (mock)";
    case cudaErrorMemoryAllocation: return "out of memory (mock)";
    case cudaErrorNoDevice: return "no CUDA-capable device is detected (mock)";
    default: return "injected failure (mock)";
  }
}
cudaError_t cudaGetLastError() { cudaError_t e = tl_last; tl_last = cudaSuccess; return e; }
cudaError_t cudaGetDeviceCount(int* n) {
  int d; { std::lock_guard<std::mutex> l(G().mu); d = G().devices; }
  *n = d;
  return ret(d > 0 ? cudaSuccess : cudaErrorNoDevice);
}
cudaError_t cudaSetDevice(int dev) {
  int d; { std::lock_guard<std::mutex> l(G().mu); d = G().devices; }
  if (dev < 0 || dev >= d) return ret(cudaErrorInvalidValue);
  tl_device = dev;
  return cudaSuccess;
}
static void sync_all() {
  std::deque<MockStream*> ss;
  { std::lock_guard<std::mutex> l(G().mu); ss = G().streams; }
  for (MockStream* s : ss) s->sync();
}
cudaError_t cudaDeviceSynchronize() { sync_all(); return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  if (inject("cudaStreamCreateWithFlags")) return ret(cudaErrorUnknown);
  *s = new MockStream;
  std::lock_guard<std::mutex> l(G().mu); G().streams.push_back(*s);
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  s->sync();
  { std::lock_guard<std::mutex> l(G().mu); for (auto it = G().streams.begin(); it != G().streams.end(); ++it) if (*it == s) { G().streams.erase(it); break; } }
  { std::lock_guard<std::mutex> l(s->mu); s->stop = true; }
  s->cv.notify_all();
  s->th.join();
  delete s;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { if (inject("cudaStreamSynchronize")) return ret(cudaErrorUnknown); s->sync(); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { if (inject("cudaEventCreate")) return ret(cudaErrorUnknown); *e = new MockEvent; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  if (inject("cudaEventRecord")) return ret(cudaErrorUnknown);
  long seq; { std::lock_guard<std::mutex> l(e->mu); seq = ++e->recorded; }
  s->push([e, seq] { { std::lock_guard<std::mutex> l(e->mu); e->t = Clock::now(); e->completed = seq; } e->cv.notify_all(); });
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  if (inject("cudaEventSynchronize")) return ret(cudaErrorUnknown);
  std::unique_lock<std::mutex> l(e->mu);
  const long want = e->recorded;
  e->cv.wait(l, [&] { return e->completed >= want; });
  return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  Clock::time_point ta, tb;
  { std::lock_guard<std::mutex> l(a->mu); if (a->completed < a->recorded || !a->recorded) return ret(cudaErrorInvalidValue); ta = a->t; }
  { std::lock_guard<std::mutex> l(b->mu); if (b->completed < b->recorded || !b->recorded) return ret(cudaErrorInvalidValue); tb = b->t; }
  *ms = std::chrono::duration<float, std::milli>(tb - ta).count();
  return cudaSuccess;
}
static cudaError_t alloc(void** p, size_t bytes, cudaMemoryType type, const char* api) {
  if (inject(api)) return ret(cudaErrorMemoryAllocation);
  void* q = malloc(bytes ? bytes : 1);
  if (!q) return ret(cudaErrorMemoryAllocation);
  memset(q, 0xA5, bytes);                                    // never hand out zeroed memory: stale reads must show
  { std::lock_guard<std::mutex> l(G().mu); G().allocs[(uintptr_t)q] = {bytes, type, tl_device}; }
  *p = q;
  return cudaSuccess;
}
static cudaError_t release(void* p, cudaMemoryType type) {
  if (!p) return cudaSuccess;
  sync_all();                                                // like the real cudaFree: waits for the device
  {
    std::lock_guard<std::mutex> l(G().mu);
    auto it = G().allocs.find((uintptr_t)p);
    if (it == G().allocs.end() || it->second.type != type) { G().violations++; return ret(cudaErrorInvalidValue); }
    if (G().expect_zero_free.load()) {                        // the engine promises to wipe secret-bearing buffers before freeing them
      const unsigned char* q = (const unsigned char*)p;
      for (size_t i = 0; i < it->second.bytes; i++) if (q[i]) { G().violations++; break; }
    }
    G().allocs.erase(it);
  }
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t bytes) { return alloc(p, bytes, cudaMemoryTypeDevice, "cudaMalloc"); }
cudaError_t cudaFree(void* p) { return release(p, cudaMemoryTypeDevice); }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return alloc(p, bytes, cudaMemoryTypeHost, "cudaHostAlloc"); }
cudaError_t cudaFreeHost(void* p) { return release(p, cudaMemoryTypeHost); }
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned) {
  if (inject("cudaHostRegister")) return ret(cudaErrorUnknown);
  std::lock_guard<std::mutex> l(G().mu); G().allocs[(uintptr_t)p] = {bytes, cudaMemoryTypeHost, tl_device};
  return cudaSuccess;
}
cudaError_t cudaHostUnregister(void* p) {
  sync_all();
  std::lock_guard<std::mutex> l(G().mu);
  auto it = G().allocs.find((uintptr_t)p);
  if (it == G().allocs.end()) { G().violations++; return ret(cudaErrorInvalidValue); }
  G().allocs.erase(it);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t s) {
  if (inject("cudaMemcpyAsync")) return ret(cudaErrorUnknown);
  if (!range_ok(dst, bytes) || !range_ok(src, bytes)) { G().violations++; return ret(cudaErrorInvalidValue); }
  G().memcpy_bytes += (long)bytes;
  s->push([=] { memcpy(dst, src, bytes); });
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
  if (!range_ok(dst, bytes) || !range_ok(src, bytes)) { G().violations++; return ret(cudaErrorInvalidValue); }
  sync_all();
  memcpy(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s) {
  if (!range_ok(p, bytes)) { G().violations++; return ret(cudaErrorInvalidValue); }
  s->push([=] { memset(p, v, bytes); });
  return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void* p) {
  std::lock_guard<std::mutex> l(G().mu);
  const uintptr_t a = (uintptr_t)p;
  auto it = G().allocs.upper_bound(a);
  at->type = cudaMemoryTypeUnregistered; at->device = 0;
  if (it != G().allocs.begin()) {
    auto b = std::prev(it);
    if (a < b->first + b->second.bytes) { at->type = b->second.type; at->device = b->second.device; }
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------- controls for the tests
extern "C" {
__attribute__((visibility("default"))) void mock_set_device_count(int n) { std::lock_guard<std::mutex> l(G().mu); G().devices = n; }
// the nth call from now of `api` fails (1 = the next one; 0 clears)
__attribute__((visibility("default"))) void mock_fail_nth(const char* api, int nth) { std::lock_guard<std::mutex> l(G().mu); if (nth > 0) G().fail_nth[api] = nth; else G().fail_nth.erase(api); }
__attribute__((visibility("default"))) void mock_set_jitter_us(int us) { G().jitter_us.store(us); }
__attribute__((visibility("default"))) void mock_set_kernel_delay_us(int dev, int us) { G().kernel_delay_us[dev & 15].store(us); }
__attribute__((visibility("default"))) long mock_violations() { return G().violations.load(); }
__attribute__((visibility("default"))) long mock_memcpy_bytes() { return G().memcpy_bytes.load(); }
// live allocations of a type (2 = device, 1 = pinned host) and their total size
__attribute__((visibility("default"))) long mock_live_allocs(int type, long* bytes) {
  std::lock_guard<std::mutex> l(G().mu);
  long n = 0, b = 0;
  for (auto& kv : G().allocs) if ((int)kv.second.type == type) { n++; b += (long)kv.second.bytes; }
  if (bytes) *bytes = b;
  return n;
}
// number of live allocations of `type` (2 = device, 1 = pinned host) of at least `min_bytes` bytes that hold a non-zero byte
__attribute__((visibility("default"))) long mock_nonzero_allocs(int type, long min_bytes) {
  sync_all();
  std::lock_guard<std::mutex> l(G().mu);
  long n = 0;
  for (auto& kv : G().allocs) {
    if ((int)kv.second.type != type || (long)kv.second.bytes < min_bytes) continue;
    const unsigned char* q = (const unsigned char*)kv.first;
    for (size_t i = 0; i < kv.second.bytes; i++) if (q[i]) { n++; break; }
  }
  return n;
}
// while on, freeing a device or pinned buffer that is not all zeros counts as a violation (fq_trim's wipe)
__attribute__((visibility("default"))) void mock_expect_zero_on_free(int on) { G().expect_zero_free.store(on); }
__attribute__((visibility("default"))) int mock_is_pinned(const void* p) {
  cudaPointerAttributes at; cudaPointerGetAttributes(&at, p); return at.type == cudaMemoryTypeHost;
}
}
