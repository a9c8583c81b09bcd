// capi.cu -- the C ABI of include/fourq_b200.h and the host engine behind it.
//
// Host engine.  A call splits its rows into contiguous per-GPU slices (every row is independent, there
// is no exchange step, hence no collective).  Each GPU runs its slice as a software pipeline of chunks over three
// CUDA streams: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c.  Chunks of all devices are
// enqueued round-robin before anything is waited on.  Device staging buffers are kept per (device, stream) and
// grow on demand; nothing else is cached between calls.  There is no CPU implementation of any operation here.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/fourq_b200.h"
#include "kernels.h"

namespace {

constexpr int kStreams = 3;
constexpr int kMaxDev = 16;
constexpr size_t kFlushBytes = 256u << 20;      // > 126 MB L2

thread_local char tl_err[512] = "";
thread_local float tl_kernel_ms = 0.f;
thread_local float tl_phase_ms[3] = {0.f, 0.f, 0.f};
std::mutex g_mu;
int g_dev_base = 0;
int g_strict = -1;            // table selection: 0 = masked loads (default), 1 = strict scan, -1 = not yet read from FQ_STRICT_SELECT

int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(tl_err, sizeof(tl_err), fmt, ap); va_end(ap);
  return code;
}
#define CU(call)                                                                                       \
  do { cudaError_t e_ = (call);                                                                        \
       if (e_ != cudaSuccess) return fail(FQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// device staging buffers of one stream slot (a, b, out, status) and, for pageable result buffers, pinned host staging (2, 3)
struct Slot {
  void* buf[4] = {nullptr, nullptr, nullptr, nullptr}; size_t cap[4] = {0, 0, 0, 0};
  void* hbuf[4] = {nullptr, nullptr, nullptr, nullptr}; size_t hcap[4] = {0, 0, 0, 0};
  // a chunk whose results still sit in hbuf[2] / hbuf[3] and have to be copied to the caller's (pageable) buffers
  bool pending = false; size_t p_r0 = 0, p_rows = 0;
};
struct DevCtx {
  bool ready = false;
  cudaStream_t st[kStreams];
  Slot slot[kStreams];
  void* flush = nullptr;
  void* comb = nullptr;       // per-digit fixed-base tables (kernels_comb.cu)
  void* dh_scratch[kStreams] = {nullptr, nullptr, nullptr};     // table / plan / projective result handed between the DH (and comb) kernels
  size_t dh_scratch_cap[kStreams] = {0, 0, 0};
  int sms = 0;
};
DevCtx g_ctx[kMaxDev];

int device_count() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(FQ_ERR_NO_DEVICE, "no CUDA device available (%s); fourq_b200 has no CPU path", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e)); }
  return n > kMaxDev ? kMaxDev : n;
}

int ctx_init(int dev) {
  DevCtx& c = g_ctx[dev];
  CU(cudaSetDevice(dev));
  if (c.ready) return FQ_OK;
  for (int i = 0; i < kStreams; i++) CU(cudaStreamCreateWithFlags(&c.st[i], cudaStreamNonBlocking));
  CU(fqk_device_init(c.st[0]));
  CU(fqk_comb_init(&c.comb, c.st[0]));
  CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
  c.ready = true;
  return FQ_OK;
}

int slot_reserve(Slot& s, int which, size_t bytes) {
  if (bytes <= s.cap[which]) return FQ_OK;
  if (s.buf[which]) CU(cudaFree(s.buf[which]));
  s.buf[which] = nullptr; s.cap[which] = 0;
  CU(cudaMalloc(&s.buf[which], bytes));
  s.cap[which] = bytes;
  return FQ_OK;
}

bool is_dh_op(int op) { return op == FQ_DEVOP_DH || op == FQ_DEVOP_DH_AFFINE || op == FQ_DEVOP_DH_ENDO || op == FQ_DEVOP_DH_ENDO_AFFINE; }
// fixed-base ops: their kernels hand (X, Y, Z) to k_dh_finish through a small scratch
bool is_comb_op(int op) {
  return op == FQ_DEVOP_DH_BASE_COMB || op == FQ_DEVOP_MUL_BASE_COMB || op == FQ_DEVOP_DH_BASE || op == FQ_DEVOP_MUL_BASE ||
         op == FQ_DEVOP_DH_ENDO_BASE || op == FQ_DEVOP_MUL_ENDO_BASE;
}
bool needs_scratch(int op) { return is_dh_op(op) || is_comb_op(op) || op == FQ_DEVOP_X25519; }

int strict_mode() {
  if (g_strict < 0) { const char* e = getenv("FQ_STRICT_SELECT"); g_strict = (e && e[0] == '1') ? 1 : 0; }
  return g_strict;
}

int hslot_reserve(Slot& s, int which, size_t bytes) {
  if (bytes <= s.hcap[which]) return FQ_OK;
  if (s.hbuf[which]) CU(cudaFreeHost(s.hbuf[which]));
  s.hbuf[which] = nullptr; s.hcap[which] = 0;
  CU(cudaHostAlloc(&s.hbuf[which], bytes, cudaHostAllocPortable));
  s.hcap[which] = bytes;
  return FQ_OK;
}
// true if p is page-locked host memory (cudaHostAlloc / cudaHostRegister): asynchronous copies can use it directly
bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// grows the kernel scratch of stream slot `si` to what `op` needs for `rows` rows
int dh_scratch_reserve(DevCtx& c, int si, int op, size_t rows) {
  size_t bytes = is_comb_op(op) ? fqk_comb_scratch_bytes(rows) : op == FQ_DEVOP_X25519 ? fqk_x25519_scratch_bytes(rows) : fqk_dh_scratch_bytes(rows);
  if (bytes <= c.dh_scratch_cap[si]) return FQ_OK;
  if (c.dh_scratch[si]) CU(cudaFree(c.dh_scratch[si]));
  c.dh_scratch[si] = nullptr; c.dh_scratch_cap[si] = 0;
  CU(cudaMalloc(&c.dh_scratch[si], bytes));
  c.dh_scratch_cap[si] = bytes;
  return FQ_OK;
}
// bytes per row of each operand of an operation
struct OpDesc { int op; int a_bytes, b_bytes, out_bytes; bool status; size_t chunk_rows; };

// rows per full pipeline chunk of the variable-base DH ops: a whole number of waves of k_dh_ladder (2 CTAs x 128 rows per SM)
// and of k_dh_prep (3 per SM) keeps the tail of each chunk short; FQ_DH_CHUNK_ROWS overrides it (tuning knob, read once).
size_t dh_chunk_rows() {
  static size_t v = 0;
  if (v == 0) {
    const char* e = getenv("FQ_DH_CHUNK_ROWS");
    long long x = e ? atoll(e) : 0;
    v = x >= 128 ? (size_t)x : (size_t)148 * 2 * 128 * 12;      // 454,656 rows = 12 waves of k_dh_ladder, 8 of k_dh_prep
  }
  return v;
}

// FQ_PIPELINE_RAMP=0 disables the ramped chunk schedule of run_host (all chunks full size)
bool ramp_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FQ_PIPELINE_RAMP"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

OpDesc describe(int op) {
  switch (op) {
    case FQ_DEVOP_FP2_MUL: case FQ_DEVOP_FP2_ADD: case FQ_DEVOP_FP2_SUB: return {op, 32, 32, 32, false, (size_t)1 << 20};
    case FQ_DEVOP_FP2_SQR: case FQ_DEVOP_FP2_NEG: case FQ_DEVOP_FP2_CONJ: return {op, 32, 0, 32, false, (size_t)1 << 20};
    case FQ_DEVOP_FP2_INV: return {op, 32, 0, 32, false, (size_t)1 << 18};
    case FQ_DEVOP_FP_BASE + FQ_FP_MUL: case FQ_DEVOP_FP_BASE + FQ_FP_ADD: case FQ_DEVOP_FP_BASE + FQ_FP_SUB: return {op, 16, 16, 16, false, (size_t)1 << 20};
    case FQ_DEVOP_FP_BASE + FQ_FP_SQR: case FQ_DEVOP_FP_BASE + FQ_FP_NEG: return {op, 16, 0, 16, false, (size_t)1 << 20};
    case FQ_DEVOP_FP_BASE + FQ_FP_INV: case FQ_DEVOP_FP_BASE + FQ_FP_INVSQRT: return {op, 16, 0, 16, false, (size_t)1 << 18};
    case FQ_DEVOP_DECODE: case FQ_DEVOP_DECODE_SPEC: return {op, 32, 0, 64, true, (size_t)1 << 18};
    case FQ_DEVOP_ENCODE: return {op, 64, 0, 32, false, (size_t)1 << 20};
    case FQ_DEVOP_ON_CURVE: return {op, 64, 0, 1, false, (size_t)1 << 20};
    case FQ_DEVOP_DH: case FQ_DEVOP_DH_ENDO: return {op, 32, 32, 32, true, dh_chunk_rows()};
    case FQ_DEVOP_DH_AFFINE: case FQ_DEVOP_DH_ENDO_AFFINE: return {op, 32, 64, 64, true, dh_chunk_rows()};
    case FQ_DEVOP_DH_BASE: case FQ_DEVOP_DH_ENDO_BASE: return {op, 32, 0, 32, true, (size_t)1 << 17};
    case FQ_DEVOP_MUL_BASE: case FQ_DEVOP_MUL_ENDO_BASE: return {op, 32, 0, 32, false, (size_t)1 << 17};
    // k_comb keeps 2 CTAs of 256 rows resident per SM, each walking over tiles: 148 x 2 x 256 x 4 rows = 4 full tiles per CTA
    case FQ_DEVOP_DH_BASE_COMB: return {op, 32, 0, 32, true, (size_t)148 * 2 * 256 * 4};
    case FQ_DEVOP_MUL_BASE_COMB: return {op, 32, 0, 32, false, (size_t)148 * 2 * 256 * 4};
    case FQ_DEVOP_X25519: return {op, 32, 32, 32, false, (size_t)1 << 17};
    default: return {-1, 0, 0, 0, false, 0};
  }
}

// si: stream slot whose DH scratch is used (reserved by the caller); ev: optional per-kernel events of the DH pipeline
cudaError_t launch(const DevCtx& cx, int op, const void* a, const void* b, void* out, void* status, size_t n, cudaStream_t s, int si = 0, cudaEvent_t* ev = nullptr) {
  switch (op) {
    case FQ_DEVOP_DH_BASE_COMB: return fqk_comb(1, strict_mode(), cx.comb, a, out, status, n, cx.dh_scratch[si], cx.sms, s);
    case FQ_DEVOP_MUL_BASE_COMB: return fqk_comb(0, strict_mode(), cx.comb, a, out, nullptr, n, cx.dh_scratch[si], cx.sms, s);
    case FQ_DEVOP_FP2_MUL: return fqk_fp2_op(FQK_MUL, a, b, out, n, s);
    case FQ_DEVOP_FP2_SQR: return fqk_fp2_op(FQK_SQR, a, b, out, n, s);
    case FQ_DEVOP_FP2_INV: return fqk_fp2_op(FQK_INV, a, b, out, n, s);
    case FQ_DEVOP_FP2_ADD: return fqk_fp2_op(FQK_ADD, a, b, out, n, s);
    case FQ_DEVOP_FP2_SUB: return fqk_fp2_op(FQK_SUB, a, b, out, n, s);
    case FQ_DEVOP_FP2_NEG: return fqk_fp2_op(FQK_NEG, a, b, out, n, s);
    case FQ_DEVOP_FP2_CONJ: return fqk_fp2_op(FQK_CONJ, a, b, out, n, s);
    case FQ_DEVOP_FP_BASE + FQ_FP_MUL: case FQ_DEVOP_FP_BASE + FQ_FP_SQR: case FQ_DEVOP_FP_BASE + FQ_FP_INV: case FQ_DEVOP_FP_BASE + FQ_FP_ADD:
    case FQ_DEVOP_FP_BASE + FQ_FP_SUB: case FQ_DEVOP_FP_BASE + FQ_FP_NEG: case FQ_DEVOP_FP_BASE + FQ_FP_INVSQRT:
      return fqk_fp_op(op - FQ_DEVOP_FP_BASE, a, b, out, n, s);
    case FQ_DEVOP_DECODE: return fqk_decode(0, a, out, status, n, s);
    case FQ_DEVOP_DECODE_SPEC: return fqk_decode(1, a, out, status, n, s);
    case FQ_DEVOP_ENCODE: return fqk_encode(a, out, n, s);
    case FQ_DEVOP_ON_CURVE: return fqk_on_curve(a, out, n, s);
    case FQ_DEVOP_DH: return fqk_dh(0, 0, strict_mode(), a, b, out, status, n, cx.dh_scratch[si], s, ev);
    case FQ_DEVOP_DH_AFFINE: return fqk_dh(1, 0, strict_mode(), a, b, out, status, n, cx.dh_scratch[si], s, ev);
    case FQ_DEVOP_DH_BASE: return fqk_fixed_base(1, 0, strict_mode(), a, out, status, n, cx.dh_scratch[si], s);
    case FQ_DEVOP_MUL_BASE: return fqk_fixed_base(0, 0, strict_mode(), a, out, nullptr, n, cx.dh_scratch[si], s);
    case FQ_DEVOP_DH_ENDO: return fqk_dh(0, 1, strict_mode(), a, b, out, status, n, cx.dh_scratch[si], s, ev);
    case FQ_DEVOP_DH_ENDO_AFFINE: return fqk_dh(1, 1, strict_mode(), a, b, out, status, n, cx.dh_scratch[si], s, ev);
    case FQ_DEVOP_DH_ENDO_BASE: return fqk_fixed_base(1, 1, strict_mode(), a, out, status, n, cx.dh_scratch[si], s);
    case FQ_DEVOP_MUL_ENDO_BASE: return fqk_fixed_base(0, 1, strict_mode(), a, out, nullptr, n, cx.dh_scratch[si], s);
    case FQ_DEVOP_X25519: return fqk_x25519(a, b, out, n, cx.dh_scratch[si], s);
    default: return cudaErrorInvalidValue;
  }
}

struct ChunkEv { int dev; cudaEvent_t e0, e1; };

int run_host(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint8_t* status, size_t n, int ndev) {
  std::lock_guard<std::mutex> lock(g_mu);
  tl_kernel_ms = 0.f;
  OpDesc d = describe(op);
  if (d.op < 0) return fail(FQ_ERR_ARG, "unknown operation %d", op);
  if (n == 0) return FQ_OK;
  if (!a || !out || (d.b_bytes && !b) || (d.status && !status)) return fail(FQ_ERR_ARG, "null buffer");
  int count = device_count();
  if (count < 0) return count;
  if (ndev < 1 || g_dev_base + ndev > count) return fail(FQ_ERR_ARG, "ndev=%d with device base %d but %d device(s) visible", ndev, g_dev_base, count);

  size_t per = (n + ndev - 1) / ndev;
  std::vector<ChunkEv> evs;
  // Chunk schedule of one slice (the same for every device): ramp up from a small chunk (doubling) so that the first kernels
  // start after a short copy, full chunks in the middle, ramp down by halves so that little work and a short copy-back remain
  // exposed at the end.  Batches of at most two small chunks are not split.
  std::vector<size_t> bounds;                        // chunk c covers [bounds[c], bounds[c+1]) of the slice
  {
    const size_t full = d.chunk_rows;
    size_t small = full / 8 >= 16384 ? full / 8 : 16384;
    if (small > full) small = full;
    const bool ramp = ramp_enabled() && per > 2 * small;
    size_t pos = 0, sz = ramp ? small : full;
    bounds.push_back(0);
    while (pos < per) {
      size_t left = per - pos;
      size_t take = sz < left ? sz : left;
      if (ramp && left > small && left <= 2 * take) take = (left / 2 + 127) / 128 * 128;   // ramp down by halves
      pos += take; bounds.push_back(pos);
      if (sz < full) sz = sz * 2 < full ? sz * 2 : full;
    }
  }
  size_t max_chunks = bounds.size() - 1;
  // Results for pageable host buffers go through pinned staging buffers of the stream slot: the D2H copy is asynchronous, and
  // the host thread copies a chunk's results to the caller's memory once its stream has finished -- at the latest when the
  // slot is needed again three chunks later -- so these copies (and the page faults of a freshly allocated output array)
  // overlap the kernels of the chunks in between.  Left to the driver, D2H into pageable memory is staged synchronously:
  // 39 M instead of 73 M DH rows/s with plain numpy arrays.  Pageable INPUTS are left to the driver (staging them here measured
  // no better); page-locked buffers (fq_host_alloc, pinned_empty) are used directly in both directions.
  const bool pin_o = is_pinned(out), pin_s = !d.status || is_pinned(status);
  // results of the slot's previous chunk: wait for its stream, then staging -> caller's memory
  auto drain = [&](DevCtx& cx, int si) -> int {
    Slot& s = cx.slot[si];
    if (!s.pending) return FQ_OK;
    s.pending = false;
    CU(cudaStreamSynchronize(cx.st[si]));
    if (!pin_o) memcpy(out + s.p_r0 * d.out_bytes, s.hbuf[2], s.p_rows * d.out_bytes);
    if (d.status && !pin_s) memcpy(status + s.p_r0, s.hbuf[3], s.p_rows);
    return FQ_OK;
  };
  // one chunk of one device: copies in, kernels, copies out on the chunk's stream.  A failing call returns its error from the
  // lambda; the caller then stops enqueueing and still waits for and cleans up everything that is already in flight.
  auto enqueue = [&](size_t c, int i) -> int {
    size_t lo = (size_t)i * per, hi = lo + per < n ? lo + per : n;
    size_t r0 = lo + bounds[c];
    if (lo >= n || r0 >= hi) return FQ_OK;
    size_t r1 = lo + bounds[c + 1] < hi ? lo + bounds[c + 1] : hi;
    size_t rows = r1 - r0;
    int dev = g_dev_base + i, e;
    if ((e = ctx_init(dev)) != FQ_OK) return e;
    DevCtx& cx = g_ctx[dev];
    int si = (int)(c % kStreams);
    Slot& s = cx.slot[si];
    cudaStream_t st = cx.st[si];
    if ((e = drain(cx, si)) != FQ_OK) return e;
    if ((e = slot_reserve(s, 0, d.chunk_rows * d.a_bytes)) != FQ_OK) return e;
    if (d.b_bytes && (e = slot_reserve(s, 1, d.chunk_rows * d.b_bytes)) != FQ_OK) return e;
    if ((e = slot_reserve(s, 2, d.chunk_rows * d.out_bytes)) != FQ_OK) return e;
    if (d.status && (e = slot_reserve(s, 3, d.chunk_rows)) != FQ_OK) return e;
    if (needs_scratch(op) && (e = dh_scratch_reserve(cx, si, op, d.chunk_rows)) != FQ_OK) return e;
    if (!pin_o && (e = hslot_reserve(s, 2, d.chunk_rows * d.out_bytes)) != FQ_OK) return e;
    if (d.status && !pin_s && (e = hslot_reserve(s, 3, d.chunk_rows)) != FQ_OK) return e;
    CU(cudaMemcpyAsync(s.buf[0], a + r0 * d.a_bytes, rows * d.a_bytes, cudaMemcpyHostToDevice, st));
    if (d.b_bytes) CU(cudaMemcpyAsync(s.buf[1], b + r0 * d.b_bytes, rows * d.b_bytes, cudaMemcpyHostToDevice, st));
    ChunkEv ev; ev.dev = i;
    CU(cudaEventCreate(&ev.e0));
    if (cudaEventCreate(&ev.e1) != cudaSuccess) { cudaEventDestroy(ev.e0); return fail(FQ_ERR_CUDA, "cudaEventCreate failed"); }
    evs.push_back(ev);                               // from here on the events are destroyed by the common cleanup
    CU(cudaEventRecord(ev.e0, st));
    CU(launch(cx, op, s.buf[0], s.buf[1], s.buf[2], s.buf[3], rows, st, si));
    CU(cudaEventRecord(ev.e1, st));
    CU(cudaMemcpyAsync(pin_o ? (void*)(out + r0 * d.out_bytes) : s.hbuf[2], s.buf[2], rows * d.out_bytes, cudaMemcpyDeviceToHost, st));
    if (d.status) CU(cudaMemcpyAsync(pin_s ? (void*)(status + r0) : s.hbuf[3], s.buf[3], rows, cudaMemcpyDeviceToHost, st));
    if (!pin_o || !pin_s) { s.pending = true; s.p_r0 = r0; s.p_rows = rows; }
    return FQ_OK;
  };
  int rc = FQ_OK;
  for (size_t c = 0; c < max_chunks && rc == FQ_OK; c++)
    for (int i = 0; i < ndev && rc == FQ_OK; i++) rc = enqueue(c, i);
  // results still in staging buffers (in enqueue order, so that the oldest chunk of each device is copied out first)
  for (size_t c = max_chunks >= (size_t)kStreams ? max_chunks - kStreams : 0; c < max_chunks; c++)
    for (int i = 0; i < ndev; i++) {
      int dev = g_dev_base + i;
      if (!g_ctx[dev].ready) continue;
      cudaSetDevice(dev);
      int e = drain(g_ctx[dev], (int)(c % kStreams));
      if (e != FQ_OK && rc == FQ_OK) rc = e;
    }
  // wait for every device, then collect kernel times
  for (int i = 0; i < ndev; i++) {
    int dev = g_dev_base + i;
    if (!g_ctx[dev].ready) continue;
    cudaSetDevice(dev);
    for (int s = 0; s < kStreams; s++) {
      cudaError_t e = cudaStreamSynchronize(g_ctx[dev].st[s]);
      if (e != cudaSuccess && rc == FQ_OK) rc = fail(FQ_ERR_CUDA, "stream sync on device %d failed: %s", dev, cudaGetErrorString(e));
    }
  }
  float per_dev[kMaxDev] = {0};
  for (auto& ev : evs) {
    float ms = 0.f;
    if (rc == FQ_OK && cudaEventElapsedTime(&ms, ev.e0, ev.e1) == cudaSuccess) per_dev[ev.dev] += ms;
    cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1);
  }
  for (int i = 0; i < ndev; i++) if (per_dev[i] > tl_kernel_ms) tl_kernel_ms = per_dev[i];
  return rc;
}

}  // namespace

extern "C" {

int fq_version(void) { return FQ_VERSION; }
int fq_device_count(void) { return device_count(); }
const char* fq_last_error(void) { return tl_err; }
float fq_last_kernel_ms(void) { return tl_kernel_ms; }

int fq_set_device_base(int first) {
  std::lock_guard<std::mutex> lock(g_mu);
  int count = device_count();
  if (count < 0) return count;
  if (first < 0 || first >= count) return fail(FQ_ERR_ARG, "device base %d out of range (%d device(s))", first, count);
  g_dev_base = first;
  return FQ_OK;
}

int fq_trim(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  int count = device_count();
  if (count < 0) return count;
  for (int dev = 0; dev < count && dev < kMaxDev; dev++) {
    DevCtx& c = g_ctx[dev];
    if (!c.ready) continue;
    CU(cudaSetDevice(dev));
    CU(cudaDeviceSynchronize());
    for (int si = 0; si < kStreams; si++) {
      Slot& s = c.slot[si];
      for (int w = 0; w < 4; w++) {
        if (s.buf[w]) { CU(cudaFree(s.buf[w])); s.buf[w] = nullptr; s.cap[w] = 0; }
        if (s.hbuf[w]) { CU(cudaFreeHost(s.hbuf[w])); s.hbuf[w] = nullptr; s.hcap[w] = 0; }
      }
      if (c.dh_scratch[si]) { CU(cudaFree(c.dh_scratch[si])); c.dh_scratch[si] = nullptr; c.dh_scratch_cap[si] = 0; }
    }
    if (c.flush) { CU(cudaFree(c.flush)); c.flush = nullptr; }
  }
  return FQ_OK;
}

int fq_set_select_mode(int strict) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (strict != 0 && strict != 1) return fail(FQ_ERR_ARG, "select mode must be 0 (masked loads) or 1 (strict scan)");
  g_strict = strict;
  return FQ_OK;
}
int fq_get_select_mode(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  return strict_mode();
}

int fq_fp2_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_MUL, a, b, out, nullptr, n, ndev); }
int fq_fp2_sqr(const uint8_t* a, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_SQR, a, nullptr, out, nullptr, n, ndev); }
int fq_fp2_inv(const uint8_t* a, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_INV, a, nullptr, out, nullptr, n, ndev); }
int fq_fp2_add(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_ADD, a, b, out, nullptr, n, ndev); }
int fq_fp2_sub(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_SUB, a, b, out, nullptr, n, ndev); }
int fq_fp2_neg(const uint8_t* a, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_NEG, a, nullptr, out, nullptr, n, ndev); }
int fq_fp2_conj(const uint8_t* a, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_FP2_CONJ, a, nullptr, out, nullptr, n, ndev); }
int fq_fp_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, int ndev) {
  if (op < FQ_FP_MUL || op > FQ_FP_INVSQRT) return fail(FQ_ERR_ARG, "unknown GF(p) operation %d", op);
  return run_host(FQ_DEVOP_FP_BASE + op, a, b, out, nullptr, n, ndev);
}
int fq_decode(const uint8_t* enc, uint8_t* xy, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DECODE, enc, nullptr, xy, status, n, ndev); }
int fq_point_on_curve(const uint8_t* xy, uint8_t* ok, size_t n, int ndev) { return run_host(FQ_DEVOP_ON_CURVE, xy, nullptr, ok, nullptr, n, ndev); }
int fq_decode_spec(const uint8_t* enc, uint8_t* xy, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DECODE_SPEC, enc, nullptr, xy, status, n, ndev); }
int fq_encode(const uint8_t* xy, uint8_t* enc, size_t n, int ndev) { return run_host(FQ_DEVOP_ENCODE, xy, nullptr, enc, nullptr, n, ndev); }
int fq_dh(const uint8_t* k, const uint8_t* enc_pt, uint8_t* enc_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH, k, enc_pt, enc_out, status, n, ndev); }
int fq_dh_affine(const uint8_t* k, const uint8_t* xy, uint8_t* xy_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_AFFINE, k, xy, xy_out, status, n, ndev); }
int fq_dh_base(const uint8_t* k, uint8_t* enc_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_BASE, k, nullptr, enc_out, status, n, ndev); }
int fq_mul_base(const uint8_t* k, uint8_t* enc_out, size_t n, int ndev) { return run_host(FQ_DEVOP_MUL_BASE, k, nullptr, enc_out, nullptr, n, ndev); }

int fq_dh_endo(const uint8_t* k, const uint8_t* enc_pt, uint8_t* enc_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_ENDO, k, enc_pt, enc_out, status, n, ndev); }
int fq_dh_endo_affine(const uint8_t* k, const uint8_t* xy, uint8_t* xy_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_ENDO_AFFINE, k, xy, xy_out, status, n, ndev); }
int fq_dh_endo_base(const uint8_t* k, uint8_t* enc_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_ENDO_BASE, k, nullptr, enc_out, status, n, ndev); }
int fq_mul_endo_base(const uint8_t* k, uint8_t* enc_out, size_t n, int ndev) { return run_host(FQ_DEVOP_MUL_ENDO_BASE, k, nullptr, enc_out, nullptr, n, ndev); }
int fq_dh_base_comb(const uint8_t* k, uint8_t* enc_out, uint8_t* status, size_t n, int ndev) { return run_host(FQ_DEVOP_DH_BASE_COMB, k, nullptr, enc_out, status, n, ndev); }
int fq_mul_base_comb(const uint8_t* k, uint8_t* enc_out, size_t n, int ndev) { return run_host(FQ_DEVOP_MUL_BASE_COMB, k, nullptr, enc_out, nullptr, n, ndev); }
int fq_x25519(const uint8_t* k, const uint8_t* u, uint8_t* out, size_t n, int ndev) { return run_host(FQ_DEVOP_X25519, k, u, out, nullptr, n, ndev); }

int fq_host_alloc(void** p, size_t bytes) {
  if (!p) return fail(FQ_ERR_ARG, "null pointer");
  int count = device_count();
  if (count < 0) return count;
  CU(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
  return FQ_OK;
}
int fq_host_free(void* p) { if (p) CU(cudaFreeHost(p)); return FQ_OK; }

static int dev_enter(int dev) {
  int count = device_count();
  if (count < 0) return count;
  if (dev < 0 || dev >= count) return fail(FQ_ERR_ARG, "device %d out of range (%d device(s))", dev, count);
  return ctx_init(dev);
}
int fq_dev_alloc(int dev, void** p, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  if (!p) return fail(FQ_ERR_ARG, "null pointer");
  CU(cudaMalloc(p, bytes ? bytes : 1));
  return FQ_OK;
}
int fq_dev_free(int dev, void* p) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  if (p) CU(cudaFree(p));
  return FQ_OK;
}
int fq_dev_upload(int dev, void* dst, const void* src, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return FQ_OK;
}
int fq_dev_download(int dev, void* dst, const void* src, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return FQ_OK;
}
int fq_dev_run(int op, int dev, const void* a, const void* b, void* out, void* status, size_t n, int iters, float* ms) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  if (describe(op).op < 0 || iters < 1) return fail(FQ_ERR_ARG, "bad op/iters");
  DevCtx& cx = g_ctx[dev];
  cudaStream_t st = cx.st[0];
  const bool dh = is_dh_op(op);
  if (needs_scratch(op) && (rc = dh_scratch_reserve(cx, 0, op, n)) != FQ_OK) return rc;
  cudaEvent_t e0, e1, ph[4];
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  for (int i = 0; i < 4; i++) CU(cudaEventCreate(&ph[i]));
  CU(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; i++) CU(launch(cx, op, a, b, out, status, n, st, 0, (dh && i == iters - 1) ? ph : nullptr));
  CU(cudaEventRecord(e1, st));
  CU(cudaEventSynchronize(e1));
  float t = 0.f;
  CU(cudaEventElapsedTime(&t, e0, e1));
  tl_phase_ms[0] = tl_phase_ms[1] = tl_phase_ms[2] = 0.f;
  if (dh && n > 0) for (int i = 0; i < 3; i++) CU(cudaEventElapsedTime(&tl_phase_ms[i], ph[i], ph[i + 1]));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (int i = 0; i < 4; i++) cudaEventDestroy(ph[i]);
  if (ms) *ms = t / iters;
  return FQ_OK;
}
int fq_dev_last_phase_ms(float* ms3) {
  if (!ms3) return fail(FQ_ERR_ARG, "null pointer");
  for (int i = 0; i < 3; i++) ms3[i] = tl_phase_ms[i];
  return FQ_OK;
}
int fq_dev_flush_l2(int dev) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  DevCtx& c = g_ctx[dev];
  if (!c.flush) CU(cudaMalloc(&c.flush, kFlushBytes));
  CU(cudaMemsetAsync(c.flush, 0, kFlushBytes, c.st[0]));
  CU(cudaStreamSynchronize(c.st[0]));
  return FQ_OK;
}

int fq_imad_peak(int dev, double* wide_per_s, double* imad32_per_s) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = dev_enter(dev); if (rc != FQ_OK) return rc;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  int blocks = prop.multiProcessorCount * 4, trips = 8192;
  void* scratch = nullptr;
  CU(cudaMalloc(&scratch, (size_t)blocks * 256 * 4));
  cudaStream_t st = g_ctx[dev].st[0];
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double res[2] = {0, 0};
  for (int v = 0; v < 2; v++) {
    CU(fqk_imad_peak(v, scratch, blocks, trips / 8, st));     // warm-up
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      CU(cudaEventRecord(e0, st));
      CU(fqk_imad_peak(v, scratch, blocks, trips, st));
      CU(cudaEventRecord(e1, st));
      CU(cudaEventSynchronize(e1));
      float t; CU(cudaEventElapsedTime(&t, e0, e1));
      if (t < best) best = t;
    }
    res[v] = (double)blocks * 256.0 * trips * 128.0 / (best * 1e-3);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CU(cudaFree(scratch));
  if (wide_per_s) *wide_per_s = res[0];
  if (imad32_per_s) *imad32_per_s = res[1];
  return FQ_OK;
}

}  // extern "C"
