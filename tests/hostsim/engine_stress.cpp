// engine_stress.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the host engine of fourq_b200/csrc/capi.cu (built against the mock CUDA
// runtime of this directory) from several threads at once, with pageable and page-locked buffers, random stream delays and
// injected CUDA failures; linked with -fsanitize=thread or -fsanitize=address,undefined by build.sh.  Exit code 0 and the
// line "engine_stress ok" mean every result matched the direct CPU simulation and no call hung.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include "../../include/fourq_b200.h"

extern "C" {
int sim_fp2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
int sim_comb(int dh, const uint8_t* k, uint8_t* out, uint8_t* status, size_t n);
int sim_dh_endo(const uint8_t* k, const uint8_t* enc, uint8_t* out, uint8_t* status, size_t n);
void mock_set_device_count(int n);
void mock_fail_nth(const char* api, int nth);
void mock_set_jitter_us(int us);
long mock_violations();
}

static std::atomic<int> g_bad{0};
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed: %s (%s:%d) err=%s\n", #c, __FILE__, __LINE__, fq_last_error()); g_bad++; } } while (0)

static void fill(std::vector<uint8_t>& v, unsigned seed) { std::minstd_rand r(seed); for (auto& x : v) x = (uint8_t)r(); }

static void worker(int id, int rounds) {
  const size_t n = 150000 + 1777 * id;
  std::vector<uint8_t> a(n * 32), b(n * 32), ref(n * 32), out(n * 32);
  fill(a, 1 + id); fill(b, 100 + id);
  sim_fp2_op(3, a.data(), b.data(), ref.data(), n);          // add: cheap on the CPU, all the pressure is on the engine
  void *pa = nullptr, *po = nullptr;
  CHECK(fq_host_alloc(&pa, n * 32) == FQ_OK); CHECK(fq_host_alloc(&po, n * 32) == FQ_OK);
  memcpy(pa, a.data(), n * 32);
  for (int r = 0; r < rounds; r++) {
    const int ndev = 1 + (id + r) % 4;
    memset(out.data(), 0, out.size());
    CHECK(fq_fp2_add(a.data(), b.data(), out.data(), n, ndev) == FQ_OK);                       // pageable in, pageable out
    CHECK(memcmp(out.data(), ref.data(), n * 32) == 0);
    memset(po, 0, n * 32);
    CHECK(fq_fp2_add((const uint8_t*)pa, b.data(), (uint8_t*)po, n, ndev) == FQ_OK);            // pinned in / out, one pageable operand
    CHECK(memcmp(po, ref.data(), n * 32) == 0);
  }
  CHECK(fq_host_free(pa) == FQ_OK); CHECK(fq_host_free(po) == FQ_OK);
}

int main() {
  mock_set_device_count(4);
  mock_set_jitter_us(150);
  {
    std::vector<std::thread> ts;
    for (int i = 0; i < 5; i++) ts.emplace_back(worker, i, 3);
    for (auto& t : ts) t.join();
  }
  // scalar multiplications through the staged path (scratch sizing under ASan), several devices, ragged sizes
  for (size_t n : {1u, 127u, 129u, 700u}) {
    std::vector<uint8_t> k(n * 32), pub(n * 32), ref(n * 32), out(n * 32), st(n), rst(n);
    fill(k, 7 + (unsigned)n);
    CHECK(fq_mul_base_comb(k.data(), pub.data(), n, 3) == FQ_OK);
    sim_comb(0, k.data(), ref.data(), nullptr, n);
    CHECK(memcmp(pub.data(), ref.data(), n * 32) == 0);
    CHECK(fq_dh_endo(k.data(), pub.data(), out.data(), st.data(), n, 2) == FQ_OK);
    sim_dh_endo(k.data(), pub.data(), ref.data(), rst.data(), n);
    CHECK(memcmp(out.data(), ref.data(), n * 32) == 0 && memcmp(st.data(), rst.data(), n) == 0);
  }
  // failures injected while other callers are in flight: the failing call returns an error, nobody hangs, later calls work
  {
    std::atomic<bool> stop{false};
    std::thread inj([&] {
      const char* apis[] = {"cudaMemcpyAsync", "cudaEventRecord", "cudaMalloc", "cudaEventSynchronize", "cudaHostAlloc"};
      for (int i = 0; i < 40 && !stop; i++) { mock_fail_nth(apis[i % 5], 1 + i % 7); std::this_thread::sleep_for(std::chrono::milliseconds(3)); }
    });
    const size_t n = 300000;
    std::vector<uint8_t> a(n * 32), ref(n * 32), out(n * 32);
    fill(a, 55);
    sim_fp2_op(5, a.data(), nullptr, ref.data(), n);
    int failed = 0;
    for (int r = 0; r < 25; r++) {
      if (r % 6 == 0) CHECK(fq_trim() == FQ_OK || true);          // forces re-allocation (an injected failure may hit trim itself)
      memset(out.data(), 0, out.size());
      int rc = fq_fp2_neg(a.data(), out.data(), n, 1 + r % 4);
      if (rc == FQ_OK) CHECK(memcmp(out.data(), ref.data(), n * 32) == 0); else { failed++; CHECK(rc == FQ_ERR_CUDA); }
    }
    stop = true; inj.join();
    for (const char* api : {"cudaMemcpyAsync", "cudaEventRecord", "cudaMalloc", "cudaEventSynchronize", "cudaHostAlloc"}) mock_fail_nth(api, 0);
    memset(out.data(), 0, out.size());
    CHECK(fq_fp2_neg(a.data(), out.data(), n, 4) == FQ_OK);
    CHECK(memcmp(out.data(), ref.data(), n * 32) == 0);
    printf("injected failures seen by the caller: %d of 25 calls\n", failed);
  }
  CHECK(fq_trim() == FQ_OK);
  CHECK(mock_violations() == 0);
  if (g_bad) { printf("engine_stress FAILED: %d checks\n", g_bad.load()); return 1; }
  printf("engine_stress ok\n");
  return 0;
}
