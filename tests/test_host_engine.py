"""The host engine of fourq_b200/csrc/capi.cu (slices, chunk schedule, pinned staging, the per-GPU feeder / drainer threads,
error paths) exercised WITHOUT a GPU: tests/hostsim builds capi.cu against a mock CUDA runtime whose streams are threads and
whose kernels are the CPU simulation of the device code.  Test infrastructure only -- the product library is never built
this way and has no CPU path."""
import ctypes
import os
import subprocess
import threading

import numpy as np
import pytest

from fourq_b200 import _lib
from oracle import c_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_DIR = os.path.join(HERE, "hostsim")
CSRC = os.path.join(HERE, "..", "fourq_b200", "csrc")


def _stale(target, extra=()):
    srcs = [os.path.join(SIM_DIR, f) for f in os.listdir(SIM_DIR) if f.endswith((".cpp", ".h", ".sh"))]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")) or f == "capi.cu"] + list(extra)
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(s) for s in srcs)


@pytest.fixture(scope="module")
def eng():
    so = os.path.join(SIM_DIR, "libfq_mockengine.so")
    if _stale(so):
        subprocess.check_call(["sh", os.path.join(SIM_DIR, "build.sh")])
    L = _lib.bind(ctypes.CDLL(so))
    for name in ("mock_violations", "mock_memcpy_bytes"):
        getattr(L, name).restype = ctypes.c_long
    L.mock_live_allocs.restype = ctypes.c_long
    L.mock_live_allocs.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
    L.mock_fail_nth.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.fq_test_chunk_schedule.restype = ctypes.c_size_t
    L.fq_test_chunk_schedule.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t]
    L.mock_set_device_count(4)
    yield L
    L.mock_set_jitter_us(0)


@pytest.fixture(scope="module")
def sim():
    return ctypes.CDLL(os.path.join(SIM_DIR, "libfq_hostsim.so"))


def pinned(L, shape):
    n = int(np.prod(shape))
    p = ctypes.c_void_p()
    assert L.fq_host_alloc(ctypes.byref(p), n) == 0
    arr = np.frombuffer((ctypes.c_uint8 * n).from_address(p.value), np.uint8).reshape(shape)
    return arr, p


P = _lib.ptr


def test_chunk_schedule_never_exceeds_the_staging_size(eng):
    """with a chunk size that is not a multiple of 128 the ramp-down used to round a chunk past the buffers."""
    buf = (ctypes.c_size_t * 40000)()
    for full in (128, 1000, 16384, 100000, 131072, 303104, 454656, 500000, 1 << 20):
        for rows in (1, 127, 128, 129, full - 1, full, full + 1, 2 * full + 5, 3 * full + 77, 10 * full + 1, (1 << 21) + 3, (1 << 22) + 12345):
            if rows < 1:
                continue
            m = eng.fq_test_chunk_schedule(rows, full, buf, 40000)
            b = [buf[i] for i in range(m)]
            assert b[0] == 0 and b[-1] == rows and all(0 < y - x <= full for x, y in zip(b, b[1:])), (full, rows, b)


@pytest.mark.parametrize("ndev", [1, 2, 3, 4])
def test_fp2_mul_many_chunks_pageable_and_pinned(eng, sim, ndev):
    eng.mock_set_jitter_us(200)
    rng = np.random.default_rng(10 + ndev)
    n = 700_001 + 4099 * ndev                     # several ramped chunks per slice (full chunk = 2^20 rows, small = 2^17)
    a = rng.integers(0, 256, (n, 32), np.uint8); b = rng.integers(0, 256, (n, 32), np.uint8)
    ref = np.empty_like(a)
    sim.sim_fp2_op(0, P(a), P(b), P(ref), ctypes.c_size_t(n))
    out = np.zeros_like(a)
    assert eng.fq_fp2_mul(P(a), P(b), P(out), n, ndev) == 0, eng.fq_last_error()
    assert (out == ref).all()
    assert eng.fq_last_kernel_ms() > 0
    # page-locked buffers are used directly: no staging copies of the results
    pa, ha = pinned(eng, (n, 32)); pb, hb = pinned(eng, (n, 32)); po, ho = pinned(eng, (n, 32))
    pa[:] = a; pb[:] = b; po[:] = 0
    assert eng.fq_fp2_mul(ha, hb, ho, n, ndev) == 0, eng.fq_last_error()
    assert (po == ref).all()
    # mixed: pinned inputs, pageable output
    out[:] = 0
    assert eng.fq_fp2_mul(ha, hb, P(out), n, ndev) == 0
    assert (out == ref).all()
    for h in (ha, hb, ho):
        assert eng.fq_host_free(h) == 0
    assert eng.mock_violations() == 0


def test_dh_slices_match_single_device_and_the_oracle(eng):
    eng.mock_set_jitter_us(0)
    rng = np.random.default_rng(5)
    n = 1203
    k = rng.integers(0, 256, (n, 32), np.uint8)
    pub = np.empty_like(k)
    assert eng.fq_mul_base_comb(P(rng.integers(0, 256, (n, 32), np.uint8)), P(pub), n, 4) == 0
    pub[7] = 0xFF; pub[600] = 0                      # failing rows travel through the slices too
    want, wst = c_oracle.dh(k, pub)
    for fn in (eng.fq_dh_endo, eng.fq_dh):
        for ndev in (1, 3):
            out = np.zeros_like(k); st = np.full(n, 9, np.uint8)
            assert fn(P(k), P(pub), P(out), P(st), n, ndev) == 0, eng.fq_last_error()
            assert (out == want).all() and (st == wst).all()
    assert eng.mock_violations() == 0


def test_three_input_select_through_the_engine(eng):
    rng = np.random.default_rng(6)
    n = 5000
    x = rng.integers(0, 256, (n, 32), np.uint8); y = rng.integers(0, 256, (n, 32), np.uint8)
    c = rng.integers(0, 2, n, np.uint8)
    out = np.empty_like(x)
    assert eng.fq_fp2_select(P(c), P(x), P(y), P(out), n, 2) == 0
    assert (out == np.where(c.reshape(-1, 1) == 1, x, y)).all()
    out16 = np.empty((n, 16), np.uint8)
    x16, y16 = np.ascontiguousarray(x[:, :16]), np.ascontiguousarray(y[:, :16])
    assert eng.fq_fp_select(P(c), P(x16), P(y16), P(out16), n, 3) == 0
    assert (out16 == np.where(c.reshape(-1, 1) == 1, x16, y16)).all()


def test_gfp25519_ops_through_the_engine(eng):
    """fq_fp25519_op: binary, unary and the batched inversion (whose kernel parks prefixes in the output buffer) sliced over devices."""
    from oracle import fourq_oracle as O
    rng = np.random.default_rng(8)
    n = 3001
    a = rng.integers(0, 256, (n, 32), np.uint8); b = rng.integers(0, 256, (n, 32), np.uint8)
    a[5] = 0; a[6] = np.frombuffer(O.P25519.to_bytes(32, "little"), np.uint8)
    for op, code, binary in (("mul", 0, True), ("sqr", 1, False), ("inv", 2, False), ("add", 3, True), ("sub", 4, True)):
        out = np.full_like(a, 0xEE)
        assert eng.fq_fp25519_op(code, P(a), P(b) if binary else None, P(out), n, 3) == 0, eng.fq_last_error()
        idx = list(range(0, n, 97)) + [5, 6, n - 1]
        for i in idx:
            assert bytes(out[i]) == O.row_f25519(op, bytes(a[i]), bytes(b[i]) if binary else None), (op, i)
    assert eng.fq_fp25519_op(5, P(a), None, P(out), n, 1) != 0          # neg / invsqrt do not exist in GFp25519
    assert eng.mock_violations() == 0


def test_sliced_pinned_allocation_is_page_locked_and_freed(eng, sim):
    """fq_host_alloc_sliced: the NUMA placement is best effort (nothing to observe on this box), but the array must be
    page-locked over its whole length, usable by a multi-GPU call without staging, and released by fq_host_free."""
    n = 300_001
    before = eng.mock_live_allocs(1, None)
    bufs = []
    for _ in range(3):
        p = ctypes.c_void_p()
        assert eng.fq_host_alloc_sliced(ctypes.byref(p), n, 32, 4) == 0, eng.fq_last_error()
        bufs.append((np.frombuffer((ctypes.c_uint8 * (n * 32)).from_address(p.value), np.uint8).reshape(n, 32), p))
    (a, ha), (b, hb), (o, ho) = bufs
    assert eng.mock_is_pinned(ha) and eng.mock_is_pinned(ctypes.c_void_p(ha.value + n * 32 - 1))
    rng = np.random.default_rng(12)
    a[:] = rng.integers(0, 256, (n, 32), np.uint8); b[:] = rng.integers(0, 256, (n, 32), np.uint8); o[:] = 0
    staged = eng.mock_memcpy_bytes()
    assert eng.fq_fp2_sub(ha, hb, ho, n, 4) == 0, eng.fq_last_error()
    assert eng.mock_memcpy_bytes() - staged == 3 * n * 32                 # H2D of a and b, D2H of out: no staging copies in between
    ref = np.empty_like(a)
    sim.sim_fp2_op(4, P(a), P(b), P(ref), ctypes.c_size_t(n))
    assert (o == ref).all()
    del a, b, o, bufs
    for h in (ha, hb, ho):
        assert eng.fq_host_free(h) == 0
    assert eng.mock_live_allocs(1, None) == before and eng.mock_violations() == 0
    assert eng.fq_host_alloc_sliced(None, 4, 32, 1) == _lib.FQ_ERR_ARG


def test_slices_follow_the_measured_speeds(eng):
    """CallWork::init: equal contiguous slices by default; with known speeds the slices are proportional (within +-25 % of the equal
    share, multiples of 128 rows) and still cover every row exactly once."""
    eng.fq_test_slices.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    for n, ndev, full in ((1 << 24, 8, 303104), (8 << 20, 8, 454656), (1000, 4, 454656), (3, 8, 128), (1 << 22, 2, 454656)):
        lo = (ctypes.c_size_t * ndev)(); hi = (ctypes.c_size_t * ndev)()
        for share in (None, [225.5] * (ndev // 2) + [209.0] * (ndev - ndev // 2), [1.0] + [100.0] * (ndev - 1), [0.0] * ndev):
            sh = (ctypes.c_double * ndev)(*share) if share else None
            assert eng.fq_test_slices(n, ndev, full, sh, lo, hi) == ndev
            assert lo[0] == 0 and hi[ndev - 1] == n and all(hi[i] == lo[i + 1] for i in range(ndev - 1)) and all(lo[i] <= hi[i] for i in range(ndev))
            per = (n + ndev - 1) // ndev
            lens = [hi[i] - lo[i] for i in range(ndev)]
            if share is None or not all(share) or n < ndev * 2 * full:
                assert lens == [max(0, min(per, n - i * per)) for i in range(ndev)]
            else:
                assert all(0.65 * n / ndev <= x <= 1.30 * n / ndev + 128 * ndev for x in lens), lens
                assert all(x % 128 == 0 for x in lens[:-1])
                if share[0] > share[-1]:
                    assert lens[0] > lens[-1]


def test_a_slow_gpu_gets_help_from_the_others(eng, sim):
    """Work stealing between slices (capi.cu CallWork): with one GPU much slower than the rest, the others finish their own slices
    and take rows from the back of the slow one's; the bytes are the same as with equal slices."""
    rng = np.random.default_rng(13)
    n = 8_000_003                                   # slices of 2 M rows: more than the four chunks a GPU keeps in flight
    a = rng.integers(0, 256, (n, 32), np.uint8); out = np.zeros_like(a)
    ref = np.empty_like(a)
    sim.sim_fp2_op(5, P(a), None, P(ref), ctypes.c_size_t(n))
    rows = (ctypes.c_size_t * 4)()
    assert eng.fq_fp2_neg(P(a), P(out), n, 4) == 0
    assert eng.fq_last_rows_per_device(rows, 4) == 0 and sum(rows) == n
    eng.mock_set_kernel_delay_us(2, 60000)                       # GPU 2: 60 ms extra per kernel
    try:
        out[:] = 0
        assert eng.fq_fp2_neg(P(a), P(out), n, 4) == 0, eng.fq_last_error()
    finally:
        eng.mock_set_kernel_delay_us(2, 0)
    assert (out == ref).all()
    assert eng.fq_last_rows_per_device(rows, 4) == 0 and sum(rows) == n
    per = (n + 3) // 4
    assert rows[2] < per and max(rows[0], rows[1], rows[3]) > per, list(rows)
    assert eng.mock_violations() == 0


def test_tiny_and_ragged_batches_on_every_gpu_count(eng, sim):
    """Fewer rows than GPUs, one row, sizes around the 128-row granularity: the empty slices get no job and every row is done once."""
    rng = np.random.default_rng(14)
    for n in (1, 2, 3, 4, 5, 127, 129, 1000):
        for ndev in (1, 2, 3, 4):
            a = rng.integers(0, 256, (n, 32), np.uint8); out = np.zeros_like(a); ref = np.zeros_like(a)
            assert eng.fq_fp2_neg(P(a), P(out), n, ndev) == 0, eng.fq_last_error()
            sim.sim_fp2_op(5, P(a), None, P(ref), ctypes.c_size_t(n))
            assert (out == ref).all(), (n, ndev)
            rows = (ctypes.c_size_t * ndev)()
            assert eng.fq_last_rows_per_device(rows, ndev) == 0 and sum(rows) == n


def test_argument_errors(eng):
    a = np.zeros((4, 32), np.uint8); out = np.zeros_like(a)
    assert eng.fq_fp2_sqr(P(a), P(out), 4, 0) == _lib.FQ_ERR_ARG
    assert eng.fq_fp2_sqr(P(a), P(out), 4, 5) == _lib.FQ_ERR_ARG          # 4 mock devices
    assert eng.fq_fp2_sqr(None, P(out), 4, 1) == _lib.FQ_ERR_ARG
    assert eng.fq_fp2_sqr(P(a), P(out), 0, 1) == 0
    d = ctypes.c_void_p()
    assert eng.fq_dev_alloc(0, ctypes.byref(d), 128) == 0
    assert eng.fq_fp2_sqr(P(a), d, 4, 1) == _lib.FQ_ERR_ARG               # a device pointer is not a host buffer
    assert b"device memory" in eng.fq_last_error()
    assert eng.fq_dev_free(0, d) == 0
    eng.mock_set_device_count(0)
    try:
        assert eng.fq_fp2_sqr(P(a), P(out), 4, 1) == _lib.FQ_ERR_NO_DEVICE
    finally:
        eng.mock_set_device_count(4)


@pytest.mark.parametrize("api,nth", [(b"cudaMemcpyAsync", 1), (b"cudaMemcpyAsync", 7), (b"cudaMalloc", 1), (b"cudaEventRecord", 5),
                                     (b"cudaEventSynchronize", 2), (b"cudaHostAlloc", 1)])
def test_injected_cuda_failures_return_an_error_and_do_not_wedge_the_engine(eng, sim, api, nth):
    rng = np.random.default_rng(8)
    n = 600_000
    a = rng.integers(0, 256, (n, 32), np.uint8); out = np.zeros_like(a)
    assert eng.fq_trim() == 0                       # buffers gone: the next call has to allocate (and stage: pageable arrays)
    eng.mock_fail_nth(api, nth)
    rc = eng.fq_fp2_sqr(P(a), P(out), n, 2)
    eng.mock_fail_nth(api, 0)
    assert rc == _lib.FQ_ERR_CUDA and eng.fq_last_error() != b""
    ref = np.empty_like(a)
    sim.sim_fp2_op(1, P(a), None, P(ref), ctypes.c_size_t(n))
    assert eng.fq_fp2_sqr(P(a), P(out), n, 2) == 0, eng.fq_last_error()
    assert (out == ref).all()


def test_concurrent_callers(eng, sim):
    """Several host threads issue multi-device calls at once: every GPU serves them in submission order, results stay right."""
    eng.mock_set_jitter_us(100)
    rng = np.random.default_rng(9)
    n = 200_000
    errs = []

    def work(seed, ndev):
        r = np.random.default_rng(seed)
        a = r.integers(0, 256, (n, 32), np.uint8); b = r.integers(0, 256, (n, 32), np.uint8)
        ref = np.empty_like(a); out = np.zeros_like(a)
        sim.sim_fp2_op(3, P(a), P(b), P(ref), ctypes.c_size_t(n))
        for _ in range(3):
            out[:] = 0
            if eng.fq_fp2_add(P(a), P(b), P(out), n, ndev) != 0 or not (out == ref).all():
                errs.append((seed, ndev))
    ts = [threading.Thread(target=work, args=(100 + i, 1 + i % 4)) for i in range(6)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs and eng.mock_violations() == 0
    del rng


def test_trim_wipes_and_frees_everything(eng):
    k = np.random.default_rng(3).integers(0, 256, (300, 32), np.uint8); out = np.empty_like(k)
    assert eng.fq_mul_base_comb(P(k), P(out), 300, 2) == 0
    nb = ctypes.c_long()
    assert eng.mock_live_allocs(1, ctypes.byref(nb)) > 0                   # pinned staging exists (pageable arrays)
    eng.mock_expect_zero_on_free(1)
    try:
        assert eng.fq_trim() == 0
    finally:
        eng.mock_expect_zero_on_free(0)
    assert eng.mock_violations() == 0                                      # every freed staging / scratch buffer had been zeroed
    assert eng.mock_live_allocs(1, ctypes.byref(nb)) == 0
    assert eng.mock_live_allocs(2, ctypes.byref(nb)) <= 4                  # only the per-device comb tables stay


def test_wipe_after_call_leaves_no_secret_in_the_engine_buffers():
    """FQ_WIPE_AFTER_CALL=1 (read once per process, hence a child process): after a keygen call every staging and scratch buffer the
    engine keeps -- device and pinned -- is all zeros; the caller's own arrays are untouched."""
    so = os.path.join(SIM_DIR, "libfq_mockengine.so")
    code = """
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
from fourq_b200 import _lib
L = _lib.bind(ctypes.CDLL(%r))
L.mock_nonzero_allocs.restype = ctypes.c_long; L.mock_nonzero_allocs.argtypes = [ctypes.c_int, ctypes.c_long]
L.mock_set_device_count(2)
k = np.random.default_rng(1).integers(1, 256, (700, 32), np.uint8); out = np.zeros_like(k); st = np.zeros(700, np.uint8)
assert L.fq_dh_base_comb(_lib.ptr(k), _lib.ptr(out), _lib.ptr(st), 700, 2) == 0
assert out.any() and k.any()
print(L.mock_nonzero_allocs(2, 1024), L.mock_nonzero_allocs(1, 1024))
""" % (os.path.join(HERE, ".."), so)
    res = {}
    for wipe in ("0", "1"):
        r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, FQ_WIPE_AFTER_CALL=wipe))
        assert r.returncode == 0, r.stderr[-2000:]
        res[wipe] = [int(x) for x in r.stdout.split()]
    assert res["0"][0] > 0 and res["0"][1] > 0           # without the switch the scalars and results stay in the staging buffers
    assert res["1"] == [0, 0]                            # with it nothing of at least 1 KiB that the engine allocated holds data


@pytest.mark.parametrize("san", ["tsan", "asan"])
def test_engine_stress_under_sanitizers(san):
    """ThreadSanitizer / AddressSanitizer + UBSan over the engine's threads, staging and scratch sizing (engine_stress.cpp)."""
    exe = os.path.join(SIM_DIR, "engine_stress_" + san)
    if _stale(exe):
        subprocess.check_call(["sh", os.path.join(SIM_DIR, "build.sh"), san])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 second_deadlock_stack=1", ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "Sanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-4000:]
    assert "engine_stress ok" in r.stdout
