"""What the reference publishes: per-file Parse() of the real fixtures (benchmarks_test.go:23-78, README "Performance"
table, `copy` rows).  Host bytes in -> sj_parse -> host tape + strings out, one call at a time on one context (latency,
not pipelined throughput), pinned host buffers, median of N calls.  Beside it the CPU port (one thread) on this host and
the reference's own README numbers (unstated hardware).  Prints a markdown table.
usage: fixture_bench.py [calls per file (default 200)]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simdjson-go_b200"))
import numpy as np

import simdjson_b200 as sj
from oracle.pyoracle import FLAG_COPY_STRINGS, Oracle
from simdjson_b200 import _lib
from tests.util import load_fixture

README_COPY_MBS = {"apache_builds": 890.21, "canada": 167.77, "citm_catalog": 1270.02, "github_events": 879.67, "gsoc-2018": 2642.88,
                   "instruments": 731.15, "marine_ik": 206.90, "mesh": 172.03, "mesh.pretty": 329.57, "numbers": 165.04, "random": 508.25,
                   "twitter": 1072.59, "twitterescaped": 578.46, "update-center": 752.31}
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = sj.Context(0)
L = ctx.L
o = Oracle("best")


def pinned(nbytes):
    p = L.sj_host_alloc(nbytes)
    assert p
    return p


print("| file | bytes | sj_parse us (median) | sj_parse MB/s | launches / call | CPU port 1 thread MB/s (%s) | reference README MB/s (copy) | sj_parse / README |" % o.isa)
print("|---|---|---|---|---|---|---|---|")
for name in sorted(README_COPY_MBS):
    doc = load_fixture(name).strip()
    n = len(doc)
    h_in = pinned(n + 64)
    C.memmove(h_in, doc, n)
    tcap, scap = C.c_size_t(0), C.c_size_t(0)
    L.sj_bounds(n, C.byref(tcap), C.byref(scap))
    tcap_w = min(tcap.value, n + 1024)  # tape words <= structurals * 2, far below the bound for real documents
    h_tape, h_str = pinned(tcap_w * 8), pinned(scap.value)
    tl, sl, mo, ml = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    flags = _lib.FLAG_COPY_STRINGS

    def one():
        r = L.sj_parse(ctx.h, h_in, n, flags, h_tape, tcap_w, C.byref(tl), h_str, scap.value, C.byref(sl), C.byref(mo), C.byref(ml))
        assert r == 0, (name, r)

    for _ in range(10):
        one()
    l0 = ctx.launches()
    ts = []
    for _ in range(calls):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    per_call = (ctx.launches() - l0) / calls
    med = float(np.median(ts))
    # the parse is the reference's: same tape as the oracle
    rc, tape_o, str_o, _ = o.parse(doc)
    got = np.ctypeslib.as_array(C.cast(h_tape, C.POINTER(C.c_uint64)), shape=(tl.value,))
    assert rc == 0 and np.array_equal(got, tape_o) and C.string_at(h_str, sl.value) == str_o, name
    # CPU port, one thread, reused buffers
    arr = np.frombuffer(doc, dtype=np.uint8)
    tape = np.empty(2 * n + 64, dtype=np.uint64)
    strs = np.empty(n + 64, dtype=np.uint8)
    reps = max(3, int(2e7 // n))
    t0 = time.perf_counter()
    for _ in range(reps):
        o.lib.sjo_parse(arr.ctypes.data, n, FLAG_COPY_STRINGS, tape.ctypes.data, tape.size, C.byref(tl), strs.ctypes.data, strs.size,
                        C.byref(sl), C.byref(mo), C.byref(ml))
    cpu = n * reps / (time.perf_counter() - t0) / 1e6
    mbs = n / med / 1e6
    print("| %s | %d | %.0f | %.0f | %.0f | %.0f | %.0f | %.2f |" % (name, n, med * 1e6, mbs, per_call, cpu, README_COPY_MBS[name], mbs / README_COPY_MBS[name]))
    for p_ in (h_in, h_tape, h_str):
        L.sj_host_free(p_)
