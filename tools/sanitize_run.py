"""Small end-to-end workload for `compute-sanitizer` (memcheck / racecheck / synccheck / initcheck): every kernel of
the path on inputs small enough for the tool's 10-100x slowdown, each result still checked against the oracle.
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py [quick]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simdjson-go_b200"))
import numpy as np

import simdjson_b200 as sj
from oracle.pyoracle import Oracle
from tests.util import load_fixture, tricky_ndjson

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = sj.Context(0)          # streaming stage 2 (default with copy_strings)
ctx_legacy = sj.Context(0)   # per-structural stage 2
ctx_legacy.set_stage2_impl(1)
o = Oracle("native")
n = 0


def same(msg, ndjson=False, copy=True):
    global n
    rc_o, tape_o, str_o, win_o = o.parse(msg, ndjson=ndjson, copy_strings=copy)
    for cx in (ctx, ctx_legacy):
        rc_g, tape_g, str_g, win_g = cx.parse(msg, ndjson=ndjson, copy_strings=copy)
        assert rc_g == rc_o and win_g == win_o, (rc_g, rc_o)
        if rc_o == 0:
            assert np.array_equal(tape_g, tape_o) and str_g == str_o
    ok_g, d_g = ctx.find_structural_indices(msg, ndjson)
    ok_o, d_o = o.find_structural_indices(msg, ndjson)
    assert ok_g == ok_o and (not ok_o or np.array_equal(d_g, d_o))
    n += 1


docs = ["twitter", "twitterescaped", "numbers", "apache_builds", "github_events"] if not quick else ["github_events"]
for name in docs:
    d = load_fixture(name).strip()
    same(d)
    same(d, copy=False)
pk = load_fixture("parking-citations").strip()
same(pk[:120000 if quick else len(pk)].rsplit(b"\n", 1)[0], ndjson=True)
nd, _ = tricky_ndjson()
same(nd, ndjson=True)
assert ctx.parse_count_where(nd, b"Make", b"HOND")[0] == 0
# carries across slab / tile edges, long strings (warp-cooperative paths), dense brackets, number-heavy
g = (__import__("ctypes").c_uint32 * 4)()
ctx.L.sj_test_geometry(g)
slab, tile = int(g[2]), int(g[3])
for edge in (slab, tile):
    for run in (1, 33, 64):
        same(b'["' + b"x" * (edge - 2 - run // 2) + b"\\" * run + b'"q\\\\", "tail",true , 12]')
same(b'{"k":"' + b"s" * (tile + 77) + b'","t":[1,2,{"u":null}]}')
same(b'["' + b"\\u30c6\\u30b9\\ud83d\\ude00\\n" * 300 + b'","' + b"abc" * 200 + b'"]')
same(b"[" + b"[1," * 3000 + b"1" + b"]" * 3000 + b"]")
same(b"[" + b",".join(b"%d.%de%d" % (i, i * 7919 % 100000, i % 30 - 15) for i in range(4000)) + b"]")
same(b'{"a":tru}')
same(b'["abc\\q"]')
# the sharded ParseND with the exchange kernel over peer memory (exchange.cuh): two host threads = two ranks on this GPU
import ctypes as C
import threading

import torch

from simdjson_b200.parallel import ShardedParse, split_at_newlines, trimmed_window

stream = pk[:60000].rsplit(b"\n", 1)[0] + b'\n{"esc":"a\\u00e9\\n","n":[1,2.5,-3],"t":true}'
rc_o, tape_o, str_o, (off_o, len_o) = o.parse(stream, ndjson=True, copy_strings=True)
wins = [trimmed_window(stream, a, b) for a, b in split_at_newlines(stream, 2)]
cx = [sj.Context(0), sj.Context(0)]
for r, c in enumerate(cx):
    assert c.L.sj_exchange_create(c.h, r, 2, 1, None) == 0
    assert c.L.sj_exchange_set_timeout_ms(c.h, 120000) == 0
locs = (C.c_void_p * 2)(*[c.L.sj_exchange_local(c.h) for c in cx])
res = [None, None]


def rank(r):
    c = cx[r]
    a, b = wins[r]
    assert c.L.sj_exchange_connect_ptrs(c.h, locs) == 0
    assert c.L.sj_exchange_set_gap(c.h, wins[1][0] - b if r == 0 else 0) == 0
    dev = torch.device("cuda:0")
    d_msg = torch.full((b - a + 256,), 0x20, dtype=torch.uint8, device=dev)
    d_msg[: b - a] = torch.frombuffer(bytearray(stream[a:b]), dtype=torch.uint8).to(dev)
    torch.cuda.synchronize()
    sp = ShardedParse(c)
    rc, tot = sp.count(d_msg.data_ptr(), b - a, True)
    assert rc == 0, rc
    d_tape = torch.empty(tot[1] + 8, dtype=torch.int64, device=dev)
    d_str = torch.empty(tot[2] + 64, dtype=torch.uint8, device=dev)
    assert sp.emit(0, 0, 0, d_tape.data_ptr(), d_tape.numel(), d_str.data_ptr(), d_str.numel(), c.L.sj_exchange_bases(c.h)) == 0
    res[r] = (d_tape[: tot[1]].cpu().numpy().view(np.uint64), d_str[: tot[2]].cpu().numpy().tobytes())


ts = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
for t in ts:
    t.start()
for t in ts:
    t.join()
assert res[0] is not None and res[1] is not None
assert np.array_equal(np.concatenate([res[0][0], res[1][0]]), tape_o) and res[0][1] + res[1][1] == str_o
n += 1
print("sanitize_run ok: %d documents, %d kernel launches" % (n, ctx.launches()))
