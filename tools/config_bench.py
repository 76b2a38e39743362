"""Device-resident timings of the BASELINE.json configs that are not the bench.py line:
single documents built by replicating a fixture inside one array (`[doc,doc,...]`), parsed with
sj_parse_device (stage 1 + stage 2) and with the stage-1 kernel alone.  Prints a markdown table.
usage: config_bench.py [MiB per document (default 256)] [comma-separated input names (default all)]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simdjson-go_b200"))
import numpy as np
import torch

import simdjson_b200 as sj
from simdjson_b200 import _lib
from tests.util import load_fixture

target = int(float(sys.argv[1]) * (1 << 20)) if len(sys.argv) > 1 else 256 << 20
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
ctx = sj.Context(0)
L = ctx.L
print("| input (replicated to ~%d MiB) | bytes | structurals | tape words | string bytes | stage1 ms | stage1 input GB/s | stage1 alg. GB/s | parse ms | parse GB/s |" % (target >> 20))
print("|---|---|---|---|---|---|---|---|---|---|")
for name, nd in (("twitter", False), ("twitterescaped", False), ("canada", False), ("gsoc-2018", False), ("citm_catalog", False),
                 ("parking-citations", True)):
    if only and name not in only:
        continue
    doc = load_fixture(name).strip()
    k = max(1, target // (len(doc) + 1))
    msg = (b"\n".join([doc] * k)) if nd else (b"[" + b",".join([doc] * k) + b"]")
    n = len(msg)
    d_msg = torch.empty(n + 65536, dtype=torch.uint8, device=dev)
    d_msg[:n] = torch.frombuffer(bytearray(msg), dtype=torch.uint8).to(dev)
    d_msg[n:] = 0x20
    flags = (_lib.FLAG_NDJSON if nd else 0) | _lib.FLAG_COPY_STRINGS
    tcap = 2 * n // 3 + (1 << 20) if name != "parking-citations" else 2 * n
    d_tape = torch.empty(tcap, dtype=torch.int64, device=dev)
    d_str = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    tl, sl = C.c_size_t(0), C.c_size_t(0)

    def parse():
        r = L.sj_parse_device(ctx.h, d_msg.data_ptr(), n, flags, d_tape.data_ptr(), tcap, C.byref(tl), d_str.data_ptr(),
                              d_str.numel(), C.byref(sl))
        assert r == 0, (name, r)

    for _ in range(3):
        parse()
    ms = C.c_float(0)
    reps = 5
    L.sj_event_record(ctx.h, 0)
    for _ in range(reps):
        parse()
    L.sj_event_record(ctx.h, 1)
    L.sj_event_elapsed_ms(ctx.h, C.byref(ms))
    t_parse = ms.value / reps
    info = sj.Stage1Info()
    cap = n // 3 + 1024
    d_idx = torch.empty(cap, dtype=torch.int32, device=dev)
    assert L.sj_stage1_device(ctx.h, d_msg.data_ptr(), n, int(nd), 0, d_idx.data_ptr(), cap, C.byref(info)) == 0
    for _ in range(3):
        L.sj_stage1_launch(ctx.h, d_msg.data_ptr(), n, int(nd), 0, d_idx.data_ptr(), cap)
    L.sj_ctx_sync(ctx.h)
    L.sj_event_record(ctx.h, 0)
    for _ in range(reps):
        L.sj_stage1_launch(ctx.h, d_msg.data_ptr(), n, int(nd), 0, d_idx.data_ptr(), cap)
    L.sj_event_record(ctx.h, 1)
    L.sj_event_elapsed_ms(ctx.h, C.byref(ms))
    t_s1 = ms.value / reps
    print("| %s | %d | %d | %d | %d | %.3f | %.0f | %.0f | %.3f | %.1f |" % (
        name, n, info.n_idx, tl.value, sl.value, t_s1, n / t_s1 / 1e6, (n + 4 * info.n_idx) / t_s1 / 1e6, t_parse,
        n / t_parse / 1e6))
    del d_msg, d_tape, d_str, d_idx
    torch.cuda.empty_cache()
