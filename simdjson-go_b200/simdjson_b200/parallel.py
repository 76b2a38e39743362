"""NDJSON sharding across ranks.

Every raw 0x0A in a valid NDJSON stream is a record boundary (a newline inside a string is a
stage-1 error, find_quote_mask_and_bits_amd64.s:69-80), so a shard boundary is simply the
first '\\n' at or after byte k*N/G -- the rule ParseNDStream uses for its 10 MiB chunks
(simdjson_amd64.go:165-174).  Shards parse independently with fresh carries; the only
exchange is one all-gather of three integers per rank (shard bytes, tape words, string
bytes), whose exclusive prefix tells each rank where its tape / strings / message would sit
in one concatenated ParsedJson.  `rebase_shard_tape` applies those offsets.
"""
import numpy as np

TAG_SHIFT = np.uint64(56)
VAL_MASK = np.uint64((1 << 56) - 1)
STRINGBUFBIT = np.uint64(1 << 55)


def split_at_newlines(buf, world):
    """[(start, stop)] per rank: contiguous, newline-aligned, covering buf."""
    n = len(buf)
    cuts = [0]
    for r in range(1, world):
        k = max(cuts[-1], r * n // world)
        j = buf.find(b"\n", k) if isinstance(buf, (bytes, bytearray)) else _find_nl(buf, k)
        cuts.append(n if j < 0 else j + 1)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def _find_nl(arr, k):
    hit = np.nonzero(np.asarray(arr[k:]) == 0x0A)[0]
    return -1 if hit.size == 0 else int(hit[0]) + k


def exchange_totals(local, group=None, device=None):
    """all_gather of (shard_bytes, tape_words, string_bytes); returns (exclusive prefix of this
    rank, list of all ranks' totals).  Works on gloo (CPU tensors) and nccl (device tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(x) for x in local], dtype=torch.int64, device=device)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    totals = [tuple(int(x) for x in t.tolist()) for t in allv]
    base = tuple(sum(t[i] for t in totals[:rank]) for i in range(3))
    return base, totals


def rebase_shard_tape(tape, tape_base, strings_base, msg_base):
    """Shift every tape-relative, string-buffer-relative and message-relative payload of a
    shard-local tape (Appendix B of SURVEY.md) so the shard can be concatenated after
    `tape_base` tape words / `strings_base` string bytes / `msg_base` message bytes."""
    t = np.asarray(tape, dtype=np.uint64)
    tags = (t >> TAG_SHIFT).astype(np.uint8)
    two_word = np.isin(tags, np.frombuffer(b'"lud', dtype=np.uint8))
    # the raw second word of a string / number can look like anything: inside a run of
    # consecutive candidates the real first words sit at even positions
    idx = np.arange(t.size)
    run_start = two_word & ~np.concatenate(([False], two_word[:-1]))
    start_idx = np.maximum.accumulate(np.where(run_start, idx, 0))
    first = two_word & (((idx - start_idx) & 1) == 0)
    second = np.concatenate(([False], first[:-1]))
    out = t.copy()
    links = ~second & np.isin(tags, np.frombuffer(b"r{}[]", dtype=np.uint8))
    out[links] = t[links] + np.uint64(tape_base)
    strs = first & (tags == ord('"'))
    in_buf = strs & ((t & STRINGBUFBIT) != 0)
    out[in_buf] = t[in_buf] + np.uint64(strings_base)
    in_msg = strs & ~in_buf
    out[in_msg] = t[in_msg] + np.uint64(msg_base)
    return out


def trimmed_window(buf, a, b):
    """[a, b) shrunk by the ASCII blanks at both ends (shards are cut behind a newline; the parse wants the trimmed
    window, parse_json_amd64.go:55).  Shard boundaries are ASCII newlines, so the ASCII rule is exact here."""
    ws = b" \t\n\v\f\r"
    while a < b and buf[a] in ws:
        a += 1
    while b > a and buf[b - 1] in ws:
        b -= 1
    return a, b


class ShardedParse:
    """ParseND over the ranks of a process group: the two halves of sj_parse_nd_sharded_* around ONE all-gather of the
    shard totals.  Each rank ends up with its slice of the single ParsedJson the reference returns
    (simdjson_amd64.go:82-93), already rebased: rank r's slice starts at tape word `tape_base` / Strings.B byte
    `strings_base` of the whole.  Buffers are device pointers (integers); this class owns none."""

    def __init__(self, ctx, group=None, device=None):
        self.ctx, self.group, self.device = ctx, group, device
        self.bases_ptr = None  # device pointer to the bases once connect_exchange succeeded

    def count(self, d_msg, n, copy_strings=True, d_totals=None):
        """d_totals: optional device pointer that receives the same four integers on the context's stream"""
        import ctypes as C
        from . import _lib
        t = _lib.ShardTotals()
        flags = _lib.FLAG_NDJSON | (_lib.FLAG_COPY_STRINGS if copy_strings else 0)
        rc = self.ctx.L.sj_parse_nd_sharded_count(self.ctx.h, d_msg, n, flags, C.byref(t), d_totals)
        return rc, (int(t.msg_bytes), int(t.tape_words), int(t.string_bytes), int(t.records))

    def emit(self, msg_base, tape_base, strings_base, d_tape, tape_cap, d_strings, strings_cap, d_bases=None):
        """d_bases: optional device pointer to { msg_base, tape_base, strings_base } (then the three scalars are ignored)"""
        return self.ctx.L.sj_parse_nd_sharded_emit(self.ctx.h, msg_base, tape_base, strings_base, d_bases, d_tape, tape_cap,
                                                   d_strings, strings_cap)

    def exchange(self, totals):
        """all-gather of this rank's (msg_bytes, tape_words, string_bytes): (exclusive prefix, every rank's totals)"""
        return exchange_totals(totals[:3], self.group, self.device)

    # ---- the exchange as a kernel over peer memory (exchange.cuh): set up once, then count() is a collective call whose
    # result -- the bases -- is already in device memory (`bases_ptr`) when it returns ----
    def connect_exchange(self, rank, world, gap_bytes=1):
        """One process per GPU: all-gather of the CUDA IPC handles of the ranks' exchange buffers over `group`, then
        every rank maps its peers' buffers.  Returns the library's return code (0 = ok); anything else means this box
        cannot share device memory between processes and the caller keeps exchanging through its collective library."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        L = self.ctx.L
        mine = (C.c_uint8 * _lib.EXCHANGE_HANDLE_BYTES)()
        rc = L.sj_exchange_create(self.ctx.h, rank, world, gap_bytes, mine)
        t = torch.tensor([rc] + list(bytes(mine)), dtype=torch.int64, device=self.device)
        allv = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allv, t, group=self.group)
        rows = [[int(v) for v in a.tolist()] for a in allv]
        if any(r[0] != 0 for r in rows):
            return next(r[0] for r in rows if r[0] != 0)
        blob = bytes(b for r in rows for b in r[1:])
        rc = L.sj_exchange_connect(self.ctx.h, blob)
        ok = torch.tensor([rc], dtype=torch.int64, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MAX, group=self.group)  # all ranks or none
        worst = int(ok.item())
        self.bases_ptr = L.sj_exchange_bases(self.ctx.h) if worst == 0 else None
        return rc if rc else worst

    def exchange_result(self):
        """{ msg_base, tape_base, strings_base, records_base, whole x4, status, epoch } of the last count()"""
        import ctypes as C
        out = (C.c_uint64 * 10)()
        rc = self.ctx.L.sj_exchange_result(self.ctx.h, out)
        return rc, [int(v) for v in out]


def reduce_counts(local, group=None, device=None):
    """countWhere / countObjects over a sharded stream (consume.cuh): records are independent, so the
    answer for the whole stream is the sum of the per-shard (roots, matches) -- one all_reduce of
    two integers.  Works on gloo (CPU tensors) and nccl (device tensors)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(x) for x in local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) for x in t.tolist())
