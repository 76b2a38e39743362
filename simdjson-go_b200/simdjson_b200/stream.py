"""ParseNDStream (simdjson_amd64.go:116-215) over the GPU path.

The reference reads 10 MiB, extends the chunk to the next '\\n' (:157-174), parses chunks
concurrently (at most (GOMAXPROCS+1)/2 in flight, :132) and delivers the results in input
order (:134-152).  Here the chunk is sized to fill a GPU (default 256 MiB), `inflight` host
threads each own one context (= one CUDA stream + scratch), so the H2D copy, the kernels and
the D2H copy of consecutive chunks overlap; results are yielded in input order, each one an
independent ParsedJson exactly as in the reference.
"""
from collections import deque
from concurrent.futures import ThreadPoolExecutor

from . import ERR_STAGE1, ERR_STAGE2, Context, ParsedJson, ParseError


def _chunks(reader, chunk_bytes):
    """newline-aligned chunks of about chunk_bytes from a binary file-like object"""
    carry = b""
    while True:
        blk = reader.read(chunk_bytes)
        if not blk:
            if carry.strip():
                yield carry
            return
        buf = carry + blk
        cut = buf.rfind(b"\n")
        if cut < 0:  # no record boundary yet: keep reading (one record larger than a chunk)
            carry = buf
            continue
        yield buf[:cut + 1]
        carry = buf[cut + 1:]


def ParseNDStream(reader, chunk_bytes=256 << 20, inflight=3, copy_strings=True, device=0):
    """Generator of ParsedJson, one per newline-aligned chunk, in input order.
    Raises ParseError ("parsing input: ...") at the position of the first bad chunk, like the
    reference's Stream{Error} (simdjson_amd64.go:196)."""
    ctxs = [Context(device) for _ in range(max(1, inflight))]
    free = deque(ctxs)

    def work(ctx, data):
        rc, tape, strings, (off, ln) = ctx.parse(data, ndjson=True, copy_strings=copy_strings)
        return rc, ParsedJson(data[off:off + ln], tape, strings) if rc == 0 else None

    pending = deque()
    try:
        with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
            for data in _chunks(reader, chunk_bytes):
                if not free:  # oldest result first: keeps delivery in input order
                    ctx, fut = pending.popleft()
                    rc, pj = fut.result()
                    free.append(ctx)
                    if rc in (ERR_STAGE1, ERR_STAGE2):
                        raise ParseError("parsing input: " + ("Failed to find all structural indices for stage 1"
                                                              if rc == ERR_STAGE1 else "Bad parsing while executing stage 2"))
                    yield pj
                ctx = free.popleft()
                pending.append((ctx, ex.submit(work, ctx, data)))
            while pending:
                ctx, fut = pending.popleft()
                rc, pj = fut.result()
                free.append(ctx)
                if rc in (ERR_STAGE1, ERR_STAGE2):
                    raise ParseError("parsing input: " + ("Failed to find all structural indices for stage 1"
                                                          if rc == ERR_STAGE1 else "Bad parsing while executing stage 2"))
                yield pj
    finally:
        for c in ctxs:
            c.close()
