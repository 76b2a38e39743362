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


def ParseNDStreamNative(reader, chunk_bytes=256 << 20, inflight=3, copy_strings=True, device=0, read_bytes=16 << 20):
    """The same stream driven through the library's own pipeline (sj_stream_*, sj_stream.inl): the
    chunking at record boundaries, the pinned staging of (pageable) input, the worker threads and the
    ordered delivery all live behind the C ABI, which is what a cgo caller binds.  One Python thread
    drives it: write until the pipeline takes nothing more, then take the oldest result."""
    import ctypes as C

    import numpy as np

    from . import _lib
    L = _lib.load()
    h = C.c_void_p()
    rc = L.sj_stream_create(device, inflight, chunk_bytes, _lib.FLAG_COPY_STRINGS if copy_strings else 0, C.byref(h))
    if rc != _lib.OK:
        raise _lib.SjError(rc)

    def take():
        """next result as a ParsedJson (copied out of the pinned slot), None when nothing is in flight / at the end"""
        res = _lib.StreamResult()
        rc = L.sj_stream_next(h, C.byref(res))
        if rc in (_lib.STREAM_EMPTY, _lib.STREAM_END):
            return rc, None
        if rc in (ERR_STAGE1, ERR_STAGE2):
            raise ParseError("parsing input: " + ("Failed to find all structural indices for stage 1"
                                                  if rc == ERR_STAGE1 else "Bad parsing while executing stage 2"))
        if rc != _lib.OK:
            raise _lib.SjError(rc)
        msg = C.string_at(res.message, res.message_len)
        tape = np.frombuffer(C.string_at(res.tape, res.tape_len * 8), dtype=np.uint64)
        strings = C.string_at(res.strings, res.strings_len) if res.strings_len else b""
        L.sj_stream_release(h, C.byref(res))
        return rc, ParsedJson(msg, tape, strings)

    try:
        taken = C.c_size_t(0)
        while True:
            blk = reader.read(read_bytes)
            if not blk:
                break
            view = np.frombuffer(blk, dtype=np.uint8)
            off = 0
            while off < view.size:
                rc = L.sj_stream_write(h, view[off:].ctypes.data, view.size - off, C.byref(taken))
                if rc in (ERR_STAGE1, ERR_STAGE2):
                    take()  # raises the stream's error
                if rc != _lib.OK:
                    raise _lib.SjError(rc)
                off += taken.value
                if taken.value == 0:  # every slot is busy: deliver the oldest chunk
                    rc, pj = take()
                    if pj is None:
                        raise RuntimeError("sj_stream: nothing taken and nothing in flight")
                    yield pj
        while True:
            rc = L.sj_stream_close_input(h)
            if rc != _lib.STREAM_BUSY:
                break
            rc, pj = take()
            if pj is not None:
                yield pj
        if rc != _lib.OK:
            raise _lib.SjError(rc)
        while True:
            rc, pj = take()
            if pj is None:
                break
            yield pj
    finally:
        L.sj_stream_destroy(h)
