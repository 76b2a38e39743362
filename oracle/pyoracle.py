"""ctypes binding for the CPU oracle (test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))

FLAG_NDJSON = 1
FLAG_COPY_STRINGS = 2

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
_LIBS = ("libsjoracle.so", "libsjoracle_native.so", "libsjoracle_avx512.so")


def build(force=False):
    """Compile the oracle with gcc (idempotent)."""
    need = force or not all(os.path.exists(os.path.join(_DIR, f)) for f in _LIBS)
    if not need:
        src_m = max(os.path.getmtime(os.path.join(_DIR, f)) for f in ("sjoracle.c", "sjoracle.h"))
        need = any(os.path.getmtime(os.path.join(_DIR, f)) < src_m for f in _LIBS)
    if need:
        subprocess.check_call(["make", "-C", _DIR, "-s", "-B"])


def host_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    fl = set(ln.split(":", 1)[1].split())
                    return {"avx512f", "avx512bw", "avx512vl", "pclmulqdq"} <= fl
    except OSError:
        pass
    return False


def _buf(b):
    b = bytes(b)
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


class Oracle:
    def __init__(self, variant="scalar"):
        build()
        if variant == "best":  # the widest mask routines this host can run (the reference picks AVX-512 the same way)
            variant = "avx512" if host_has_avx512() else "native"
        name = {"scalar": "libsjoracle.so", "native": "libsjoracle_native.so", "avx512": "libsjoracle_avx512.so"}[variant]
        L = self.lib = C.CDLL(os.path.join(_DIR, name))
        L.sjo_isa.restype = C.c_char_p
        self.isa = L.sjo_isa().decode()
        L.sjo_find_odd_backslash_sequences.restype = C.c_uint64
        L.sjo_find_odd_backslash_sequences.argtypes = [C.c_void_p, u64p]
        L.sjo_find_quote_mask_and_bits.restype = C.c_uint64
        L.sjo_find_quote_mask_and_bits.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p]
        L.sjo_find_whitespace_and_structurals.restype = None
        L.sjo_find_whitespace_and_structurals.argtypes = [C.c_void_p, u64p, u64p]
        L.sjo_finalize_structurals.restype = C.c_uint64
        L.sjo_finalize_structurals.argtypes = [C.c_uint64] * 4 + [u64p]
        L.sjo_find_newline_delimiters.restype = C.c_uint64
        L.sjo_find_newline_delimiters.argtypes = [C.c_void_p, C.c_uint64]
        L.sjo_flatten_bits_incremental.restype = None
        L.sjo_flatten_bits_incremental.argtypes = [u32p, C.POINTER(C.c_int), C.c_uint64, u64p, u64p]
        L.sjo_find_structural_bits.restype = C.c_uint64
        L.sjo_find_structural_bits.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_uint64, u64p]
        L.sjo_find_structural_bits_in_slice.restype = C.c_uint64
        L.sjo_find_structural_bits_in_slice.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p, u64p, u32p,
                                                        C.POINTER(C.c_int), u64p, u64p, C.c_uint64]
        L.sjo_find_structural_indices.restype = C.c_int
        L.sjo_find_structural_indices.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                                  C.POINTER(C.c_size_t)]
        L.sjo_parse_string_validate_only.restype = C.c_int
        L.sjo_parse_string_validate_only.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, u64p, u64p]
        L.sjo_parse_string.restype = C.c_int
        L.sjo_parse_string.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, u64p]
        L.sjo_parse_number.restype = C.c_uint64
        L.sjo_parse_number.argtypes = [C.c_void_p, C.c_size_t, u64p]
        for n in ("true", "false", "null"):
            f = getattr(L, "sjo_is_valid_%s_atom" % n)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t]
        L.sjo_trim_space.restype = None
        L.sjo_trim_space.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.sjo_parse.restype = C.c_int
        L.sjo_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                C.POINTER(C.c_size_t)]
        L.sjo_count_where.restype = C.c_uint64
        L.sjo_count_where.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                      C.c_size_t, u64p]
        L.sjo_stage1_count.restype = C.c_size_t
        L.sjo_stage1_count.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_int)]

    # ---- block-level mirrors of the reference's Go stubs ---------------------
    def find_odd_backslash_sequences(self, in64, prev):
        p = C.c_uint64(prev)
        r = self.lib.sjo_find_odd_backslash_sequences(_buf(in64), C.byref(p))
        return r, p.value

    def find_quote_mask_and_bits(self, in64, odd_ends, prev_inside, error_mask=0):
        pi, qb, em = C.c_uint64(prev_inside), C.c_uint64(0), C.c_uint64(error_mask)
        qm = self.lib.sjo_find_quote_mask_and_bits(_buf(in64), odd_ends, C.byref(pi), C.byref(qb), C.byref(em))
        return qm, qb.value, pi.value, em.value

    def find_whitespace_and_structurals(self, in64):
        ws, st = C.c_uint64(0), C.c_uint64(0)
        self.lib.sjo_find_whitespace_and_structurals(_buf(in64), C.byref(ws), C.byref(st))
        return ws.value, st.value

    def finalize_structurals(self, structurals, whitespace, quote_mask, quote_bits, prev_pseudo):
        pp = C.c_uint64(prev_pseudo)
        r = self.lib.sjo_finalize_structurals(structurals, whitespace, quote_mask, quote_bits, C.byref(pp))
        return r, pp.value

    def find_newline_delimiters(self, in64, quote_mask):
        return self.lib.sjo_find_newline_delimiters(_buf(in64), quote_mask)

    def flatten_bits(self, masks, carried=0, position=(1 << 64) - 1):
        base = (C.c_uint32 * (64 * len(masks) + 8))()
        idx, car, pos = C.c_int(0), C.c_uint64(carried), C.c_uint64(position)
        for m in masks:
            self.lib.sjo_flatten_bits_incremental(base, C.byref(idx), m, C.byref(car), C.byref(pos))
        return list(base[:idx.value]), car.value, pos.value

    def find_structural_bits(self, in64, prev_odd, prev_inside, error_mask, prev_pseudo):
        a, b, c, d = C.c_uint64(prev_odd), C.c_uint64(prev_inside), C.c_uint64(error_mask), C.c_uint64(prev_pseudo)
        r = self.lib.sjo_find_structural_bits(_buf(in64), C.byref(a), C.byref(b), C.byref(c), 0, C.byref(d))
        return r, a.value, b.value, c.value, d.value

    def find_structural_bits_in_slice(self, buf, carried, position, ndjson=0, state=None):
        st = state or dict(prev_odd=0, prev_inside=0, error_mask=0, prev_pseudo=1)
        a, b = C.c_uint64(st["prev_odd"]), C.c_uint64(st["prev_inside"])
        c, d = C.c_uint64(st["error_mask"]), C.c_uint64(st["prev_pseudo"])
        idx = (C.c_uint32 * 1536)()
        n, car, pos = C.c_int(0), C.c_uint64(carried), C.c_uint64(position)
        processed = self.lib.sjo_find_structural_bits_in_slice(_buf(buf), len(buf), C.byref(a), C.byref(b), C.byref(c),
                                                               C.byref(d), idx, C.byref(n), C.byref(car), C.byref(pos),
                                                               ndjson)
        st.update(prev_odd=a.value, prev_inside=b.value, error_mask=c.value, prev_pseudo=d.value)
        return processed, list(idx[:n.value]), car.value, pos.value, st

    # ---- driver / stage 2 ------------------------------------------------------
    def find_structural_indices(self, msg, ndjson=False):
        import numpy as np
        out = np.empty(len(msg) + 64, dtype=np.uint32)
        n = C.c_size_t(0)
        ptr, keep = _ptr(msg)
        ok = self.lib.sjo_find_structural_indices(ptr, len(msg), int(ndjson), out.ctypes.data, out.size, C.byref(n))
        del keep
        return ok == 1, out[:n.value].copy()

    def parse_string_validate_only(self, buf, max_string_size):
        sl, dl = C.c_uint64(0), C.c_uint64(0)
        ok = self.lib.sjo_parse_string_validate_only(_buf(buf), len(buf), max_string_size, C.byref(sl), C.byref(dl))
        return bool(ok), sl.value, dl.value

    def parse_string(self, buf):
        dst = (C.c_uint8 * (len(buf) + 64))()
        dl = C.c_uint64(0)
        ok = self.lib.sjo_parse_string(_buf(buf), len(buf), dst, C.byref(dl))
        return bool(ok), bytes(dst[:dl.value]) if ok else b""

    def parse_number(self, buf):
        v = C.c_uint64(0)
        tag = self.lib.sjo_parse_number(_buf(buf), len(buf), C.byref(v))
        return tag, v.value

    def atom(self, kind, buf):
        return bool(getattr(self.lib, "sjo_is_valid_%s_atom" % kind)(_buf(buf), len(buf)))

    def trim_space(self, buf):
        a, b = C.c_size_t(0), C.c_size_t(0)
        self.lib.sjo_trim_space(_buf(buf), len(buf), C.byref(a), C.byref(b))
        return a.value, b.value

    def parse(self, msg, ndjson=False, copy_strings=True):
        """Returns (rc, tape uint64[], strings bytes, (msg_off, msg_len))."""
        import numpy as np
        n = len(msg)
        tape = np.empty(2 * n + 64, dtype=np.uint64)
        strings = np.empty(n + 64, dtype=np.uint8)
        tl, sl, mo, ml = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        flags = (FLAG_NDJSON if ndjson else 0) | (FLAG_COPY_STRINGS if copy_strings else 0)
        ptr, keep = _ptr(msg)
        rc = self.lib.sjo_parse(ptr, n, flags, tape.ctypes.data, tape.size, C.byref(tl), strings.ctypes.data,
                                strings.size, C.byref(sl), C.byref(mo), C.byref(ml))
        del keep
        if rc != 0:
            return rc, None, None, (mo.value, ml.value)
        return rc, tape[:tl.value].copy(), strings[:sl.value].tobytes(), (mo.value, ml.value)

    def count_where(self, tape, strings, message, key, value):
        """countWhere(key, value) + countObjects over a finished tape: (roots, matches)."""
        import numpy as np
        t = np.ascontiguousarray(tape, dtype=np.uint64)
        sb, mb = _buf(strings), _buf(message)
        roots = C.c_uint64(0)
        n = self.lib.sjo_count_where(t.ctypes.data, t.size, C.addressof(sb), C.addressof(mb), bytes(key), len(key),
                                     bytes(value), len(value), C.byref(roots))
        return roots.value, n

    def stage1_count(self, msg, ndjson=False):
        ok = C.c_int(0)
        ptr, keep = _ptr(msg)
        n = self.lib.sjo_stage1_count(ptr, len(msg), int(ndjson), C.byref(ok))
        del keep
        return n, ok.value == 1


def _ptr(msg):
    """(address, keep-alive) for bytes-like or numpy uint8 input."""
    try:
        import numpy as np
        if isinstance(msg, np.ndarray):
            return msg.ctypes.data, msg
    except ImportError:
        pass
    keep = _buf(msg)
    return C.addressof(keep), keep
