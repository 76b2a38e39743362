"""The streaming stage 2 (simdjson-go_b200/csrc/s2s_core.h, s2s_slab.h) executed on the CPU by a 32-fiber warp emulation
(tests/emu/s2s_emu.cpp compiles the very templates the CUDA kernels instantiate) and compared bit for bit with the oracle:
tape, string buffer and accept / reject.  This is how the kernels' logic is verified on a machine without a GPU; the GPU
suite (test_gpu_stage2.py) then runs the same inputs through the real kernels."""
import numpy as np
import pytest

from tests.emu_util import same_as_oracle
from tests.util import SMALL_FILES, TAPE_FILES, fuzz_corpus, golden, load_fixture, unhex, tricky_ndjson

SLAB, STEP = 6144, 2048


@pytest.mark.parametrize("name", TAPE_FILES + SMALL_FILES)
def test_fixtures(oracle_native, name):
    assert same_as_oracle(oracle_native, load_fixture(name)) == 0


def test_ndjson(oracle_native):
    pk = load_fixture("parking-citations").strip()
    assert same_as_oracle(oracle_native, pk, True) == 0
    assert same_as_oracle(oracle_native, b"\n".join([pk[:50000].rsplit(b"\n", 1)[0]] * 3), True) == 0
    nd, _ = tricky_ndjson()
    assert same_as_oracle(oracle_native, nd, True) == 0
    g = golden("G12_ndjson_tape")
    assert same_as_oracle(oracle_native, unhex(g["input"]), True) == 0
    for doc in (b'{"a":1}\n\n\n{"b":2}', b'{"a":1}\n{"b":2}\n', b'{"a":1} \n \n [1,2]\n{}', b'{"a":1}\n"str"\n{"b":2}', b'{"a":1}\n1\n{}',
                b'\n\n{"a":1}', b'{"a":1}{"b":2}', b'{"a":1}\n{"b":2}x', b'{"a":"x\\ny"}\n[]', b'[1]\n' * 3000, b'{}\n' + b" " * 7000 + b"\n[]",
                b'{"k":"' + b"v" * 7000 + b'"}\n{"k":2}'):
        same_as_oracle(oracle_native, doc, True)
        same_as_oracle(oracle_native, doc, False)


def test_golden_documents(oracle_native):
    g = golden("G16_G17_documents")
    n = 0
    for tc in g["parse_nd"]:
        same_as_oracle(oracle_native, unhex(tc["js"]), True)
        n += 1
    for tc in g["fail_cases"] + g["pass_cases"]:
        rc = same_as_oracle(oracle_native, unhex(tc["js"]))
        assert (rc != 0) == bool(tc["want_err"]), tc["name"]
        same_as_oracle(oracle_native, unhex(tc["js"]), True)
        n += 1
    for js in g["ndjson_emptylines"]:
        assert same_as_oracle(oracle_native, unhex(js), True) == 0
        n += 1
    assert n > 100


def _escape_soup(rng, n):
    alphabet = [b"a", b"b", b"\\\\", b'\\"', b"\\n", b"\\/", b"\\t", b"\\u00e9", b"\\u20AC", b"\\u0041", b"\\ud83d\\ude00", b"\\uD800\\uDC00",
                b"\\udbff\\u1234", b"\\ud800\\ud800\\udc00", b"\\udc00", b"\xc3\xa9", b"xyz" * 3, b" ", b"0123456789abcdef", b"\\u0022", b"\\u005c"]
    return b"".join(alphabet[j] for j in rng.integers(0, len(alphabet), n))


def test_escapes_across_every_edge(oracle_native):
    """escapes (simple, \\uXXXX with 1-3 byte results, surrogate pairs, chains of high surrogates) straddling 64-byte
    block, 2 KiB step and 6 KiB slab edges at every offset"""
    escapes = [b"\\n", b"\\u00e9", b"\\u20AC", b"\\u0041", b"\\ud83d\\ude00", b"\\udbff\\u1234", b"\\ud800\\ud800\\udc00", b"\\\\\\n", b'\\"']
    for edge in (64, 128, STEP, 2 * STEP, SLAB, SLAB + STEP, 2 * SLAB):
        for esc in escapes:
            for a in range(0, len(esc) + 1):  # `a` bytes of the escape in front of the edge
                pre = edge - 2 - a            # '["' is 2 bytes
                if pre < 0:
                    continue
                doc = b'["' + b"p" * pre + esc + b'tail","' + esc + b'",{"k' + esc + b'":"v"}]'
                assert same_as_oracle(oracle_native, doc) == 0, (edge, esc, a)
    rng = np.random.default_rng(5)
    for _ in range(60):
        body = _escape_soup(rng, int(rng.integers(50, 3000)))
        doc = b'{"' + body + b'":["' + _escape_soup(rng, int(rng.integers(1, 2000))) + b'","' + body[:100] + b'"]}'
        same_as_oracle(oracle_native, doc)


def test_invalid_escapes_and_strings(oracle_native):
    bad = [b'["\\q"]', b'["\\u12"]', b'["\\u12G4"]', b'["\\ud800"]', b'["\\ud800\\n"]', b'["\\ud800x\\udc00"]', b'["\\u00"', b'["\\u"]',
           b'["a\\"]', b'["\\ud83d\\ude0"]', b'["\\ud83d\\u"]', b'["\\u+123"]', b'["\\u 123"]', b'["\\u0\\"00"]', b'["\\ud83d\\ud83d"]',
           b'["\\ud83d\\ud83d\\ude00"]', b'["\\udfff\\udfff"]', b'["\\u-123"]']
    for d in bad:
        same_as_oracle(oracle_native, d)
        for edge in (64, STEP, SLAB):
            for a in range(0, 13):
                pre = edge - 2 - a
                same_as_oracle(oracle_native, b'["' + b"p" * pre + d[2:])


def test_strings_across_edges(oracle_native):
    """strings that open in one lane / step / slab and close in another, empty strings at the edges, quotes at the last
    and first byte of blocks, steps and slabs"""
    for edge in (64, STEP, SLAB, 3 * SLAB):
        for ln in (0, 1, 2, 62, 63, 64, 65, 127, 128, 129, 2047, 2048, 2049, 6143, 6144, 6145, 13000):
            for a in (0, 1, 2, 3, 63, 64):
                pre = edge - a
                if pre < 2:
                    continue
                doc = b"[" + b" " * (pre - 2) + b'"' + b"s" * ln + b'","' + b"t" * (ln % 7) + b'",1,true,null,"","' + b"u" * ln + b'"]'
                assert same_as_oracle(oracle_native, doc) == 0


def test_structure_and_grammar(oracle_native):
    docs = [b"[" * 3000 + b"]" * 3000, b"[" + b"[1," * 2500 + b"1" + b"]" * 2500 + b"]", b"[" + b",".join([b"[]"] * 5000) + b"]",
            b"[" + b",".join([b'{"a":[1,2],"b":{}}'] * 1500) + b"]", b'{"a":{"b":{"c":[{"d":[[[{"e":null}]]]}]}}}',
            b'{"a":1,"b":2', b'{"a":1,,"b":2}', b'{"a":1 "b":2}', b'{"a","b":2}', b'["a":1]', b'[1,2}', b'{"a":[1,2}', b"[1 2]", b"[1,]", b"[,1]", b'{"a":}',
            b'{:1}', b'{"a" 1}', b"[tru]", b"[truex]", b"[nul]", b"[falsey]", b"[-]", b"[1.e3]", b"[01]", b"[1e+1111]", b"[x]", b'{"a":1}}', b"]", b"[]]",
            b'{"a":"b","c":{"d":["e",{"f":"g"}]},"h":[]}', b'"str"', b"1", b"[1]x", b"[1] 2", b'{"a":1}[]']
    for d in docs:
        same_as_oracle(oracle_native, d)
        same_as_oracle(oracle_native, d, True)
    rng = np.random.default_rng(11)
    toks = [b"{", b"}", b"[", b"]", b":", b",", b'"k"', b'"v\\n"', b"1", b"-2.5e3", b"true", b"false", b"null", b" ", b"\n", b"x", b'"', b"\\"]
    for _ in range(400):
        d = b"".join(toks[j] for j in rng.integers(0, len(toks), int(rng.integers(1, 60))))
        same_as_oracle(oracle_native, d)
        same_as_oracle(oracle_native, d, True)
    # valid random documents with every construct, several slabs long
    def gen(depth):
        r = rng.integers(0, 10)
        if depth > 4 or r < 3:
            return [b"1", b"-0.5", b"true", b"false", b"null", b'"s"', b'"\\u00e9\\n"', b'""', b"12345678901234567890", b"1e300"][int(rng.integers(0, 10))]
        if r < 6:
            return b"[" + b",".join(gen(depth + 1) for _ in range(int(rng.integers(0, 6)))) + b"]"
        return b"{" + b",".join(b'"k%d":' % i + gen(depth + 1) for i in range(int(rng.integers(0, 6)))) + b"}"
    for _ in range(30):
        d = b"[" + b" ,\n".join(gen(0) for _ in range(int(rng.integers(1, 200)))) + b"]"
        assert same_as_oracle(oracle_native, d) == 0


@pytest.mark.parametrize("which,limit", [("corpus", 1200), ("go-corpus", 356)])
def test_fuzz_corpus_sample(oracle_native, which, limit):
    n = 0
    for name, data in fuzz_corpus(which, limit=limit, max_size=120_000):
        same_as_oracle(oracle_native, data)
        if n % 4 == 0:
            same_as_oracle(oracle_native, data, True)
        n += 1
    assert n > 300


def test_lane_order_does_not_matter():
    """the fibers of the emulated warp run in ascending lane order between two collectives; correct GPU code cannot depend on
    that (escapes are patched into the shared image in place while other lanes still read it), so the edge and escape
    cases run once more with the lanes in DESCENDING order (-DS2S_EMU_REVERSE build of the emulator)"""
    import os
    import subprocess
    import sys
    if os.environ.get("S2S_EMU_FLAGS"):
        pytest.skip("already a variant build")
    env = dict(os.environ, S2S_EMU_FLAGS="-DS2S_EMU_REVERSE")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_s2s_emulation.py", "-x", "-q", "-k",
                          "escapes or strings_across or twitterescaped or ndjson or golden"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]


def test_escape_fast_path_is_the_definition():
    """esc_u_fast (word loads + SWAR hex digits) must agree with esc_decode wherever it claims to apply: random steps dense
    in "\\uXXXX" with proper, improper and quirky digits (bytes below '0', raw quotes, upper / lower case, >= 0x80),
    surrogates and look-alikes of a high surrogate six bytes in front, at every alignment and at both ends of the step"""
    import ctypes as C
    from tests import emu_util
    L = emu_util.lib()
    L.s2s_emu_esc_fast_check.restype = C.c_long
    L.s2s_emu_esc_fast_check.argtypes = [C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(20260923)
    digits = b"0123456789abcdefABCDEF"
    odd = b"gG/:@`\"' \x00\x10\x19\x2f\x3a\x7f\x80\xff\\u"
    applied = 0
    for it in range(400):
        out = bytearray()
        while len(out) < 2048 + 16:
            k = int(rng.integers(0, 10))
            if k < 6:
                hx = bytes(digits[int(j)] for j in rng.integers(0, len(digits), 4))
                if rng.integers(0, 5) == 0:
                    hx = (b"d" if rng.integers(0, 2) else b"D") + bytes([b"89abAB cdefCDEF"[int(rng.integers(0, 15))]]) + hx[2:]
                if rng.integers(0, 6) == 0:
                    j = int(rng.integers(0, 4))
                    hx = hx[:j] + bytes([odd[int(rng.integers(0, len(odd)))]]) + hx[j + 1:]
                out += b"\\u" + hx
            elif k < 8:
                out += bytes(rng.integers(0x20, 0x7f, int(rng.integers(1, 7)), dtype=np.uint8).tolist()).replace(b"\\", b"x")
            else:
                out += b"\\" + bytes([b'nrt"/\\bfq'[int(rng.integers(0, 9))]])
        step = np.frombuffer(bytes(out[:2048]), dtype=np.uint8).copy()
        avail = 2048 if it % 3 else int(rng.integers(1, 2049))
        r = L.s2s_emu_esc_fast_check(step.ctypes.data, avail)
        assert r >= 0, (it, -r - 1, bytes(step[max(0, -r - 8): -r + 8]))
        applied += r
    assert applied > 20000


def test_atom_fast_path_is_the_definition():
    """atom_ok_fast (word compares on the step image) against atom_ok_p (stage2_build_tape_amd64.go:124-158, 455-476): steps
    full of true / false / null, near misses and every kind of byte behind them, at every alignment and near the end"""
    import ctypes as C
    from tests import emu_util
    L = emu_util.lib()
    L.s2s_emu_atom_fast_check.restype = C.c_long
    L.s2s_emu_atom_fast_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
    rng = np.random.default_rng(7)
    words = [b"true", b"false", b"null", b"tru", b"fals", b"nul", b"trve", b"falsf", b"nulL", b"True", b"t", b"f", b"n", b"truetrue", b"nullfalse"]
    follow = [bytes([c]) for c in (0, 9, 10, 13, 32, 44, 58, 91, 93, 123, 125, 11, 12, 31, 33, 34, 43, 45, 59, 92, 94, 122, 124, 126, 127, 128, 255, 48, 101)]
    applied = 0
    for it in range(300):
        out = bytearray()
        while len(out) < 2048 + 16:
            out += words[int(rng.integers(0, len(words)))] + follow[int(rng.integers(0, len(follow)))] * int(rng.integers(1, 3))
        step = np.frombuffer(bytes(out[:2048]), dtype=np.uint8).copy()
        avail = 2048 if it % 3 else int(rng.integers(1, 2049))
        r = L.s2s_emu_atom_fast_check(step.ctypes.data, avail, avail)
        assert r >= 0, (it, -r - 1, bytes(step[max(0, -r - 4): -r + 10]))
        applied += r
    assert applied > 50000
