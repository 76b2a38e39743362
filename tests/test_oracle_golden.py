"""Pins the CPU oracle against the reference's own golden vectors (SURVEY.md
appendix D, G1..G19), extracted by tests/golden/make_golden.py.  CPU only.

Each test names the reference test it replays.
"""
import json
import struct

import numpy as np
import pytest

from tests.util import SMALL_FILES, TAPE_FILES, golden, load_fixture, tricky_ndjson, unhex

M64 = (1 << 64) - 1


@pytest.fixture(params=["scalar", "native", "avx512"])
def o(request, oracle, oracle_native):
    """the three builds of the oracle: scalar restatement, AVX2+PCLMUL mask routines, AVX-512BW mask routines (the
    reference's *_avx512 path; skipped on hosts without AVX-512BW)"""
    if request.param == "avx512":
        from oracle.pyoracle import Oracle, host_has_avx512
        if not host_has_avx512():
            pytest.skip("host CPU has no AVX-512BW")
        return Oracle("avx512")
    return oracle if request.param == "scalar" else oracle_native


def test_g1_finalize_structurals(o):
    """find_subroutines_amd64_test.go:32 TestFinalizeStructurals"""
    for i, tc in enumerate(golden("G1_finalize_structurals")):
        got, pp = o.finalize_structurals(tc["structurals"], tc["whitespace"], tc["quote_mask"], tc["quote_bits"], 0)
        assert (got, pp) == (tc["expected_strls"], tc["expected_pseudo"]), i


def test_g2_newline_delimiters(o):
    """find_subroutines_amd64_test.go:73 testFindNewlineDelimiters, :103 quoted newline"""
    g = golden("G2_newline_delimiters")
    nd = unhex(g["input"])
    for k, off in enumerate(range(0, len(nd) - 64, 64)):
        assert o.find_newline_delimiters(nd[off:off + 64], 0) == g["want"][k]
    q = bytearray(unhex(g["quoted_case"]["input"]))
    for p in g["quoted_case"]["newline_at"]:
        q[p] = 0x0A
    qm, _, _, _ = o.find_quote_mask_and_bits(bytes(q), 0, 0)
    assert o.find_newline_delimiters(bytes(q), qm) == g["quoted_case"]["want"]


def test_g3_odd_backslash(o):
    """find_subroutines_amd64_test.go:145 testFindOddBackslashSequences (+ the 1..128 sweep)"""
    for i, tc in enumerate(golden("G3_odd_backslash")):
        got, carry = o.find_odd_backslash_sequences(unhex(tc["input"]), tc["prev_ends_odd"])
        assert (got, carry) == (tc["expected"], tc["ends_odd_backslash"]), i
    for i in range(1, 129):
        t = b" " * (i - 1) + b'\\"' + b" " * (62 + 64)
        lo, c = o.find_odd_backslash_sequences(t[:64], 0)
        hi, c = o.find_odd_backslash_sequences(t[64:128], c)
        assert (lo, hi) == ((1 << i, 0) if i < 64 else (0, (1 << (i - 64)) & M64)), i  # Go: uint64 shift wraps to 0 at i=128


def test_g4_quote_mask_and_bits(o):
    """find_subroutines_amd64_test.go:215 testFindQuoteMaskAndBits"""
    g = golden("G4_quote_mask")
    for i, tc in enumerate(g["cases"]):
        qm, qb, piiq, em = o.find_quote_mask_and_bits(unhex(tc["input"]), tc["odd_ends"], 0)
        assert (qm, qb, piiq, em) == (tc["expected"], tc["quote_bits"], tc["piiq"], tc["error_mask"]), i
    for i, tc in enumerate(g["piiq_cases"]):
        _, _, piiq, _ = o.find_quote_mask_and_bits(unhex(tc["input"]), 0, tc["piiq_in"])
        assert piiq == tc["piiq_out"], i


def test_g5_whitespace_and_structurals(o):
    """find_subroutines_amd64_test.go:643 testFindWhitespaceAndStructurals"""
    for i, tc in enumerate(golden("G5_whitespace_structurals")):
        ws, st = o.find_whitespace_and_structurals(unhex(tc["input"])[:64])
        assert (ws, st) == (tc["ws"], tc["structurals"]), i


def test_g6_fused_equals_composed(o):
    """find_subroutines_amd64_test.go:311 testFindStructuralBits"""
    fused = dict(a=0, b=0, c=0, d=1)
    comp = dict(odd=0, inside=0, err=0, pseudo=1)
    for hx in golden("G6_fused_inputs"):
        blk = unhex(hx)
        s1, fused["a"], fused["b"], fused["c"], fused["d"] = o.find_structural_bits(blk, fused["a"], fused["b"],
                                                                                     fused["c"], fused["d"])
        oe, comp["odd"] = o.find_odd_backslash_sequences(blk, comp["odd"])
        qm, qb, comp["inside"], comp["err"] = o.find_quote_mask_and_bits(blk, oe, comp["inside"], comp["err"])
        ws, st = o.find_whitespace_and_structurals(blk)
        s2, comp["pseudo"] = o.finalize_structurals(st, ws, qm, qb, comp["pseudo"])
        assert s1 == s2


def test_g7_tail_whitespace_padding(o):
    """find_subroutines_amd64_test.go:373 testFindStructuralBitsWhitespacePadding"""
    msg = unhex(golden("G7_tail_padding")["msg"])
    for l in range(len(msg), -1, -1):
        processed, idx, carried, _, _ = o.find_structural_bits_in_slice(msg[:l], M64, M64)
        assert processed == l and len(idx) == l
        last = sum(idx) & M64
        if l > 0:
            assert last == l - 1
        else:
            assert last == ((l - 1) - carried) & M64


def test_g8_twitter_structural_loop(o):
    """find_subroutines_amd64_test.go:427 testFindStructuralBitsLoop"""
    g = golden("G8_twitter_loop")
    msg = load_fixture("twitter")
    indexes, processed, carried, position, st = [], 0, M64, M64, None
    while processed < len(msg):
        p, idx, carried, position, st = o.find_structural_bits_in_slice(msg[processed:], carried, position, 0, st)
        processed += p
        indexes += idx
    assert len(indexes) == g["count"]
    pos = len(msg) - 1
    for j, ch in enumerate(g["reversed_tail"]):
        assert msg[pos:pos + 1].decode() == ch
        pos -= indexes[len(indexes) - 1 - j]


def test_g9_flatten_bits_incremental(o):
    """find_subroutines_amd64_test.go:706 TestFlattenBitsIncremental"""
    for i, tc in enumerate(golden("G9_flatten_bits")):
        got, _, _ = o.flatten_bits(tc["masks"], 0, M64)
        assert got == tc["expected"], i


def _rev64(x):
    return int("{:064b}".format(x)[::-1], 2)


def test_g10_stage1_marks_and_indices(o):
    """stage1_find_marks_amd64_test.go:28 TestStage1FindMarks, :86 TestFindStructuralIndices"""
    g = golden("G10_stage1_marks")
    demo = unhex(g["demo_json"])
    blk = demo[:64]
    want = {k: int(v, 2) for k, v in g["masks_msb_first_reversed"].items()}
    oe, _ = o.find_odd_backslash_sequences(blk, 0)
    assert oe == 0
    qm, qb, _, _ = o.find_quote_mask_and_bits(blk, oe, 0)
    assert _rev64(qm) == want["quoted"]
    ws, st = o.find_whitespace_and_structurals(blk)
    assert _rev64(st) == want["structurals"] and _rev64(ws) == want["whitespace"]
    fin, _ = o.finalize_structurals(st, ws, qm, qb, 0)
    assert _rev64(fin) == want["structurals_finalized"]
    ok, deltas = o.find_structural_indices(demo)
    assert ok
    pos = (np.cumsum(deltas.astype(np.int64)) - 1).tolist()
    assert pos == g["positions"]


def test_g11_stage2_tapes(o):
    """stage2_build_tape_amd64_test.go:26 TestStage2BuildTape (no-copy tapes)"""
    for i, tc in enumerate(golden("G11_tapes")):
        rc, tape, strs, _ = o.parse(unhex(tc["input"]), copy_strings=False)
        assert rc == 0, i
        assert [int(x) for x in tape] == tc["tape"], i


def test_g12_demo_ndjson_tape(o):
    """ndjson_test.go:36 verifyDemoNdjson / parse_json_amd64_test.go:34 TestDemoNdjson"""
    g = golden("G12_ndjson_tape")
    rc, tape, _, _ = o.parse(unhex(g["input"]), ndjson=True, copy_strings=False)
    assert rc == 0
    assert [int(x) for x in tape] == list(g["tape"]) and len(tape) == 153


def test_g13_atoms(o):
    """stage2_build_tape_amd64_test.go:195-262"""
    g = golden("G13_atoms")
    for kind in ("true", "false", "null"):
        for tc in g[kind]:
            assert o.atom(kind, unhex(tc["input"])) == tc["expected"], (kind, tc)


def test_g14_strings(o):
    """parse_string_test.go:19 tests, driven like parse_json_amd64_test.go:540 / :568"""
    for tc in golden("G14_strings"):
        buf = b'"' + unhex(tc["str"]) + b'"'
        ok, sl, dl = o.parse_string_validate_only(buf, len(buf))
        assert ok == tc["success"], tc["name"]
        if ok:
            want = unhex(tc["want"])
            assert dl == len(want), tc["name"]
            ok2, out = o.parse_string(buf)
            assert ok2 and out == want, tc["name"]


def _tagname(tag):
    return chr(tag >> 56) if tag else ""


def test_g15_numbers(o):
    """parse_json_amd64_test.go:222 TestParseNumber, :320 TestParseInt64, :504 TestParseFloat64,
    parse_number_test.go:30 TestNumberIsValid"""
    g = golden("G15_numbers")
    for tc in g["parse_number"]:
        tag, val = o.parse_number(tc["input"].encode() + b":")
        assert _tagname(tag) == tc["tag"], tc
        assert tag & ((1 << 56) - 1) == tc["flags"], tc
        if tc["tag"] == "d":
            assert struct.unpack("<d", struct.pack("<Q", val))[0] == float(tc["d"]), tc
        elif tc["tag"] == "l":
            assert val == tc["i"] & M64, tc
        else:
            assert val == tc["u"], tc
    for tc in g["parse_int64"]:
        tag, val = o.parse_number(tc["input"].encode() + b":")
        assert _tagname(tag) == tc["tag"], tc
        if tc["tag"] == "l":
            assert val == tc["out"] & M64, tc
    for tc in g["atof"]:
        tag, val = o.parse_number(tc["input"].encode() + b":")
        t = _tagname(tag)
        if t == "":
            assert tc["err"], tc
        elif t == "d":
            want = float(tc["out"].replace("+Inf", "inf").replace("-Inf", "-inf"))
            assert struct.pack("<Q", val) == struct.pack("<d", want), tc
        else:
            assert str(val if t == "u" else struct.unpack("<q", struct.pack("<Q", val))[0]) == tc["out"], tc
    for s in g["valid"]:
        assert o.parse_number(s.encode())[0] != 0, s
    for s in g["invalid"]:
        assert o.parse_number(s.encode())[0] == 0, s


def _tape_to_py(tape, strings, msg):
    """Minimal Iter.Interface() equivalent for checking parsed values."""
    STRBIT = 1 << 55
    i, out = 0, []

    def val(i):
        w = int(tape[i])
        t, p = chr(w >> 56), w & ((1 << 56) - 1)
        if t == '"':
            ln = int(tape[i + 1])
            s = strings[p - STRBIT:p - STRBIT + ln] if p & STRBIT else msg[p:p + ln]
            return s.decode("utf-8", "surrogatepass"), i + 2
        if t == "l":
            return struct.unpack("<q", struct.pack("<Q", int(tape[i + 1])))[0], i + 2
        if t == "u":
            return int(tape[i + 1]), i + 2
        if t == "d":
            return struct.unpack("<d", struct.pack("<Q", int(tape[i + 1])))[0], i + 2
        if t in "tfn":
            return {"t": True, "f": False, "n": None}[t], i + 1
        if t == "[":
            arr, i2 = [], i + 1
            while chr(int(tape[i2]) >> 56) != "]":
                v, i2 = val(i2)
                arr.append(v)
            assert p == i2 + 1 and int(tape[i2]) & ((1 << 56) - 1) == i
            return arr, i2 + 1
        if t == "{":
            obj, i2 = {}, i + 1
            while chr(int(tape[i2]) >> 56) != "}":
                k, i2 = val(i2)
                v, i2 = val(i2)
                obj[k] = v
            assert p == i2 + 1 and int(tape[i2]) & ((1 << 56) - 1) == i
            return obj, i2 + 1
        raise AssertionError("bad tag %r at %d" % (t, i))

    while i < len(tape):
        w = int(tape[i])
        assert chr(w >> 56) == "r"
        v, j = val(i + 1)
        assert chr(int(tape[j]) >> 56) == "r" and int(tape[j]) & ((1 << 56) - 1) == i
        assert w & ((1 << 56) - 1) == j + 1
        out.append(v)
        i = j + 1
    return out


def _same(a, b):
    """Deep equality; an overflowed integer parsed as float64 equals Python's big int numerically."""
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, bool) or isinstance(b, bool) or a is None or b is None:
        return a is b
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return float(a) == float(b)
    return a == b


def test_g16_g17_documents(o):
    """simdjson_amd64_test.go:29 TestParseND, :162 TestParseFailCases, :695 TestParsePassCases;
    parse_json_amd64_test.go:47 TestNdjsonEmptyLines"""
    g = golden("G16_G17_documents")
    for key, nd in (("fail_cases", False), ("pass_cases", False), ("parse_nd", True)):
        for tc in g[key]:
            js = unhex(tc["js"])
            for copy in (True, False):
                rc, tape, strs, (off, ln) = o.parse(js, ndjson=nd, copy_strings=copy)
                assert (rc != 0) == tc["want_err"], (key, tc["name"], rc)
                if rc == 0 and tc["want"] is not None:
                    got = _tape_to_py(tape, strs, js[off:off + ln])
                    want = [json.loads(l) for l in unhex(tc["want"]).decode().split("\n")] if nd else \
                        [json.loads(unhex(tc["want"]).decode())]
                    assert _same(got, want), (key, tc["name"])
    for hx in g["ndjson_emptylines"]:
        rc, tape, _, _ = o.parse(unhex(hx), ndjson=True)
        assert rc == 0 and sum(1 for w in tape if int(w) >> 56 == ord("r")) == 4


def test_g18_parking_citations(o):
    """ndjson_test.go:250 TestNdjsonCountWhere: 1000 roots, Make == HOND 116 times"""
    g = golden("G18_G19_fixtures")["parking_citations"]
    msg = load_fixture("parking-citations")
    rc, tape, strs, (off, ln) = o.parse(msg, ndjson=True)
    assert rc == 0
    recs = _tape_to_py(tape, strs, msg[off:off + ln])
    assert len(recs) == g["roots"]
    assert sum(1 for r in recs if r.get("Make") == "HOND") == g["make_hond"]


@pytest.mark.parametrize("name", TAPE_FILES + SMALL_FILES)
def test_g19_verify_tape_files(o, name):
    """parse_json_amd64_test.go:682 TestVerifyTape: every fixture parses; values equal json.loads"""
    msg = load_fixture(name)
    rc, tape, strs, (off, ln) = o.parse(msg)
    assert rc == 0
    if len(msg) < 700_000:
        assert _same(_tape_to_py(tape, strs, msg[off:off + ln]), [json.loads(msg)])


def test_twitter_vs_twitterescaped_identical(o):
    """SURVEY.md 8c invariant: same content, one escaped => identical tape and Strings.B (copy mode)"""
    a = o.parse(load_fixture("twitter"))
    b = o.parse(load_fixture("twitterescaped"))
    assert a[0] == b[0] == 0
    assert np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_trim_space(o):
    """Go bytes.TrimSpace incl. the unicode.IsSpace fall-back"""
    cases = [(b"  {} \n", b"{}"), (b"\xc2\xa0{}\xe2\x80\x83", b"{}"), (b"\xff {} ", b"\xff {}"), (b" \t\r\n", b""),
             (b"\x0b\x0c[]\xe3\x80\x80", b"[]"), (b"{}\xc2", b"{}\xc2"), (b"\xe1\x9a\x80[1]\xc2\x85", b"[1]")]
    for src, want in cases:
        a, b = o.trim_space(src)
        assert src[a:b] == want, src


def test_count_where_oracle(o):
    """countWhere / countObjects (ndjson_test.go:421-474) restated in the oracle: the reference's golden
    (1000 roots, Make == HOND 116 times, ndjson_test.go:250-266) and agreement with a plain json.loads walk"""
    g = golden("G18_G19_fixtures")["parking_citations"]
    msg = load_fixture("parking-citations")
    for copy in (True, False):
        rc, tape, strs, (off, ln) = o.parse(msg, ndjson=True, copy_strings=copy)
        assert rc == 0
        assert o.count_where(tape, strs, msg[off:off + ln], b"Make", b"HOND") == (g["roots"], g["make_hond"])
    nd, recs = tricky_ndjson()
    for copy in (True, False):
        rc, tape, strs, (off, ln) = o.parse(nd, ndjson=True, copy_strings=copy)
        assert rc == 0
        for key, value in ((b"Make", b"HOND"), (b"Make", b""), (b"", b""), (b"Make", b"TOYT"), (b"x", b"1")):
            want = 0
            for r in recs:
                first = {}
                for k, v in json.loads(r, object_pairs_hook=list) if r.startswith(b"{") else []:
                    first.setdefault(k, v)
                v = first.get(key.decode())
                want += isinstance(v, str) and v == value.decode()
            assert o.count_where(tape, strs, nd[off:off + ln], key, value) == (len(recs), want), (key, value, copy)
