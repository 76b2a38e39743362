"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/simdjson_b200.h declares, refuses to work without a device (no CPU fallback), and its
host-only helper agrees with the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from simdjson_b200 import _lib
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "simdjson_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sj_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from simdjson_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.EXPORTS) == names


def test_no_device_means_no_service(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.sj_supported() == 0
    h = C.c_void_p()
    assert lib.sj_ctx_create(0, C.byref(h)) == 3  # SJ_ERR_NO_DEVICE: the product path has no CPU fallback
    assert not h.value


def test_trim_space_matches_oracle(lib, oracle):
    cases = [b"  {} \n", b"\xc2\xa0{}\xe2\x80\x83", b"\xff {} ", b" \t\r\n", b"\x0b\x0c[]\xe3\x80\x80", b"{}\xc2",
             b"\xe1\x9a\x80[1]\xc2\x85", b"", b"x", b"\xe2\x80\x8a[\xe2\x80\x8b]\xe2\x80\xa8"]
    for src in cases:
        a, b = C.c_size_t(0), C.c_size_t(0)
        buf = (C.c_uint8 * max(1, len(src))).from_buffer_copy(src or b"\0")
        lib.sj_trim_space(buf, len(src), C.byref(a), C.byref(b))
        assert (a.value, b.value) == oracle.trim_space(src), src


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """include/simdjson_b200.h compiled by a C11 compiler with -Werror, linked against the library from plain C
    (what cgo sees); without a device the example stops at SJ_ERR_NO_DEVICE -- no CPU fallback"""
    import subprocess
    import torch
    exe = str(tmp_path / "abi_example")
    libdir = os.path.join(ROOT, "simdjson-go_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "abi_example.c"), "-L" + libdir, "-lsimdjson_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "trimmed window: [1, 49)" in out.stdout, out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert out.returncode == 3 and "no sm_100 device" in out.stdout, out.stdout + out.stderr
