"""Shared helpers: golden-vector and data-fixture loading (no reference access)."""
import io
import json
import os
import tarfile

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA = os.path.join(GOLDEN, "data")

TAPE_FILES = ["apache_builds", "canada", "citm_catalog", "github_events", "gsoc-2018", "instruments", "numbers",
              "marine_ik", "mesh", "mesh.pretty", "twitterescaped", "twitter", "random", "update-center"]
SMALL_FILES = ["payload-small", "payload-medium", "payload-large"]


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def unhex(s):
    return bytes.fromhex(s)


def _zstd_bytes(path):
    import pyarrow as pa
    with pa.CompressedInputStream(pa.OSFile(path), "zstd") as f:
        return f.read()


_cache = {}


def load_fixture(name):
    """Decompressed bytes of tests/golden/data/<name>.json.zst (the reference's testdata)."""
    if name not in _cache:
        _cache[name] = _zstd_bytes(os.path.join(DATA, name + ".json.zst"))
    return _cache[name]


def fuzz_corpus(which="corpus", limit=None, max_size=None):
    """Yield (name, bytes) from the reference's fuzz seed corpora."""
    raw = _zstd_bytes(os.path.join(DATA, "fuzz-%s.tar.zst" % which))
    n = 0
    with tarfile.open(fileobj=io.BytesIO(raw), mode="r:") as tf:
        for m in tf:
            if not m.isfile():
                continue
            data = tf.extractfile(m).read()
            if which == "go-corpus":
                data = decode_go_fuzz(data)
                if data is None:
                    continue
            if max_size is not None and len(data) > max_size:
                continue
            yield m.name, data
            n += 1
            if limit is not None and n >= limit:
                return


def decode_go_fuzz(data):
    """'go test fuzz v1' file with one []byte("...") line -> bytes (else None)."""
    import ast
    lines = data.decode("utf-8", "replace").splitlines()
    if not lines or not lines[0].startswith("go test fuzz v1"):
        return None
    for ln in lines[1:]:
        ln = ln.strip()
        if ln.startswith("[]byte(") and ln.endswith(")"):
            lit = ln[len("[]byte("):-1]
            try:
                return _go_string(lit)
            except Exception:
                return None
    return None


def _go_string(lit):
    import re
    if lit.startswith("`"):
        return lit[1:-1].encode("utf-8")
    body = lit[1:-1]
    out = bytearray()
    i = 0
    simple = {"n": 10, "t": 9, "r": 13, "\\": 92, '"': 34, "'": 39, "a": 7, "b": 8, "f": 12, "v": 11}
    while i < len(body):
        c = body[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        e = body[i + 1]
        if e in simple:
            out.append(simple[e]); i += 2
        elif e == "x":
            out.append(int(body[i + 2:i + 4], 16)); i += 4
        elif e == "u":
            out += chr(int(body[i + 2:i + 6], 16)).encode("utf-8", "surrogatepass"); i += 6
        elif e == "U":
            out += chr(int(body[i + 2:i + 10], 16)).encode("utf-8", "surrogatepass"); i += 10
        elif e in "01234567":
            out.append(int(body[i + 1:i + 4], 8)); i += 4
        else:
            raise ValueError(e)
    return bytes(out)


def tricky_ndjson():
    """records that exercise FindKey's corner cases (parsed_object.go:97-140)"""
    recs = [
        b'{"Make":"HOND","x":1}',
        b'{"a":{"Make":"HOND"},"Make":"TOYT"}',              # nested member of the same name does not count
        b'{"a":[{"Make":"HOND"},2,[3]],"Make":"HOND"}',      # containers are skipped as a whole
        b'{"Make":"TOYT","Make":"HOND"}',                    # the FIRST member of that name decides
        b'{"Make":"HOND","Make":"TOYT"}',
        b'{"Make":1}', b'{"Make":null}', b'{"Make":["HOND"]}', b'{"Make":{"Make":"HOND"}}',
        b'{"make":"HOND"}', b'{"Make ":"HOND"}', b'{"Make":"HOND "}', b'{"Make":"HON"}', b'{"Mak":"HOND"}',
        b'["Make","HOND"]', b'[]', b'{}', b'[{"Make":"HOND"}]',
        b'{"t":true,"f":false,"n":null,"d":1.5,"l":-3,"u":18446744073709551615,"Make":"HOND"}',
        b'{"M\\u0061ke":"HOND"}',                             # escaped name: compared after unescaping
        b'{"Make":"HO\\u004eD"}',
        b'{"":"","Make":""}',
    ]
    return b"\n".join(recs), recs
