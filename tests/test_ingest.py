"""Fast ingest of the reference's text formats (include/gsgp_host.h) against the oracle's plain restatements of
read_input_data (main_cpu.go:381-421) and read_sem (:710-733): bit-identical values, same failure cases."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as ob
from go_gsgp_b200 import build
from go_gsgp_b200.host import read_dataset, read_semantic


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()
    ob.build()


def write_dataset(path, X, y, sep="\t"):
    with open(path, "w") as f:                                # README.md:31-42 / runner.py:213 (tabs)
        f.write(f"{X.shape[1]}\n{X.shape[0]}\n")
        for r, t in zip(X, y):
            f.write(sep.join(repr(float(v)) for v in r) + sep + repr(float(t)) + "\n")


def oracle_read_dataset(train, test):
    L = ob.lib()
    nv, nr, nt = C.c_int32(), C.c_int32(), C.c_int32()
    ptr = C.POINTER(C.c_double)()
    L.orc_read_dataset.restype = C.c_int
    L.orc_read_dataset.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.POINTER(C.c_double))]
    rc = L.orc_read_dataset(train.encode(), test.encode(), C.byref(nv), C.byref(nr), C.byref(nt), C.byref(ptr))
    assert rc == 0
    n = nr.value + nt.value
    return np.ctypeslib.as_array(ptr, shape=(n, nv.value + 1)).copy(), nr.value


def test_dataset_files_match_the_reference_reader(tmp_path):
    rng = np.random.default_rng(2)
    X = rng.normal(size=(3000, 7)) * 10.0 ** rng.integers(-8, 8, size=(3000, 7))
    y = rng.normal(size=3000)
    X[5, 2], X[6, 3], y[7] = 0.0, -0.0, 1e-310                  # zeros and a subnormal
    tr, te = str(tmp_path / "train.txt"), str(tmp_path / "test.txt")
    write_dataset(tr, X[:2400], y[:2400])
    write_dataset(te, X[2400:], y[2400:], sep="  ")
    Xr, yr, nrow = read_dataset(tr, te)
    assert nrow == 2400 and Xr.shape == X.shape
    np.testing.assert_array_equal(Xr.view(np.uint64), X.view(np.uint64))          # repr round-trips: bit-identical
    np.testing.assert_array_equal(yr.view(np.uint64), y.view(np.uint64))
    m, onrow = oracle_read_dataset(tr, te)
    assert onrow == nrow
    np.testing.assert_array_equal(np.concatenate([Xr, yr[:, None]], axis=1).view(np.uint64), m.view(np.uint64))
    # failure cases the reference panics on
    with pytest.raises(RuntimeError, match="cannot open"):
        read_dataset(tr, str(tmp_path / "missing.txt"))
    bad = str(tmp_path / "bad.txt")
    write_dataset(bad, X[:10, :3], y[:10])
    with pytest.raises(RuntimeError, match="same number of variables"):
        read_dataset(tr, bad)
    open(bad, "w").write("2\n3\n1.0 2.0 3.0\n4.0 5.0\n")
    with pytest.raises(RuntimeError, match="not enough values"):
        read_dataset(bad, bad)


def test_semantic_files(tmp_path):
    rng = np.random.default_rng(4)
    n = 200_000
    sem = rng.normal(size=n) * 10.0 ** rng.integers(-12, 12, size=n)
    p = str(tmp_path / "mod_0_1.dat")
    with open(p, "w") as f:
        for i, v in enumerate(sem):
            f.write(("+" if i % 1000 == 1 and v > 0 else "") + (f"{v:.17e}" if i % 3 else repr(float(v))) + "\n")
    got = read_semantic(p, n)
    np.testing.assert_array_equal(got.view(np.uint64), sem.view(np.uint64))
    ref = np.empty(n)
    L = ob.lib()
    L.orc_read_sem.restype, L.orc_read_sem.argtypes = C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_double)]
    assert L.orc_read_sem(p.encode(), n, ref.ctypes.data_as(C.POINTER(C.c_double))) == 0
    np.testing.assert_array_equal(got.view(np.uint64), ref.view(np.uint64))
    np.testing.assert_array_equal(read_semantic(p, 1000), sem[:1000])             # a prefix is fine (:719)
    with pytest.raises(RuntimeError, match="Not enough values"):
        read_semantic(p, n + 1)
    open(p, "a").write("abc\n")
    with pytest.raises(RuntimeError, match="Cannot parse"):
        read_semantic(p, n + 1)


@pytest.mark.parametrize("suffix", [".txt", ".txt.gz"])
def test_async_line_writer_matches_the_reference_format(tmp_path, suffix):
    """gsgph_writer: best-individual semantic lines (Semantic.String() + newline, main_cpu.go:123-126,1499-1500) and
    fitness lines (%v, :1495-1496), gzip when the path ends in .gz (:1236-1238).  Large lines are compressed as several
    gzip members in parallel: the file must still read back as one stream, byte for byte what the synchronous
    formatter produces."""
    import gzip
    from go_gsgp_b200 import host
    rng = np.random.default_rng(9)
    rows = [rng.normal(size=7), np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e21, 1e-7, 123456789.0, 5e-324]),
            rng.normal(size=700_000) * 1e3,                        # ~13 MB of text: several 4 MB gzip members
            np.array([]), rng.uniform(-1, 1, 33)]
    path = tmp_path / ("semantictrain" + suffix)
    with host.LineWriter(path, n_threads=4) as w:
        for r in rows:
            w.write_semantic(r)
    want = "".join(host.format_semantic(r) + "\n" if len(r) else "\n" for r in rows)
    opener = gzip.open if suffix.endswith(".gz") else open
    with opener(path, "rt") as f:
        got = f.read()
    assert got == want
    if suffix.endswith(".gz"):
        raw = open(path, "rb").read()
        assert raw.count(b"\x1f\x8b\x08") >= 4                      # parallel members
        assert len(raw) < 0.6 * len(want)                           # fastest deflate level: digits compress to ~52 %
    fpath = tmp_path / ("fitnesstrain" + suffix)
    vals = [0.5, 1e-9, 12345.678, np.nan, np.inf, 3.0]
    with host.LineWriter(fpath) as w:
        for v in vals:
            w.write_float(v)
    with opener(fpath, "rt") as f:
        assert f.read() == "".join(host.format_float(v) + "\n" for v in vals)


def test_line_writer_reports_unwritable_paths(tmp_path):
    from go_gsgp_b200 import host
    with pytest.raises(OSError):
        host.LineWriter(tmp_path / "no_such_dir" / "x.txt.gz")


@pytest.mark.parametrize("threads", [1, 2, 3, 8])
def test_parse_floats_token_boundaries(threads):
    """gsgph_parse_floats: a vectorised counting pass decides where each thread's values go, then every thread parses
    its byte range in place.  Tokens must be counted and parsed once whichever byte a range boundary falls on: random
    whitespace runs (all six kinds), tokens of every length, no trailing newline, a prefix read (cap), and a malformed
    token that only matters when it lies within the values asked for."""
    from go_gsgp_b200.host import _lib
    L = _lib()
    rng = np.random.default_rng(100 + threads)
    vals = rng.normal(size=30_000) * 10.0 ** rng.integers(-30, 30, size=30_000)
    vals[::97] = np.round(vals[::97])                               # short tokens: "3.0", "-0.0"
    ws = [" ", "\n", "\t", "\r", "\v", "\f", "  \n", "\t \t", "\r\n"]
    toks = [("+" if v > 0 and i % 50 == 0 else "") + (repr(float(v)) if i % 4 else f"{v:.17e}") for i, v in enumerate(vals)]
    text = (ws[3] + "".join(t + ws[int(k)] for t, k in zip(toks, rng.integers(0, len(ws), size=len(toks))))).rstrip() .encode()
    assert len(text) > (1 << 16)                                    # above the single-thread cut-off

    def parse(buf, cap):
        out, n = np.full(max(cap, 1), -7.0), C.c_int64()
        rc = L.gsgph_parse_floats(buf, len(buf), out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(n), threads)
        return rc, n.value, out

    rc, n, out = parse(text, len(vals))
    assert rc == 0 and n == len(vals)
    np.testing.assert_array_equal(out.view(np.uint64), vals.view(np.uint64))
    rc, n, out = parse(text, 1234)                                  # a prefix: the rest is counted, not stored
    assert rc == 0 and n == len(vals)
    np.testing.assert_array_equal(out[:1234], vals[:1234])
    bad = text + b" 1.5abc 2.0"
    rc, n, _ = parse(bad, len(vals))                                # malformed token beyond the values asked for: ignored
    assert rc == 0 and n == len(vals) + 2
    rc, n, _ = parse(bad, len(vals) + 2)
    assert rc == 2
    for junk in (b"1e", b"--1", b"0x10", b"1,5", b"."):
        rc, _, _ = parse(text[: 70_000].rsplit(b" ", 1)[0] + b" " + junk + b" " + text[70_000:], len(vals) + 1)
        assert rc == 2, junk
    rc, n, out = parse(b"  \n\t ", 4)
    assert rc == 0 and n == 0
    rc, n, out = parse(b"1e999 -1e999 1e-999 inf nan", 5)            # out of range keeps ParseFloat's value
    assert rc == 0 and n == 5 and out[0] == np.inf and out[1] == -np.inf and out[2] == 0.0 and out[3] == np.inf and np.isnan(out[4])


def test_gzip_member_round_trips_any_bytes():
    """gsgph_gzip_member (one dynamic-Huffman block of literals, built directly; zlib's Z_HUFFMAN_ONLY when a code would
    exceed 15 bits): whatever the bytes, gzip reads the member back, members concatenate, and digit text comes out at
    the entropy-coding ratio."""
    import gzip
    from go_gsgp_b200.host import gzip_member, format_semantic
    rng = np.random.default_rng(12)
    text = (format_semantic(rng.normal(size=200_000)) + "\n").encode()
    fib = b"".join(bytes([i]) * f for i, f in enumerate([1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584,
                                                          4181, 6765, 10946, 17711, 28657, 46368, 75025]))   # depth 24 > 15: fallback
    cases = [text, b"", b"7", b"aaaaaaaaaaaaaaaaaaaaaaaa", bytes(range(256)) * 5, rng.integers(0, 256, 100_000, dtype=np.uint8).tobytes(),
             b"0123456789" * 1000 + b"\x00", fib, text[:3], text[:4], text[:5]]
    members = [gzip_member(c) for c in cases]
    for c, m in zip(cases, members):
        assert m[:3] == b"\x1f\x8b\x08"
        assert gzip.decompress(m) == c
    assert gzip.decompress(b"".join(members)) == b"".join(cases)              # a multi-member file reads as one stream
    assert len(members[0]) < 0.48 * len(text)
