"""CPU tests of the C++ host side (carpedeam_b200/host/formats.cpp): the text formats between the stages
(QueryMatcher::prefilterHitToBuffer / parsePrefilterHits, Matcher::resultToBuffer / parseAlignmentRecord) against the
reference's own pref / aln databases (tests/golden)."""
import ctypes as C

import numpy as np
import pytest

from support import load_golden


@pytest.fixture(scope="module")
def host():
    import carpedeam_b200 as cd
    cd.build()
    H = C.CDLL(cd.HOST_LIB_PATH)
    H.cdh_pref_roundtrip.restype = C.c_size_t
    H.cdh_pref_roundtrip.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    H.cdh_format_aln_line.restype = C.c_size_t
    H.cdh_format_aln_line.argtypes = [C.c_char_p, C.c_uint32, C.c_int32, C.c_float, C.c_double] + [C.c_int32] * 6
    H.cdh_parse_aln_entry.restype = C.c_size_t
    H.cdh_parse_aln_entry.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    return H


@pytest.mark.parametrize("name", ["pref_0", "pref_asm_5", "pref_asm_9"])
def test_pref_entries_round_trip(host, name):
    """parse (diagonal read back as unsigned short) + format (written as signed short) reproduces every entry byte for byte"""
    keys, payloads, _ = load_golden(name)
    out = C.create_string_buffer(1 << 22)
    neg = 0
    for p in payloads:
        n = host.cdh_pref_roundtrip(p, out, len(out))
        assert out.raw[:n] == p
        neg += p.count(b"\t-")
    assert neg > 100      # negative scores (reverse strand) and negative diagonals occur


@pytest.mark.parametrize("name", ["aln_0", "aln_asm_5"])
def test_aln_lines_format_and_parse(host, name):
    """every alignment line of the reference re-formatted by the host formatter from its own fields (seqId quirks: "1.00",
    truncation to three decimals, %.3E E-values), and the parser returns the fields ancient_correction reads"""
    keys, payloads, _ = load_golden(name)
    buf = C.create_string_buffer(256)
    checked = 0
    for p in payloads[:4000]:
        lines = [l for l in p.split(b"\n") if l]
        for i, line in enumerate(lines):
            f = line.split(b"\t")
            sid_txt = f[2]
            sid = 1.0 if sid_txt == b"1.00" else float(sid_txt) + 0.0004      # any value that truncates to the printed decimals
            n = host.cdh_format_aln_line(buf, int(f[0]), int(f[1]), sid, float(f[3]), *[int(x) for x in f[4:10]])
            assert buf.raw[:n] == line + b"\n", (buf.raw[:n], line)
            if i == len(lines) - 1 and i < 3:
                t = C.c_uint32(); s = C.c_float(); c4 = (C.c_int32 * 4)()
                cnt = host.cdh_parse_aln_entry(p, i, C.byref(t), C.byref(s), c4)
                assert cnt == len(lines) and t.value == int(f[0]) and list(c4) == [int(f[4]), int(f[5]), int(f[7]), int(f[8])]
                assert s.value == np.float32(float(sid_txt))
            checked += 1
    assert checked > 5000


def _write_db(path, keys, payloads, exts, dbtype, split=1, shuffle=False):
    """an MMseqs2 database as the reference's DBWriter leaves it: data file(s) (`split` > 1: per-thread files db.0, db.1, ...
    with offsets over their concatenation), 4-column index, 4-byte dbtype"""
    import struct
    order = list(range(len(keys)))
    if shuffle:
        np.random.default_rng(0).shuffle(order)
    blobs = [b""] * split
    rows = {}
    per = (len(order) + split - 1) // split
    offset_base = [0] * split
    for j, i in enumerate(order):
        f = min(j // per, split - 1)
        rows[keys[i]] = (f, len(blobs[f]), len(payloads[i]) + 1, exts[i])
        blobs[f] += payloads[i] + b"\0"
    for f in range(1, split):
        offset_base[f] = offset_base[f - 1] + len(blobs[f - 1])
    for f in range(split):
        with open(path if split == 1 else "%s.%d" % (path, f), "wb") as fh:
            fh.write(blobs[f])
    with open(path + ".index", "w") as fh:
        for k in (sorted(rows) if not shuffle else list(rows)):
            f, off, ln, e = rows[k]
            fh.write("%d\t%d\t%d\t%d\n" % (k, offset_base[f] + off, ln, e))
    with open(path + ".dbtype", "wb") as fh:
        fh.write(struct.pack("<i", dbtype))


def _read_db(path):
    data = open(path, "rb").read()
    out = {}
    for line in open(path + ".index"):
        k, off, ln, e = [int(x) for x in line.split()]
        out[k] = (data[off:off + ln - 1], e)
    return out


@pytest.mark.parametrize("split,shuffle", [(1, False), (3, False), (1, True)])
def test_db_reader_writer_round_trip(tmp_path, split, shuffle):
    """mmdb.cpp reads what the reference writes (single or per-thread data files, index in any order, CarpeDeam's ext column)
    and writes what the reference reads: same keys, payloads, ext flags and dbtype after a read -> write round trip"""
    import carpedeam_b200 as cd
    cd.build()
    H = C.CDLL(cd.HOST_LIB_PATH)
    H.cdh_db_roundtrip.restype = C.c_long
    H.cdh_db_roundtrip.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
    keys, payloads, exts = load_golden("assembly_reads_4")      # reads and ext-flagged contigs
    keys, payloads, exts = keys[:3000], payloads[:3000], exts[:3000]
    src = str(tmp_path / "in"); dst = str(tmp_path / "out")
    _write_db(src, keys, payloads, exts, 1, split=split, shuffle=shuffle)
    t = C.c_int()
    n = H.cdh_db_roundtrip(src.encode(), dst.encode(), C.byref(t))
    assert n == len(keys) and t.value == 1
    got = _read_db(dst)
    assert sorted(got) == keys
    assert any(exts) and not all(exts)
    for k, p, e in zip(keys, payloads, exts):
        assert got[k] == (p, e)
    idx_keys = [int(l.split()[0]) for l in open(dst + ".index")]
    assert idx_keys == sorted(idx_keys)          # DBWriter::close sorts the index by key


def test_fast_formatters_equal_the_c_library(host):
    """the hand-written integer / %.3E formatters and field parsers of the stage binaries (0.10 us per alignment line instead of 0.90 with
    sprintf) against snprintf / strtol / strtod on 4 million seeded values: random doubles over 120 decades, values next to decimal ties,
    short decimals, zero"""
    host.cdh_format_selftest.restype = C.c_uint64
    host.cdh_format_selftest.argtypes = [C.c_uint64, C.c_uint64]
    for seed in (1, 2):
        assert host.cdh_format_selftest(2000000, seed) == 0


@pytest.mark.parametrize("parts", [1, 2, 7, 64])
def test_db_writer_parts_closed_into_one_database(tmp_path, parts):
    """the stage binaries format their entries with one writer per host thread over contiguous blocks of queries and close them into ONE
    database (DbWriter::close_parts: payloads written part by part, index offsets shifted, no concatenated copy): same entries as a single
    writer, also with more parts than entries per part and an unordered input index"""
    import carpedeam_b200 as cd
    cd.build()
    H = C.CDLL(cd.HOST_LIB_PATH)
    H.cdh_db_roundtrip_parts.restype = C.c_long
    H.cdh_db_roundtrip_parts.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    keys, payloads, exts = load_golden("assembly_reads_4")
    keys, payloads, exts = keys[:500], payloads[:500], exts[:500]
    src = str(tmp_path / "in"); dst = str(tmp_path / "out")
    _write_db(src, keys, payloads, exts, 1, split=2, shuffle=True)
    assert H.cdh_db_roundtrip_parts(src.encode(), dst.encode(), parts) == len(keys)
    got = _read_db(dst)
    assert sorted(got) == keys and all(got[k] == (p, e) for k, p, e in zip(keys, payloads, exts))
    lines = [l.split() for l in open(dst + ".index")]
    assert [int(l[0]) for l in lines] == keys and all(len(l) == 4 for l in lines)
    # offsets are contiguous in key order (block-partitioned writers in thread order)
    assert all(int(a[1]) + int(a[2]) == int(b[1]) for a, b in zip(lines, lines[1:]))
