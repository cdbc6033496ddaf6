"""CPU tests of the packed hand-off's host side (carpedeam_b200/host/pack.cpp): 2-bit packing against a numpy restatement of the
layout, expansion of compact alignment records with the reference's E-value / bit-score formulas (checked against the oracle,
which is pinned by the reference's own alignment DBs), patch application."""
import ctypes as C

import numpy as np
import pytest

from support import SeqSet, synth_reads


def _np_pack(s):
    code = np.zeros(256, np.uint64)
    for ch, c in zip(b"ACTG", range(4)):
        code[ch] = c
    words, masks = [], []
    for i in range(s.n):
        b = np.frombuffer(s.seq(i), np.uint8)
        for w0 in range(0, len(b), 32):
            chunk = b[w0:w0 + 32]
            word = 0; xm = 0
            for j, ch in enumerate(chunk.tolist()):
                if ch == ord("N"):
                    xm |= 1 << (31 - j)
                else:
                    word |= int(code[ch]) << (62 - 2 * j)
            words.append(word); masks.append(xm)
    return np.array(words, np.uint64), np.array(masks, np.uint32)


def test_pack_seqs_layout():
    import carpedeam_b200 as cd
    cd.build()
    seqs = synth_reads(300, 5000, 9, lmin=5, lmax=97, n_frac=0.02) + [b"", b"A", b"GT", b"N" * 33, b"ACGT" * 16]
    s = SeqSet(seqs)
    words, nmask, has_n = cd.pack_seqs(s.blob, s.off, s.len)
    w2, m2 = _np_pack(s)
    assert has_n and np.array_equal(words, w2) and np.array_equal(nmask, m2)
    with pytest.raises(cd.CdgError):
        cd.pack_seqs(*[getattr(SeqSet([b"ACGTacgt"]), f) for f in ("blob", "off", "len")])
    with pytest.raises(cd.CdgError):
        cd.pack_seqs(*[getattr(SeqSet([b"ACGTR"]), f) for f in ("blob", "off", "len")])
    with pytest.raises(cd.CdgError):     # N but the caller declined the mask
        cd.pack_seqs(s.blob, s.off, s.len, with_mask=False)


def test_expand_alns_and_patches(oracle):
    import carpedeam_b200 as cd
    cd.build()
    lens = np.array([50, 60, 40, 0, 35], np.uint32)
    # query 0: implicit self + one forward hit; query 1: explicit self (score 0, start -1) + reverse hit; query 2: nothing; 3: implicit... no: empty
    count = np.array([1 | (1 << 31), 2, 0, 0, 1 << 31], np.uint32)
    tgt = np.array([1, 1, 4], np.uint32)
    score = np.array([70, 0, 61], np.uint16); idc = np.array([38, 0, 33], np.uint16)
    qs = np.array([10, -1, 34], np.int16); ds = np.array([0, -1, 2], np.int16)
    alr = np.array([40 << 1, 1 << 1, (33 << 1) | 1], np.uint16)
    keep = [count, tgt, score, idc, qs, ds, alr]
    ca = cd.AlnsCompact(5, 3, 0, count.ctypes.data_as(C.POINTER(C.c_uint32)), tgt.ctypes.data_as(C.POINTER(C.c_uint32)), 2, score.ctypes.data,
                        idc.ctypes.data, qs.ctypes.data, ds.ctypes.data, alr.ctypes.data, 12345678)
    a = cd.expand_alns(ca, lens)
    assert a["line_off"].tolist() == [0, 2, 4, 4, 4, 5]
    assert a["tgt"].tolist() == [0, 1, 1, 4, 4]
    assert a["score"].tolist() == [100, 70, 0, 61, 70]
    assert a["qs"].tolist() == [0, 10, -1, 34, 0] and a["qe"].tolist() == [49, 49, -1, 2, 34]
    assert a["ds"].tolist() == [0, 0, -1, 2, 0] and a["de"].tolist() == [49, 39, -1, 34, 34]
    assert a["seqid"].tolist() == [1.0, np.float32(38) / np.float32(40), 0.0, 1.0, 1.0]
    qlen = [50, 50, 60, 60, 35]
    for i in range(5):
        e, b = oracle.evalue(float(a["score"][i]), float(qlen[i]), 12345678.0)
        assert abs(a["evalue"][i] - e) <= 1e-12 * abs(e) and a["bits"][i] == int(b + 0.5)
    # same with 32-bit fields
    wide = [x.astype(t) for x, t in zip((score, idc, qs, ds, alr), (np.uint32, np.uint32, np.int32, np.int32, np.uint32))]
    cb = cd.AlnsCompact(5, 3, 0, ca.count, ca.target, 4, *[x.ctypes.data for x in wide], 12345678)
    b2 = cd.expand_alns(cb, lens)
    for f in a:
        assert np.array_equal(a[f], b2[f]), f
    # patches
    blob = np.frombuffer(b"ACGT\n\0GGGG\n\0", np.uint8)
    ent = np.array([(1 << 8) | ord("T"), (9 << 8) | ord("A")], np.uint64)
    p = cd.Patches(2, ent.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert cd.apply_patches(blob, p).tobytes() == b"ATGT\n\0GGGA\n\0"
    bad = np.array([(99 << 8) | ord("T")], np.uint64)
    with pytest.raises(cd.CdgError):
        cd.apply_patches(blob, cd.Patches(1, bad.ctypes.data_as(C.POINTER(C.c_uint64))))
    assert keep
