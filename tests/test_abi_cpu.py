"""CPU-only checks of the boundary: libgenztok.so loads, exports every symbol include/genztok.h declares,
its host-side loader reproduces the reference's dict semantics, and compute calls on a device-less handle
fail loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import tempfile

import pytest

from genz_tokenize_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L.build()
    return L.load()


def host_tok(vocab=None, codes=None, **kw):
    from genz_tokenize_b200 import Tokenize
    if vocab is None:
        return Tokenize(devices=[], **kw)
    return Tokenize.fromFile(vocab, codes, devices=[], **kw)


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "genztok.h")).read()
    declared = set(re.findall(r"\b(genztok_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.genztok_version()


def test_bundled_tables_host_only(golden):
    t = host_tok()
    assert t.vocab_size() == golden["meta"]["vocab_size"] == 48423
    assert t._special_ids() == [0, 1, 2, 3, 4]
    assert len(t.bpe_ranks) == golden["meta"]["n_ranks"] == 50001
    assert t.bpe_ranks[('#version:', '0.2')] == 0 and t.bpe_ranks[('n', 'g</w>')] == 1
    assert t.encoder['sinh_viên'] == 770 and t.decoder[770] == 'sinh_viên'
    assert t._lib.genztok_device_count(t._h) == 0


def test_loader_quirks_match_reference(golden, oracle):
    with tempfile.TemporaryDirectory() as td:
        for i, Lq in enumerate(golden["loaders"]):
            vp, mp = os.path.join(td, "v.txt"), os.path.join(td, "m.codes")
            open(vp, "wb").write(Lq["vocab"].encode("utf-8"))
            open(mp, "wb").write(Lq["merges"].encode("utf-8"))
            t = host_tok(vp, mp)
            assert t.encoder == Lq["encoder"], i
            assert {str(k): v for k, v in t.decoder.items()} == Lq["decoder"], i    # pins dict insertion order too
            assert t.vocab_size() == Lq["vocab_size"], i
            assert t._special_ids() == Lq["special_ids"], i
            assert len(t.bpe_ranks) == Lq["n_ranks"], i
            for a, b, r in Lq["ranks2"]:
                assert t.bpe_ranks[(a, b)] == r
                assert t._lib.genztok_rank_get(t._h, a.encode(), len(a.encode()), b.encode(), len(b.encode())) == r
            for k in range(-1, Lq["vocab_size"] + 3):
                key, n = C.POINTER(C.c_uint8)(), C.c_int64()
                t._lib.genztok_decoder_get(t._h, k, C.byref(key), C.byref(n))
                got = C.string_at(key, n.value).decode() if key else None
                assert got == Lq["decoder"].get(str(k)), (i, k)


def test_custom_specials_host_only(golden):
    cs = golden["custom_specials"]
    t = host_tok(pad_token=cs["specials"][0], bos_token=cs["specials"][1], eos_token=cs["specials"][2],
                 mask_token=cs["specials"][3], unk_token=cs["specials"][4])
    assert t.vocab_size() == cs["vocab_size"]
    assert t.encoder["[PAD]"] == 0 and t.encoder["[UNK]"] == 4


def test_errors_like_the_reference():
    from genz_tokenize_b200 import Tokenize
    with pytest.raises(FileNotFoundError):
        Tokenize.fromFile("/nonexistent/vocab.txt", "/nonexistent/bpe.codes", devices=[])
    with tempfile.TemporaryDirectory() as td:
        vp, mp = os.path.join(td, "v.txt"), os.path.join(td, "m.codes")
        open(vp, "wb").write(b"ok 1\n\xff\xfe 2\n")
        open(mp, "wb").write(b"a b\n")
        with pytest.raises(UnicodeDecodeError):
            Tokenize.fromFile(vp, mp, devices=[])


def test_no_cpu_fallback():
    from genz_tokenize_b200 import GenztokError
    t = host_tok()
    with pytest.raises(GenztokError, match="no CPU path"):
        t("xin chào")
    with pytest.raises(GenztokError, match="no CPU path"):
        t.decode([1, 770, 2])
    with pytest.raises(GenztokError, match="no CPU path"):
        t.bpe("hello")
    with pytest.raises(TypeError):
        t(["xin chào"])          # the reference has no list API: re.findall raises TypeError (tokenize.py:106)
    with pytest.raises(TypeError):
        t(b"xin")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genz_tokenize_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f == "workload.py", os.path.join(dirpath, f)
