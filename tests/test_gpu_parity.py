"""GPU parity: the CUDA engine (through the C ABI) against the reference-generated golden vectors and,
on seeded synthetic batches, against the CPU oracle.  Bit-exact for every output (integer work)."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from golden_util import check_cases
from parity_util import assert_matches_oracle, pad_run_rows as _pad_run_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tok():
    from genz_tokenize_b200 import Tokenize
    return Tokenize()


def no_offset(cases):
    return [c for c in cases if not c["kw"].get("return_offset")]


def test_readme_vector(tok, golden):
    check_cases(tok, golden["readme"], "readme")
    assert tok.decode([1, 770, 2]) == "<s> sinh_viên </s>"
    out = tok("sinh_viên công_nghệ", "hello", max_len=10, padding=True, truncation=True)
    assert out["input_ids"] == [1, 770, 1444, 2, 2, 30469, 2, 0, 0, 0]
    assert out["attention_mask"] == [1, 1, 1, 1, 1, 1, 1, 0, 0, 0]
    assert out["sequence_id"] == [0, 0, 0, 0, 1, 1, 1]
    assert out["token_type_ids"] == [0, 0, 0, 0, 1, 1, 1, 0, 0, 0]
    assert tok("xin chào", text_pair="hello", max_len=8) == tok("xin chào", "hello", max_len=8)


def test_corner_calls(tok, golden):
    check_cases(tok, golden["calls"], "calls")      # includes return_offset=True cases


def test_return_offset_batch_vs_oracle(tok, oracle):
    from genz_tokenize_b200 import workload
    for seed, paired, kw in [(501, True, dict(max_len=12)), (502, False, dict()), (503, True, dict())]:
        t = workload.generate(seed, 1200, 0, 12, 0.1)
        p = workload.generate(seed + 5000, 1200, 0, 12, 0.1) if paired else None
        be = tok.encode_batch(t, p, return_offset=True, **kw)
        orc = oracle.encode_batch(t, p, return_offset=True, threads=8, **kw)
        assert np.array_equal(be["span_off"], orc["span_off"]), seed
        assert np.array_equal(be["spans"], orc["span"]), seed
        assert np.array_equal(be["input_ids"], orc["ids"]), seed
    assert tok.encode("xin chào\n các bạn", True) == ([1, 217, 30075, 1742, 4, 35, 175, 2], [(0, 0), (1, 1), (2, 4), (5, 5), (6, 6), (7, 7)])


def test_random_rows_single_calls(tok, golden):
    for blk in golden["random"][:3]:
        pairs = blk["pairs"] or [None] * len(blk["texts"])
        cases = [{"text": t, "pair": p, "kw": blk["kw"], "out": o} for t, p, o in zip(blk["texts"], pairs, blk["out"])]
        check_cases(tok, cases[:120], "random seed %d" % blk["gen"]["seed"])


def test_random_rows_batch(tok, golden):
    for blk in golden["random"]:
        be = tok.encode_batch(blk["texts"], blk["pairs"], **blk["kw"])
        for i, exp in enumerate(blk["out"]):
            if "raises" in exp:
                with pytest.raises(ValueError):
                    be.row(i)
            else:
                assert be.row(i) == exp, (blk["gen"], i, blk["texts"][i], blk["pairs"][i] if blk["pairs"] else None)


def test_encode_digests_85k_reference_rows(tok):
    # the CUDA path against the reference itself (no oracle in between): SHA-256 over every row's reference dict on 85,000
    # seeded rows -- BASELINE-shaped batches, noisy / heavily truncated (ValueError rows) / ragged / unpadded regimes
    from genz_tokenize_b200 import workload
    from golden_util import load_encode_digests, row_bytes
    for c in load_encode_digests():
        t = workload.generate(c["seed"], c["n"], c["lo"], c["hi"], c["noise"])
        p = workload.generate(c["seed"] + 1000, c["n"], c["lo"], c["hi"], c["noise"]) if c["paired"] else None
        be = tok.encode_batch(t, p, **c["kw"])
        h, errors = hashlib.sha256(), 0
        for i in range(c["n"]):
            try:
                r = be.row(i)
            except ValueError:
                h.update(row_bytes(1))
                errors += 1
                continue
            h.update(row_bytes(0, r["input_ids"], r["attention_mask"], r.get("sequence_id"), r.get("token_type_ids")))
        assert errors == c["value_errors"], c
        assert h.hexdigest() == c["sha256"], c
        # and decode() of those rows: fixed planes through the fixed-width kernels, ragged ones a thread per id
        texts = tok.decode_batch(be["input_ids"]) if be["input_ids"].ndim == 2 else tok.decode_batch(be["input_ids"], be["row_off"])
        hd = hashlib.sha256()
        ok = (be["row_status"] == 0) if "row_status" in be else np.ones(c["n"], dtype=bool)    # rows the reference raised on have no ids
        for i, s in enumerate(texts):
            if ok[i]:
                hd.update(s.encode("utf-8", "surrogatepass") + b"\n")
        assert hd.hexdigest() == c["decode_sha256"], c


def test_bpe_strings(tok, golden):
    for c in golden["bpe"]:
        assert tok.bpe(c["w"]) == c["out"], c["w"]


def test_sequence_id_state_machine(tok, golden):
    for ids, raw, tt in golden["seqid"]:           # all 5,460 sequences of the reference
        assert tok.get_sequence_id(ids) == raw, ids
        if tt == "ValueError":
            with pytest.raises(ValueError):
                tok.get_token_type(tok.get_sequence_id(ids))
        else:
            assert tok.get_token_type(tok.get_sequence_id(ids)) == tt, ids
            assert tok._sequence_id(ids, True) == tt, ids


def test_attention_mask_helper(tok):
    assert tok.get_atttention_mask([1, 5, 0, 2, 0, None, 7.0]) == [1, 1, 0, 1, 0, 1, 1]
    assert tok.get_atttention_mask([]) == []


def test_decode(tok, golden):
    for c in golden["decode"]:
        assert tok.decode(c["ids"]) == c["out"], c["ids"]
    assert tok.decode(np.array([1, 770, 2])) == "<s> sinh_viên </s>"
    assert tok.decode([1.0, 770.0, 2.5, None, "770"]) == "<s> sinh_viên <unk> <unk> <unk>"
    rows = [c["ids"] for c in golden["decode"]]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    flat = np.array([v if -2**31 <= v < 2**31 else -1 for r in rows for v in r], dtype=np.int32)
    assert tok.decode_batch(flat, off) == [c["out"] for c in golden["decode"]]


def test_loader_quirks(golden):
    from genz_tokenize_b200 import Tokenize
    with tempfile.TemporaryDirectory() as td:
        for i, L in enumerate(golden["loaders"]):
            vp, mp = os.path.join(td, "v.txt"), os.path.join(td, "m.codes")
            open(vp, "wb").write(L["vocab"].encode("utf-8"))
            open(mp, "wb").write(L["merges"].encode("utf-8"))
            t = Tokenize.fromFile(vp, mp)
            t.set_option("max_chunk_bytes", 1 << 16)
            assert t.vocab_size() == L["vocab_size"]
            check_cases(t, L["calls"], "loader %d" % i)
            for d in L["decode"]:
                assert t.decode(d["ids"]) == d["out"], (i, d["ids"])


def test_custom_specials(golden):
    from genz_tokenize_b200 import Tokenize
    cs = golden["custom_specials"]
    t = Tokenize(*cs["specials"])
    check_cases(t, cs["calls"], "custom specials")
    for d in cs["decode"]:
        assert t.decode(d["ids"]) == d["out"]


def test_bpe_digest_all_vocab(tok, golden, oracle):
    # every vocab word and merge concatenation, one word per document, against the oracle (which is pinned to the
    # reference's digest in test_oracle_golden.py)
    words = [w[:-2] if w.endswith("@@") else w for w in tok.encoder.keys()]
    words += ["".join(k).replace("</w>", "") for k in tok.bpe_ranks.keys()]
    words = [w for w in words if w and not any(c.isspace() for c in w)]
    assert len(words) == golden["bpe_digest"]["n"]
    be = tok.encode_batch(words)
    orc = oracle.encode_batch(words, threads=8)
    assert_matches_oracle(be, orc, what="all vocab words")


CONFIGS = [
    # seed, n, lo, hi, noise, paired, kw
    (101, 3000, 3, 13, 0.0, False, dict(max_len=128)),
    (102, 3000, 3, 13, 0.02, True, dict(max_len=256)),
    (103, 2000, 0, 9, 0.2, True, dict(max_len=16)),
    (104, 2000, 0, 9, 0.2, True, dict(max_len=7)),
    (105, 2000, 0, 9, 0.2, False, dict(max_len=5)),
    (106, 1500, 0, 9, 0.2, True, dict()),
    (107, 1500, 0, 9, 0.2, True, dict(max_len=12, padding=False)),
    (108, 1500, 0, 9, 0.2, True, dict(max_len=12, truncation=False)),
    (109, 1500, 0, 9, 0.2, True, dict(max_len=0)),
    (110, 1500, 0, 9, 0.2, True, dict(max_len=-2)),
    (111, 1000, 0, 9, 0.2, False, dict(max_len=1)),
    (112, 1000, 0, 9, 0.2, True, dict(max_len=2)),
    (113, 600, 30, 120, 0.05, True, dict(max_len=100)),
    (114, 300, 200, 900, 0.02, True, dict(max_len=512)),
    (115, 1000, 3, 13, 0.05, True, dict(max_len=33)),      # width not a multiple of 16: scalar store path
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "seed%d" % c[0])
def test_synthetic_vs_oracle(tok, oracle, cfg):
    from genz_tokenize_b200 import workload
    seed, n, lo, hi, noise, paired, kw = cfg
    t = workload.generate(seed, n, lo, hi, noise)
    p = workload.generate(seed + 5000, n, lo, hi, noise) if paired else None
    be = tok.encode_batch(t, p, **kw)
    orc = oracle.encode_batch(t, p, threads=8, **kw)
    assert_matches_oracle(be, orc, what=str(cfg))


@pytest.mark.parametrize("group", [1, 2, 3, 8, 17, 32])
def test_every_tile_size_and_small_chunks(oracle, group):
    # force the number of documents per warp tile, tiny chunks (many launches, cache resets)
    from genz_tokenize_b200 import Tokenize, workload
    tok = Tokenize()
    tok.set_option("max_chunk_bytes", 1 << 15)
    tok.set_option("chunk_rows", 257)
    tok.set_option("group", group)
    tok.set_option("wide_rows", group % 2)      # odd tile sizes also exercise the int32 row staging
    for seed, paired, kw in [(201, True, dict(max_len=24)), (202, False, dict(max_len=64)), (203, True, dict())]:
        t = workload.generate(seed, 1500, 0, 20, 0.1)
        p = workload.generate(seed + 5000, 1500, 0, 20, 0.1) if paired else None
        be = tok.encode_batch(t, p, **kw)
        orc = oracle.encode_batch(t, p, threads=8, **kw)
        assert_matches_oracle(be, orc, what="group %d seed %d" % (group, seed))


@pytest.mark.parametrize("variant", [("fused", 1, 0, 0), ("fused", 0, 16, 0), ("fused", 0, 32, 3), ("fused", 0, 48, 8), ("fused", 0, 64, 0), ("fused", 0, 256, 5),
                                     ("fused", 0, 0, 32), ("flat", 0, 0, 8), ("flat", 0, 16, 1), ("flat", 0, 32, 5), ("flat", 0, 64, 16), ("flat", 0, 256, 4), ("flat", 0, 32, 32)],
                         ids=lambda v: "%s-no_tma%d-cols%d-rows%d" % v)
def test_fixed_layout_pipelines(oracle, variant):
    # the fixed planes through every pipeline: the byte-parallel one (k_flat_words + k_flat_rows) and the fused row kernel with
    # TMA or store-instruction write-out; every staged width (rows that do not fit take the generic second pass), tile
    # sizes, partial last tiles, multi-token words, special ids inside the text -- all against the oracle
    from genz_tokenize_b200 import Tokenize, workload
    pipeline, no_tma, cols, rows = variant
    tok = Tokenize()
    tok.set_option("no_flat", int(pipeline == "fused"))
    tok.set_option("no_tma", no_tma)
    tok.set_option("tma_columns", cols)
    if pipeline == "fused":
        tok.set_option("group", rows)
    else:
        tok.set_option("flat_rows", rows)
    tok.set_profiling(True)
    for seed, n, lo, hi, noise, paired, kw in [(701, 3001, 3, 13, 0.02, False, dict(max_len=128)), (702, 2999, 3, 13, 0.02, True, dict(max_len=256)),
                                               (703, 2000, 0, 9, 0.2, True, dict(max_len=16)), (704, 1000, 10, 60, 0.05, True, dict(max_len=96)),
                                               (705, 1500, 0, 40, 0.1, False, dict(max_len=32)), (706, 700, 100, 300, 0.02, False, dict(max_len=512)),
                                               (707, 5000, 0, 3, 0.5, True, dict(max_len=16)), (708, 20000, 3, 13, 0.0, False, dict(max_len=64))]:
        t = workload.generate(seed, n, lo, hi, noise)
        p = workload.generate(seed + 5000, n, lo, hi, noise) if paired else None
        be = tok.encode_batch(t, p, **kw)
        orc = oracle.encode_batch(t, p, threads=8, **kw)
        assert_matches_oracle(be, orc, what="variant %r seed %d" % (variant, seed))
    kernels = tok.profile_report()
    want = "k_flat_rows" if pipeline == "flat" else ("k_rows_fixed" if no_tma or not cols else "k_rows_fixed_tma")
    assert want in kernels, sorted(kernels)
    if pipeline == "fused":
        assert "k_flat_rows" not in kernels


def test_flat_pipeline_text_edges(oracle):
    # byte-parallel pipeline on the shapes that stress its block / granule boundaries: exotic whitespace next to block ends,
    # words across 8 KiB blocks, very long words, runs of empty documents, documents of one byte, offsets that do not start at 0
    from genz_tokenize_b200 import Tokenize
    from oracle.oracle import pack_strings
    rng = np.random.default_rng(11)
    ws = [" ", "\n", "\t", "\r\n", "\x1c", "\x85", "\xa0", "\u1680", "\u2003", "\u2028", "\u202f", "\u205f", "\u3000", "  ", " \n "]
    near = ["\u200b", "\ufeff", "\u180e", "\xc2", "\u2040", "\u3001", "\xe1", "\u00e2\u0080"]
    words = ["sinh_viên", "công_nghệ", "hello", "xin", "chào", "a", "b" * 23, "c" * 24, "d" * 25, "e" * 40, "ế" * 9, "thế_giới", "x" * 300, "\x00", "é"]
    docs = []
    for i in range(6000):
        k = int(rng.integers(0, 12))
        parts = []
        for _ in range(k):
            parts.append(words[int(rng.integers(0, len(words)))])
            parts.append(ws[int(rng.integers(0, len(ws)))] if rng.random() < 0.9 else near[int(rng.integers(0, len(near)))])
        docs.append("".join(parts))
    docs += [""] * 50 + ["a"] * 300 + ["\n"] * 20 + ["y" * 9000, "z" * 70000] + ["\u2003"] * 10
    order = rng.permutation(len(docs))
    docs = [docs[int(i)] for i in order]
    tok = Tokenize()
    tok.set_profiling(True)
    packed = pack_strings(docs)
    for kw in (dict(max_len=64), dict(max_len=16), dict(max_len=256)):
        be = tok.encode_batch(packed, **kw)
        orc = oracle.encode_batch(packed, None, threads=8, **kw)
        assert_matches_oracle(be, orc, what="flat edges %r" % (kw,))
        n_docs = len(packed[1]) - 1
        lens = np.diff(packed[1])
        ridx = np.arange(n_docs)[::-1]                                   # the same documents in reverse order as side B
        pieces = [packed[0][packed[1][i]:packed[1][i + 1]] for i in ridx]
        other = (np.concatenate(pieces) if pieces else packed[0][:0], np.concatenate([[0], np.cumsum(lens[ridx])]).astype(np.int64))
        be = tok.encode_batch(packed, other, **kw)
        orc = oracle.encode_batch(packed, other, threads=8, **kw)
        assert_matches_oracle(be, orc, what="flat edges pairs %r" % (kw,))
    assert "k_flat_rows" in tok.profile_report()
    # far more rows than text (empty documents): the pad columns are then written by k_flat_rows itself
    sparse = pack_strings([""] * 70000 + ["xin chào", "a"] * 5 + [""] * 3000)
    for kw in (dict(max_len=32), dict(max_len=128)):
        be = tok.encode_batch(sparse, **kw)
        orc = oracle.encode_batch(sparse, None, threads=8, **kw)
        assert_matches_oracle(be, orc, what="flat sparse %r" % (kw,))
        be = tok.encode_batch(sparse, sparse, **kw)
        orc = oracle.encode_batch(sparse, sparse, threads=8, **kw)
        assert_matches_oracle(be, orc, what="flat sparse pairs %r" % (kw,))
    # a chunk that starts in the middle of the caller's buffer (absolute offsets, unaligned base)
    tok2 = Tokenize()
    tok2.set_option("chunk_rows", 777)
    be = tok2.encode_batch(packed, max_len=32)
    orc = oracle.encode_batch(packed, None, threads=8, max_len=32)
    assert_matches_oracle(be, orc, what="flat edges chunked")


def test_word_cache_follows_the_chunk_size(oracle):
    # the word cache is sized for the worst case of one chunk: a handle that only sees small batches must not take the 3 GiB that
    # max_chunk_bytes = 64 MiB costs; a larger batch regrows (and empties) it, a smaller one afterwards reuses it; results unchanged
    import torch
    from genz_tokenize_b200 import Tokenize, workload
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    t = Tokenize(devices=[0])
    small = workload.generate(41, 2000, 0, 13, 0.02)
    ref_small = oracle.encode_batch(small, None, threads=8, max_len=32)
    assert_matches_oracle(t.encode_batch(small, max_len=32), ref_small, what="small batch, small cache")
    torch.cuda.synchronize()
    used_small = free0 - torch.cuda.mem_get_info(0)[0]
    assert used_small < (1 << 30), used_small                            # (tables, work arrays, lazily loaded kernels and a 32 MiB slot table)
    big = workload.generate(42, 300000, 3, 13, 0.01)                      # ~15 MB: the cache regrows
    assert_matches_oracle(t.encode_batch(big, max_len=32), oracle.encode_batch(big, None, threads=8, max_len=32), what="large batch, regrown cache")
    torch.cuda.synchronize()
    used_big = free0 - torch.cuda.mem_get_info(0)[0]
    assert used_big > used_small
    assert_matches_oracle(t.encode_batch(small, max_len=32), ref_small, what="small batch again")
    assert_matches_oracle(t.encode_batch(small), oracle.encode_batch(small, None, threads=8), what="small batch, ragged")
    assert t.check_errors() == 0
    fixed = Tokenize(devices=[0])
    fixed.set_option("fixed_cache", 1)                                    # the old behaviour, on request
    assert_matches_oracle(fixed.encode_batch(small, max_len=32), ref_small, what="small batch, full-size cache")


def test_decode_roundtrip_batch(tok, oracle):
    from genz_tokenize_b200 import workload
    t = workload.generate(301, 4000, 3, 13, 0.02)
    p = workload.generate(5301, 4000, 3, 13, 0.02)
    be = tok.encode_batch(t, p, max_len=64)
    texts = tok.decode_batch(be["input_ids"])
    ref = oracle.decode_batch(be["input_ids"].reshape(-1), np.arange(0, 4000 * 64 + 1, 64, dtype=np.int64), threads=8)
    assert texts == ref
    rag = tok.encode_batch(t, p)
    assert tok.decode_batch(rag["input_ids"], rag["row_off"]) == oracle.decode_batch(rag["input_ids"], rag["row_off"], threads=8)
    # more rows than one scan tile: the multi-block offset scan
    t2 = workload.generate(302, 30000, 0, 9, 0.02)
    big = tok.encode_batch(t2, max_len=16)
    assert tok.decode_batch(big["input_ids"]) == oracle.decode_batch(big["input_ids"].reshape(-1), np.arange(0, 30000 * 16 + 1, 16, dtype=np.int64), threads=8)


def test_decode_pad_runs(tok, oracle):
    # the write pass sends the trailing run of pad ids and the last piece straight to global memory: every width around the
    # 16-byte units and the 32-id batches, every shape of row, against the oracle
    rng = np.random.default_rng(77)
    exotic = [-1, -2**31, 2**31 - 1, 48423, 48422, 10**6, 3, 4]
    from genz_tokenize_b200 import Tokenize
    any_rows = Tokenize()
    any_rows.set_option("no_fixed_decode", 1)            # widths that are a multiple of 4 through the warp-per-row kernels too
    any_rows.set_option("no_token_decode", 1)            # and ragged rows (a thread per id by default)
    coop, lanes = Tokenize(), Tokenize()                 # both write kernels of the fixed-width family, whatever the average lead
    coop.set_option("decode_write", 1)
    lanes.set_option("decode_write", 2)
    for width in [1, 2, 3, 4, 7, 8, 9, 10, 11, 12, 16, 17, 31, 32, 33, 40, 64, 100, 124, 128, 132, 252, 255, 256, 257, 260, 300, 384, 516]:
        ids = _pad_run_rows(rng, 300, width, 0, 48423, exotic)
        ref = oracle.decode_batch(ids.reshape(-1), np.arange(0, 300 * width + 1, width, dtype=np.int64), threads=8)
        assert tok.decode_batch(ids) == ref, width
        assert any_rows.decode_batch(ids) == ref, width
        assert coop.decode_batch(ids) == ref, width
        assert lanes.decode_batch(ids) == ref, width
    # encoder-shaped rows (short leads, long pad runs: the rows the lane-per-row kernel is for) at every alignment of the text
    for width, lo, hi in [(128, 1, 14), (256, 10, 45), (64, 1, 40), (48, 0, 3)]:
        ids = np.zeros((500, width), dtype=np.int32)
        for r in range(500):
            k = int(rng.integers(lo, hi + 1))
            ids[r, :k] = rng.integers(1, 48423, k)
        ref = oracle.decode_batch(ids.reshape(-1), np.arange(0, 500 * width + 1, width, dtype=np.int64), threads=8)
        assert coop.decode_batch(ids) == ref and lanes.decode_batch(ids) == ref and tok.decode_batch(ids) == ref, width
    for n_rows in [1, 31, 32, 33, 64, 65]:                # tiles of 32 rows, whole and partial
        ids = _pad_run_rows(rng, n_rows, 128, 0, 48423, exotic)
        assert tok.decode_batch(ids) == oracle.decode_batch(ids.reshape(-1), np.arange(0, n_rows * 128 + 1, 128, dtype=np.int64)), n_rows
    # ragged rows of the same shapes (rows start at any alignment: the scalar id loads of the length pass)
    rows = [_pad_run_rows(rng, 1, int(w), 0, 48423, exotic)[0] for w in rng.integers(0, 200, 700)]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    flat = np.concatenate(rows).astype(np.int32)
    ref = oracle.decode_batch(flat, off, threads=8)
    assert tok.decode_batch(flat, off) == ref
    assert any_rows.decode_batch(flat, off) == ref
    sub = off[200:501]                                   # a batch whose offsets do not start at 0
    assert tok.decode_batch(flat, sub) == ref[200:500]
    assert any_rows.decode_batch(flat, sub) == ref[200:500]
    empty = np.zeros(6, dtype=np.int64)                  # nothing but empty rows
    assert tok.decode_batch(flat, empty) == [""] * 5
    # one very long row, mostly padding
    long_row = np.zeros((3, 70001), dtype=np.int32)
    long_row[:, :900] = rng.integers(0, 48423, (3, 900))
    long_row[1, -1] = 770
    assert tok.decode_batch(long_row) == oracle.decode_batch(long_row.reshape(-1), np.arange(0, 3 * 70001 + 1, 70001, dtype=np.int64))
    long4 = np.ascontiguousarray(long_row[:, :70000])
    long4[2, 69990:] = 5
    assert tok.decode_batch(long4) == oracle.decode_batch(long4.reshape(-1), np.arange(0, 3 * 70000 + 1, 70000, dtype=np.int64))


def test_decode_device_into_a_ring(tok, oracle):
    # genztok_decode_device_into: sizes and text in one go into the caller's buffer, the write kernel picked on the device by the
    # batch's average lead (all three families), the capacity checked by the kernels themselves
    import torch
    from genz_tokenize_b200 import Tokenize
    rng = np.random.default_rng(91)
    dev = torch.device("cuda:0")
    ring = torch.empty((1 << 22,), dtype=torch.uint8, device=dev)
    def texts(txt, off):
        b = txt.cpu().numpy().tobytes(); o = off.cpu().numpy()
        return [b[o[i]:o[i + 1]].decode("utf-8") for i in range(len(o) - 1)]
    for width, lo, hi in [(128, 1, 14), (256, 20, 60), (256, 70, 200), (64, 1, 60), (12, 0, 3)]:   # -> 256-byte junctions, 512-byte, warp per row, mixed
        ids = np.zeros((700, width), dtype=np.int32)
        for r in range(700):
            k = min(int(rng.integers(lo, hi + 1)), width)
            ids[r, :k] = rng.integers(1, 48423, k)
        ref = oracle.decode_batch(ids.reshape(-1), np.arange(0, 700 * width + 1, width, dtype=np.int64), threads=8)
        d_ids = torch.from_numpy(ids).to(dev)
        ring.fill_(255)
        txt, off = tok.decode_device(d_ids, out=ring)
        assert txt.data_ptr() == ring.data_ptr() and texts(txt, off) == ref, width
        assert int(ring[txt.numel():txt.numel() + 64].min()) == 255                     # nothing behind the text
        out2, off2 = tok.decode_device(d_ids, out=ring, sync=False)                     # no host read at all
        torch.cuda.synchronize()
        assert out2.data_ptr() == ring.data_ptr() and texts(out2[:int(off2[-1])], off2) == ref, width
        small = torch.full((int(off[-1]) - 1 + 16,), 7, dtype=torch.uint8, device=dev)[:int(off[-1]) - 1]   # one byte short: nothing may be written
        txt3, off3 = tok.decode_device(d_ids, out=small)
        assert txt3.data_ptr() != small.data_ptr() and texts(txt3, off3) == ref and int(small.max()) == 7 and int(small.min()) == 7, width
    # ragged rows (a thread per id) and any-width rows (a warp per row) through the same entry point
    rows = [_pad_run_rows(rng, 1, int(w), 0, 48423, [3, 4])[0] for w in rng.integers(0, 90, 400)]
    o = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    flat = np.concatenate(rows).astype(np.int32)
    ref = oracle.decode_batch(flat, o, threads=8)
    txt, off = tok.decode_device(torch.from_numpy(flat).to(dev), torch.from_numpy(o).to(dev), out=ring)
    assert txt.data_ptr() == ring.data_ptr() and texts(txt, off) == ref
    odd = _pad_run_rows(rng, 300, 37, 0, 48423, [3, 4])
    txt, off = tok.decode_device(torch.from_numpy(odd).to(dev), out=ring)
    assert texts(txt, off) == oracle.decode_batch(odd.reshape(-1), np.arange(0, 300 * 37 + 1, 37, dtype=np.int64), threads=8)
    with pytest.raises(ValueError):
        tok.decode_device(torch.from_numpy(odd).to(dev), sync=False)
    assert tok.check_errors() == 0


def test_decode_pad_runs_custom_pad_tokens():
    # pad texts of other periods: "[P] " (4 bytes, divides the unit), "\x7f" ("\x7f@@ " with the marker removed: 1 byte), "Ω" (2),
    # "a~b " (4), "ặ~ " (5), "abcdefg " (8), "abcdefgh " / "<padtok> " (9 bytes: no periodic shortcut), and strings that
    # vocab.txt holds too ("ab", "p@@": the pad id moves)
    from genz_tokenize_b200 import Tokenize
    from oracle.oracle import Oracle
    rng = np.random.default_rng(78)
    for pad_tok in ["[P]", "\x7f@@", "Ω@@", "zq@@", "a~b", "ặ~", "abcdefg", "abcdefgh", "<padtok>", "ab", "p@@"]:
        sp = [pad_tok, "<s>", "</s>", "<mask>", "<unk>"]
        t, o = Tokenize(*sp), Oracle(specials=sp)
        pad = t.encoder[pad_tok]                      # 0 unless vocab.txt holds the same string (then it moved: SURVEY A.6)
        for width in [5, 16, 37, 128, 260]:
            ids = _pad_run_rows(rng, 200, width, pad, t.vocab_size(), [-1, 10**6, 1, 2, 0])
            assert t.decode_batch(ids) == o.decode_batch(ids.reshape(-1), np.arange(0, 200 * width + 1, width, dtype=np.int64)), (pad_tok, width)


def test_decode_pad_run_vectors_from_the_reference():
    # the same shapes against what the reference's decode() printed (tests/golden/decode_v2.json.gz), through all three
    # kernel families: fixed-width (default), warp per row, and -- as ragged rows -- a thread per id
    from genz_tokenize_b200 import Tokenize
    from golden_util import load_decode_golden
    toks = {}
    for b in load_decode_golden():
        if b["pad_token"] not in toks:
            sp = [b["pad_token"], "<s>", "</s>", "<mask>", "<unk>"]
            t, a, l2 = Tokenize(*sp), Tokenize(*sp), Tokenize(*sp)
            a.set_option("no_fixed_decode", 1)
            t.set_option("decode_write", 1)
            l2.set_option("decode_write", 2)
            assert t.encoder[b["pad_token"]] == b["pad_id"]
            toks[b["pad_token"]] = (t, a, l2)
        t, a, l2 = toks[b["pad_token"]]
        ids = np.array(b["ids"], dtype=np.int64).astype(np.int32)
        n, w = ids.shape
        assert t.decode_batch(ids) == b["out"], (b["pad_token"], w)
        assert a.decode_batch(ids) == b["out"], (b["pad_token"], w)
        assert l2.decode_batch(ids) == b["out"], (b["pad_token"], w)
        assert t.decode_batch(ids.reshape(-1), np.arange(0, n * w + 1, w, dtype=np.int64)) == b["out"], (b["pad_token"], w)


def test_decode_device_interleaved_batches(tok):
    # length pass of A, length pass of B, write pass of A: the per-row description of A was overwritten and is redone
    import ctypes as C
    import torch
    rng = np.random.default_rng(79)
    dev = torch.device("cuda:0")
    a = _pad_run_rows(rng, 500, 64, 0, 48423, [-1])
    b = _pad_run_rows(rng, 500, 64, 0, 48423, [-1])
    da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    oa, ob_ = torch.empty(501, dtype=torch.int64, device=dev), torch.empty(501, dtype=torch.int64, device=dev)
    ta, tb = C.c_int64(), C.c_int64()
    st = tok._torch_stream(dev)
    lib, h = tok._lib, tok._h
    assert lib.genztok_decode_device(h, 0, da.data_ptr(), None, 500, 64, oa.data_ptr(), None, C.byref(ta), st) == 0
    assert lib.genztok_decode_device(h, 0, db.data_ptr(), None, 500, 64, ob_.data_ptr(), None, C.byref(tb), st) == 0
    xa = torch.zeros(ta.value + 16, dtype=torch.uint8, device=dev)
    xb = torch.zeros(tb.value + 16, dtype=torch.uint8, device=dev)
    assert lib.genztok_decode_device(h, 0, da.data_ptr(), None, 500, 64, oa.data_ptr(), xa.data_ptr(), None, st) == 0
    assert lib.genztok_decode_device(h, 0, db.data_ptr(), None, 500, 64, ob_.data_ptr(), xb.data_ptr(), None, st) == 0
    torch.cuda.synchronize()
    # the same with ragged rows (decoded by id: the scan of A is overwritten by B's)
    ra, rb = tok.encode_batch(["xin chào các bạn", "", "sinh_viên công_nghệ\n"] * 50), tok.encode_batch(["hello", "a b c d e f g"] * 70)
    dev_t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    ia, fa, ib, fb = dev_t(ra["input_ids"]), dev_t(ra["row_off"]), dev_t(rb["input_ids"]), dev_t(rb["row_off"])
    na, nb = len(ra["row_off"]) - 1, len(rb["row_off"]) - 1
    qa, qb = torch.empty(na + 1, dtype=torch.int64, device=dev), torch.empty(nb + 1, dtype=torch.int64, device=dev)
    sa, sb = C.c_int64(), C.c_int64()
    assert lib.genztok_decode_device(h, 0, ia.data_ptr(), fa.data_ptr(), na, 0, qa.data_ptr(), None, C.byref(sa), st) == 0
    assert lib.genztok_decode_device(h, 0, ib.data_ptr(), fb.data_ptr(), nb, 0, qb.data_ptr(), None, C.byref(sb), st) == 0
    ya = torch.zeros(sa.value + 16, dtype=torch.uint8, device=dev)
    yb = torch.zeros(sb.value + 16, dtype=torch.uint8, device=dev)
    assert lib.genztok_decode_device(h, 0, ia.data_ptr(), fa.data_ptr(), na, 0, qa.data_ptr(), ya.data_ptr(), None, st) == 0
    assert lib.genztok_decode_device(h, 0, ib.data_ptr(), fb.data_ptr(), nb, 0, qb.data_ptr(), yb.data_ptr(), None, st) == 0
    torch.cuda.synchronize()
    for enc, y, q, tot in [(ra, ya, qa, sa), (rb, yb, qb, sb)]:
        raw, off = y.cpu().numpy().tobytes(), q.cpu().numpy()
        assert off[-1] == tot.value and raw[tot.value:] == bytes(16)
        assert [raw[off[i]:off[i + 1]].decode("utf-8", "surrogatepass") for i in range(len(off) - 1)] == tok.decode_batch(enc["input_ids"], enc["row_off"])
    for ids, x, o, tot in [(a, xa, oa, ta), (b, xb, ob_, tb)]:
        raw, off = x.cpu().numpy().tobytes(), o.cpu().numpy()
        assert off[-1] == tot.value
        assert [raw[off[i]:off[i + 1]].decode("utf-8", "surrogatepass") for i in range(500)] == tok.decode_batch(ids)
        assert raw[tot.value:] == bytes(16)              # nothing written behind the text


def test_device_api_matches_host_api(tok):
    import torch
    from genz_tokenize_b200 import workload
    tb, to = workload.generate(401, 5000, 3, 13, 0.02)
    pb, po = workload.generate(5401, 5000, 3, 13, 0.02)
    host = tok.encode_batch((tb, to), (pb, po), max_len=96)
    dev = torch.device("cuda:0")
    pad16 = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
    out = tok.encode_device(pad16(tb), torch.from_numpy(to).to(dev), pad16(pb), torch.from_numpy(po).to(dev), max_len=96,
                            token_type_ids=True, sequence_id=True, text_bytes=len(tb), pair_bytes=len(pb))
    torch.cuda.synchronize()
    assert np.array_equal(out["input_ids"].cpu().numpy(), host["input_ids"])
    assert np.array_equal(out["attention_mask"].cpu().numpy(), host["attention_mask"])
    assert np.array_equal(out["token_type_ids"].cpu().numpy(), host["token_type_ids"])
    assert np.array_equal(out["row_status"].cpu().numpy(), host["row_status"])
    assert np.array_equal(out["seq_len"].cpu().numpy(), host["seq_len"])
    # DLPack hand-off of a result plane
    cap = torch.utils.dlpack.to_dlpack(out["input_ids"])
    again = torch.utils.dlpack.from_dlpack(cap)
    assert again.data_ptr() == out["input_ids"].data_ptr()
    # device decode
    db, do = tok.decode_device(out["input_ids"])
    texts = tok.decode_batch(host["input_ids"])
    raw, off = db.cpu().numpy().tobytes(), do.cpu().numpy()
    assert [raw[off[i]:off[i + 1]].decode("utf-8", "surrogatepass") for i in range(5000)] == texts
    ring = torch.zeros(int(db.numel()) + 4096, dtype=torch.uint8, device=dev)      # a caller-owned text buffer, reused
    db2, do2 = tok.decode_device(out["input_ids"], out=ring)
    assert db2.data_ptr() == ring.data_ptr() and torch.equal(db2, db) and torch.equal(do2, do) and int(ring[db.numel():].sum()) == 0
    small = torch.zeros(16, dtype=torch.uint8, device=dev)                            # too small: a new tensor is returned
    db3, _ = tok.decode_device(out["input_ids"], out=small)
    assert db3.data_ptr() != small.data_ptr() and torch.equal(db3, db)


def test_file_to_device_batches_handoff(tok, oracle):
    # SURVEY.md 8 f4, both sides of the path: lines of a file -> packed batches -> planes on the GPU -> shuffled batches
    import torch
    from genz_tokenize_b200 import DataCollection, iter_line_batches, workload
    tb, to = workload.generate(811, 3000, 0, 13, 0.0)
    raw = tb.tobytes()
    lines = [raw[to[i]:to[i + 1]].decode("utf-8") for i in range(3000)]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "docs.txt")
        open(p, "wb").write(("\n".join(lines) + "\n").encode("utf-8"))
        rows = []
        for b, o in iter_line_batches(p, docs_per_batch=1024, read_bytes=50000):
            rows.append(tok.encode_batch((b, o), max_len=32)["input_ids"].copy())
    ids = np.concatenate(rows)
    ref = oracle.encode_batch((tb, to), None, threads=8, max_len=32)
    assert np.array_equal(ids.reshape(-1), ref["ids"])
    dev = torch.device("cuda:0")
    pad16 = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 16, dtype=np.uint8)])).to(dev)
    out = tok.encode_device(pad16(tb), torch.from_numpy(to).to(dev), max_len=32, text_bytes=len(tb))
    y = torch.arange(3000, device=dev)
    dc = DataCollection.from_encoding(out, y)
    got = torch.empty((3000, 32), dtype=torch.int32, device=dev)
    for feats, yy in dc.to_torch_batches(batch_size=256, seed=1):
        assert feats["input_ids"].is_cuda and feats["attention_mask"].is_cuda and yy.is_cuda
        got[yy] = feats["input_ids"]
    assert np.array_equal(got.cpu().numpy(), ids)
    # the batches are cut by the library's gather kernel, one launch for all the fields of a batch (rows of whole 16-byte
    # vectors, an int8 plane of odd width, 8-byte labels): every field equals index_select with the same permutation
    n0 = tok.launch_count()
    odd = torch.arange(3000 * 7, device=dev, dtype=torch.int32).reshape(3000, 7).to(torch.int8)
    dc2 = DataCollection(input_ids=out["input_ids"], attention_mask=out["attention_mask"], token_type_ids=odd, y=y)
    g = torch.Generator(device="cpu"); g.manual_seed(9)
    order = torch.randperm(3000, generator=g).to(dev)
    nb = 0
    for i, (feats, yy) in enumerate(dc2.to_torch_batches(batch_size=700, seed=9, tokenizer=tok)):
        idx = order[700 * i:700 * (i + 1)]
        assert torch.equal(yy, y[idx]) and torch.equal(feats["input_ids"], out["input_ids"][idx])
        assert torch.equal(feats["attention_mask"], out["attention_mask"][idx]) and torch.equal(feats["token_type_ids"], odd[idx])
        nb += 1
    assert nb == 5 and tok.launch_count() - n0 == 5
    assert [yy.tolist() for _, yy in DataCollection(input_ids=out["input_ids"], y=y).to_torch_batches(batch_size=1500, shuffle=False)] == [list(range(1500)), list(range(1500, 3000))]
    # the reader with the file read, the line scan and the host->device copy of the next batch overlapped with this batch's kernels
    from genz_tokenize_b200 import iter_device_batches
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "docs.txt")
        open(p, "wb").write(("\n".join(lines) + "\n").encode("utf-8"))
        host = list(iter_line_batches(p, docs_per_batch=700, read_bytes=30000))
        rows, k = [], 0
        for d_b, d_o, nbytes in iter_device_batches(p, docs_per_batch=700, read_bytes=30000, device=dev):
            hb, ho = host[k]; k += 1
            assert nbytes == len(hb) and d_b.numel() % 16 == 0 and np.array_equal(d_b[:nbytes].cpu().numpy(), hb) and np.array_equal(d_o.cpu().numpy(), ho)
            rows.append(tok.encode_device(d_b, d_o, max_len=32, text_bytes=nbytes)["input_ids"])
        assert k == len(host)
        assert np.array_equal(torch.cat(rows).cpu().numpy(), ids)
        it = iter_device_batches(p, docs_per_batch=100, device=dev)       # a consumer that stops early leaves no reader behind
        next(it); it.close()


def test_config2_full_size_properties(tok, oracle):
    # BASELINE.json configs[1]: 1M single sentences, max_len=128 -- first 100k rows against the oracle, the whole
    # batch through size-independent properties (row framing, mask == non-pad, decode/encode idempotence on a sample)
    from genz_tokenize_b200 import workload
    t = workload.generate(1234, 1 << 20, 3, 13, 0.0)
    be = tok.encode_batch(t, max_len=128)
    ids, mask = be["input_ids"], be["attention_mask"]
    assert ids.shape == (1 << 20, 128)
    assert (ids[:, 0] == 1).all()
    assert np.array_equal(mask, (ids != 0).astype(np.uint8))
    rl = be["row_len"]
    assert np.array_equal(rl, mask.sum(axis=1))
    assert (ids[np.arange(len(rl)), rl - 1] == 2).all()
    assert int(be["real_tokens"]) == int(mask.sum())
    n = 100000
    sub = (t[0][:t[1][n]], t[1][:n + 1])
    orc = oracle.encode_batch(sub, None, max_len=128, threads=8)
    assert np.array_equal(ids[:n].reshape(-1), orc["ids"])
    # every word of W is one token: a row has exactly (#words + 2) real tokens
    words = np.diff(np.searchsorted(np.nonzero(t[0] == 0x20)[0], t[1])) + (np.diff(t[1]) > 0)
    assert np.array_equal(rl, np.minimum(words + 2, 128))
    # checksum is independent of chunking
    tok2 = type(tok)()
    tok2.set_option("max_chunk_bytes", 1 << 22)
    be2 = tok2.encode_batch((t[0][:t[1][200000]], t[1][:200001]), max_len=128)
    assert hashlib.sha256(be2["input_ids"].tobytes()).hexdigest() == hashlib.sha256(ids[:200000].tobytes()).hexdigest()


def _custom_model(td, n_words=4000, seed=77):
    """BASELINE configs[4]: a custom vocab / bpe.codes made of concatenations of bundled words, with learned-looking merges."""
    from genz_tokenize_b200 import workload
    rng = np.random.default_rng(seed)
    wl = workload.default_wordlist()
    base = [wl.words[int(i)] for i in rng.integers(0, len(wl.words), size=n_words)]
    words = list(dict.fromkeys(base + [a + "_" + b for a, b in zip(base[::2], base[1::2])]))
    merges, seen, vocab = [], set(), {}
    for w in words:                          # merges that build every word left to right, code point by code point
        syms = list(w[:-1]) + [w[-1] + "</w>"]
        cur = syms[0]
        for s in syms[1:]:
            if (cur, s) not in seen:
                seen.add((cur, s))
                merges.append("%s %s" % (cur, s))
            cur = cur + s
            vocab.setdefault(cur.replace("</w>", "") + ("" if cur.endswith("</w>") else "@@"), 1)
        vocab[w] = 1
    rng.shuffle(merges)
    vp, mp = os.path.join(td, "vocab.txt"), os.path.join(td, "bpe.codes")
    with open(vp, "w", encoding="utf-8") as f:
        f.write("".join("%s %d\n" % (k, v) for k, v in vocab.items()))
    with open(mp, "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n" + "\n".join(merges) + "\n")
    return vp, mp, words


def test_config5_custom_vocab_long_documents(oracle):
    # fromFile with a large custom vocab / merge table, long documents with low word reuse, max_len=4096 (heavy truncation)
    from genz_tokenize_b200 import Tokenize
    from oracle.oracle import Oracle, pack_strings
    with tempfile.TemporaryDirectory() as td:
        vp, mp, words = _custom_model(td)
        tok = Tokenize.fromFile(vp, mp)
        orc = Oracle(vp, mp)
        assert tok.vocab_size() == orc.vocab_size()
        rng = np.random.default_rng(5)
        docs = []
        for i in range(48):
            k = int(rng.integers(2000, 9000)) if i % 3 else int(rng.integers(0, 400))
            ws = [words[int(j)] for j in rng.integers(0, len(words), size=k)]
            for j in rng.integers(0, max(k, 1), size=k // 50):            # unseen material: unk, partial merges
                ws[int(j)] = ws[int(j)][::-1] + "zq"
            docs.append(" ".join(ws))
        packed = pack_strings(docs)
        for kw in (dict(max_len=4096), dict(max_len=512), dict()):
            be = tok.encode_batch(packed, **kw)
            ref = orc.encode_batch(packed, None, threads=8, **kw)
            assert_matches_oracle(be, ref, what="config5 %r" % (kw,))
        pairs = pack_strings(docs[::-1])
        be = tok.encode_batch(packed, pairs, max_len=4096)
        ref = orc.encode_batch(packed, pairs, max_len=4096, threads=8)
        assert_matches_oracle(be, ref, what="config5 pairs")
        # an empty cache in front of the fused row kernel: the byte-parallel word pass fills it first (both sides); and without it
        tok.cache_reset()
        assert_matches_oracle(tok.encode_batch(packed, pairs, max_len=4096), ref, what="config5 pairs, empty cache")
        tok.cache_reset()
        tok.set_option("no_discovery", 1)
        assert_matches_oracle(tok.encode_batch(packed, pairs, max_len=4096), ref, what="config5 pairs, empty cache, no word pass")
        tok.set_option("no_discovery", 0)
        dec = tok.decode_batch(be["input_ids"][:8])
        assert dec == orc.decode_batch(be["input_ids"][:8].reshape(-1), np.arange(0, 8 * 4096 + 1, 4096, dtype=np.int64))


def test_one_handle_many_devices(oracle):
    # SURVEY.md 8e inside one process: a handle created over several GPUs shards the rows by document
    # (one host thread per GPU, no collective) and the stitched result equals the single-GPU one
    import torch
    from genz_tokenize_b200 import Tokenize, workload
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    multi = Tokenize(devices=list(range(ndev)))
    assert multi._lib.genztok_device_count(multi._h) == ndev
    t = workload.generate(601, 20000, 0, 13, 0.05)
    p = workload.generate(5601, 20000, 0, 13, 0.05)
    for kw in (dict(max_len=64), dict(), dict(max_len=9, padding=False)):
        be = multi.encode_batch(t, p, **kw)
        orc = oracle.encode_batch(t, p, threads=8, **kw)
        assert_matches_oracle(be, orc, what="multi-device %r" % (kw,))
    be = multi.encode_batch(t, None, return_offset=True)
    orc = oracle.encode_batch(t, None, return_offset=True, threads=8)
    assert np.array_equal(be["span_off"], orc["span_off"]) and np.array_equal(be["spans"], orc["span"])
    # decode shards by row the same way: parts concatenated, offsets shifted
    fix = multi.encode_batch(t, p, max_len=64)["input_ids"]
    assert multi.decode_batch(fix) == oracle.decode_batch(fix.reshape(-1), np.arange(0, 20000 * 64 + 1, 64, dtype=np.int64), threads=8)
    rag = multi.encode_batch(t, p)
    assert multi.decode_batch(rag["input_ids"], rag["row_off"]) == oracle.decode_batch(rag["input_ids"], rag["row_off"], threads=8)
    assert multi.decode_batch(fix[:3]) == oracle.decode_batch(fix[:3].reshape(-1), np.arange(0, 3 * 64 + 1, 64, dtype=np.int64))
    few = multi.encode_batch((t[0][:t[1][3]], t[1][:4]), max_len=16)       # fewer rows than 2 x devices: one device does it all
    assert np.array_equal(few["input_ids"].reshape(-1), oracle.encode_batch((t[0][:t[1][3]], t[1][:4]), None, max_len=16)["ids"])


# ---- round 2: the bench workload (BASELINE configs[2]) -- generator, digest, a full chunk against the oracle, the reference itself ----
def _device_planes(tok, ta, tb, W, dev):
    import torch
    up = lambda a: torch.from_numpy(np.concatenate([a, np.zeros((-len(a)) % 16 + 32, dtype=np.uint8)])).to(dev)
    out = tok.encode_device(up(ta[0]), torch.from_numpy(ta[1]).to(dev), up(tb[0]), torch.from_numpy(tb[1]).to(dev), max_len=W,
                            text_bytes=len(ta[0]), pair_bytes=len(tb[0]))
    torch.cuda.synchronize()
    assert tok.check_errors(dev) == 0
    return out


@pytest.mark.parametrize("noise", [0.0, 0.03])
def test_synth_device_matches_host(tok, noise):
    """csrc/synth.cuh produces the bytes workload.generate_hashed defines, for any document range and both sides."""
    import torch
    from genz_tokenize_b200 import workload
    dev = torch.device("cuda:0")
    for doc0, n, side in ((0, 5000, 0), (99_999_000, 3000, 1), (123_456_789, 1, 0)):
        hb, ho = workload.generate_hashed(1234, doc0, n, side, 3, 13, noise)
        db, do, nb = tok.synth_device(1234, doc0, n, side, 3, 13, noise, device=dev)
        assert nb == len(hb)
        assert np.array_equal(do.cpu().numpy(), ho)
        assert np.array_equal(db[:nb].cpu().numpy(), hb)


def test_plane_digest_device_matches_host_and_is_chunking_independent(tok):
    import torch
    from genz_tokenize_b200 import workload
    dev = torch.device("cuda:0")
    n, W = 6000, 64
    ta, tb = workload.generate_hashed(7, 500, n, 0, 3, 13, 0.02), workload.generate_hashed(7, 500, n, 1, 3, 13, 0.02)
    out = _device_planes(tok, ta, tb, W, dev)
    want = workload.plane_digest(out["input_ids"].cpu().numpy(), out["attention_mask"].cpu().numpy(), out["token_type_ids"].cpu().numpy(), row0=500)
    acc = torch.zeros(1, dtype=torch.int64, device=dev)
    tok.digest_device(out, 500, acc)
    assert (int(acc.item()) & 0xFFFFFFFFFFFFFFFF) == want
    acc2 = torch.zeros(1, dtype=torch.int64, device=dev)       # the same rows in three pieces, out of order
    for lo, hi in ((4000, 6000), (0, 1500), (1500, 4000)):
        tok.digest_device({k: v[lo:hi] for k, v in out.items() if v.dim() == 2}, 500 + lo, acc2)
    assert int(acc2.item()) == int(acc.item())
    out["input_ids"][1234, 3] += 1                               # ... and it sees a single changed id
    acc3 = torch.zeros(1, dtype=torch.int64, device=dev)
    tok.digest_device(out, 500, acc3)
    assert int(acc3.item()) != int(acc.item())


@pytest.mark.parametrize("doc0", [0, 37_500_000 + 7 * (1 << 20)])
def test_config3_full_chunk_vs_oracle(oracle, doc0):
    """A whole 1,048,576-pair chunk of the bench workload (the first one, and one from the middle of another shard) at max_len 256,
    three planes, device-resident: every row against the oracle."""
    import torch
    from genz_tokenize_b200 import Tokenize, workload
    from oracle.oracle import Oracle
    dev = torch.device("cuda:0")
    tok = Tokenize(devices=[0])
    tok.set_option("max_chunk_bytes", 1 << 27)
    n, W = 1 << 20, 256
    da, oa, na = tok.synth_device(1234, doc0, n, 0, 3, 13, 0.0, device=dev)
    db, ob, nb = tok.synth_device(1234, doc0, n, 1, 3, 13, 0.0, device=dev)
    out = tok.encode_device(da, oa, db, ob, max_len=W, text_bytes=na, pair_bytes=nb)
    torch.cuda.synchronize()
    assert tok.check_errors(dev) == 0
    ha, hb = (da[:na].cpu().numpy(), oa.cpu().numpy()), (db[:nb].cpu().numpy(), ob.cpu().numpy())
    ref = oracle.encode_batch(ha, hb, max_len=W, threads=max(Oracle.max_threads(), os.cpu_count() or 1))
    assert np.array_equal(out["input_ids"].cpu().numpy().reshape(-1), ref["ids"])
    assert np.array_equal(out["attention_mask"].cpu().numpy().reshape(-1), ref["mask"])
    assert np.array_equal(out["row_status"].cpu().numpy(), ref["status"])
    tt = out["token_type_ids"].cpu().numpy()
    assert (np.diff(ref["tt_off"]) == W).all()
    assert np.array_equal(tt.reshape(-1).astype(np.int32), ref["tt"])
    assert np.array_equal(out["seq_len"].cpu().numpy(), np.diff(ref["seq_off"]))
    assert np.array_equal(out["row_len"].cpu().numpy().astype(np.int64), ref["mask"].reshape(n, W).sum(axis=1))
    if doc0 == 0:
        # BASELINE configs[3] at full size: the chunk's [1,048,576 x 256] ids decoded on the device in one go (sizes + text into a
        # ring, the write kernel picked on the device, 32,769 tiles of junctions handed out by the counter); three slices of 65,536
        # rows -- the first, one from the middle that starts inside a tile, the last -- against the oracle, byte for byte
        ring = torch.empty((1 << 31,), dtype=torch.uint8, device=dev)
        txt, off = tok.decode_device(out["input_ids"], out=ring)
        assert txt.data_ptr() == ring.data_ptr() and tok.check_errors(dev) == 0
        ho = off.cpu().numpy()
        assert ho[0] == 0 and ho[-1] == txt.numel() and (np.diff(ho) > 0).all()
        ids = ref["ids"].reshape(n, W)
        for lo in (0, 524288 + 13, n - 65536):
            hi = lo + 65536
            want = oracle.decode_batch(ids[lo:hi].reshape(-1), np.arange(0, 65536 * W + 1, W, dtype=np.int64), threads=max(Oracle.max_threads(), os.cpu_count() or 1))
            blob = txt[int(ho[lo]):int(ho[hi])].cpu().numpy().tobytes()
            rel = ho[lo:hi + 1] - ho[lo]
            assert [blob[rel[i]:rel[i + 1]].decode("utf-8") for i in range(65536)] == want, lo
        txt2, off2 = tok.decode_device(out["input_ids"])             # the two-step protocol gives the same bytes
        assert torch.equal(off2, off) and torch.equal(txt2, txt)


def test_cuda_path_against_the_reference_itself(tok):
    """When oracle/_ref travelled to this box (oracle/make_ref.py), compare the CUDA path with the UNMODIFIED Python reference
    directly -- no oracle in between -- on noisy pairs of the bench workload, encode and decode."""
    from oracle import ref_pool
    if not ref_pool.available():
        pytest.skip("oracle/_ref is not on this box")
    from genz_tokenize_b200 import workload
    n, W = 3000, 48
    ta, tb = workload.generate_hashed(99, 10, n, 0, 3, 13, 0.05), workload.generate_hashed(99, 10, n, 1, 3, 13, 0.05)
    sa, sb = workload.unpack(*ta), workload.unpack(*tb)
    rows = ref_pool.encode_rows_single(sa, sb, W)
    be = tok.encode_batch(ta, tb, max_len=W)
    for i, r in enumerate(rows):
        if r is None:
            assert be["row_status"][i] == 1, i
            continue
        assert be["row_status"][i] == 0, i
        got = tok._row_to_dict(be, i)
        assert got == r, (i, sa[i], sb[i])
    texts = tok.decode_batch(be["input_ids"][:500])
    ref_tok = ref_pool._tok
    for i in range(500):
        assert texts[i] == ref_tok.decode(be["input_ids"][i].tolist()), i


def test_token_type_padding_is_the_pad_id(oracle):
    """A pad token that is a vocab word (pad id != 0): token_type_ids is padded with the PAD ID (tokenize.py:256-258), the mask
    compares with it (tokenize.py:148-152).  Reference vectors through the drop-in call, and a batch in the fixed layout (store
    path: the TMA planes pad token types with zeros) against the oracle."""
    from golden_util import load_padtt_golden
    from genz_tokenize_b200 import Tokenize, workload
    from genz_tokenize_b200.data import bundled_paths
    from oracle.oracle import Oracle
    v, b = bundled_paths()
    for blk in load_padtt_golden():
        t = Tokenize(pad_token=blk["pad_token"], devices=[0])
        check_cases(t, blk["calls"], "pad token %r" % blk["pad_token"])
        o = Oracle(v, b, [blk["pad_token"], None, None, None, None])
        ta, tb = workload.generate_hashed(21, 0, 20000, 0, 3, 13, 0.02), workload.generate_hashed(21, 0, 20000, 1, 3, 13, 0.02)
        for W in (48, 64):
            be = t.encode_batch(ta, tb, max_len=W)
            assert_matches_oracle(be, o.encode_batch(ta, tb, max_len=W, threads=8), pad_id=blk["pad_id"], what="pad token %r W=%d" % (blk["pad_token"], W))


def test_recycled_result_planes_hold_no_stale_columns(tok, oracle):
    """The host path copies back only the columns that can differ from padding and relies on a recycled result buffer for the
    rest (genztok.cu, fixed layout): batches of the same shape with long rows first, then short rows, then long ones again, in
    several chunks, every plane against the oracle each time."""
    from genz_tokenize_b200 import workload
    n, W = 5000, 96
    tok.set_option("chunk_rows", 1536)                                     # four chunks: the two sets of chunk buffers alternate
    try:
        for k, (lo, hi, noise) in enumerate([(20, 60, 0.02), (1, 4, 0.0), (3, 13, 0.05), (0, 2, 0.0), (30, 40, 0.0)]):
            ta, tb = workload.generate_hashed(77 + k, 0, n, 0, lo, hi, noise), workload.generate_hashed(77 + k, 0, n, 1, lo, hi, noise)
            be = tok.encode_batch(ta, tb, max_len=W)
            assert_matches_oracle(be, oracle.encode_batch(ta, tb, max_len=W, threads=8), what="recycled planes, batch %d" % k)
            assert be.d2h_bytes > 0
            del be
        ta = workload.generate_hashed(5, 0, n, 0, 1, 3, 0.0)                # single sentences in a buffer that held pairs of the same size
        be = tok.encode_batch(ta, max_len=W)
        assert_matches_oracle(be, oracle.encode_batch(ta, None, max_len=W, threads=8), what="recycled planes, singles")
    finally:
        tok.set_option("chunk_rows", 1 << 18)


def test_add_vocab_and_bpe_file_after_construction(tmp_path):
    """tokenize.py:44-57 called after __init__: `encoder` grows, `decoder` stays as built (added ids decode to the unk token),
    `bpe_ranks` is replaced -- reference vectors (oracle/gen_golden_addvocab.py)."""
    import gzip, json
    from genz_tokenize_b200 import Tokenize
    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "addvocab_v1.json.gz"), "rb") as f:
        g = json.loads(f.read().decode("ascii"))
    v2, c2 = tmp_path / "v2.txt", tmp_path / "c2.codes"
    v2.write_text(g["vocab2"], encoding="utf-8")
    c2.write_text(g["codes2"], encoding="utf-8")
    tok = Tokenize(devices=[0])
    assert tok.vocab_size() == g["n0"]
    tok.add_vocab_file(str(v2))
    s = g["steps"][0]
    assert tok.vocab_size() == s["vocab_size"] and len(tok.decoder) == s["decoder_len"]
    assert {w: tok.encoder.get(w) for w in s["enc"]} == s["enc"]
    for c in s["calls"]:
        assert tok(c["text"], max_len=12) == c["out"], c["text"]
    for c in s["pairs"]:
        assert tok(c["text"], c["pair"], max_len=16) == c["out"]
    for d in s["decode"]:
        assert tok.decode(d["ids"]) == d["out"], d["ids"]
    tok.add_bpe_file(str(c2))
    s = g["steps"][1]
    assert tok.vocab_size() == s["vocab_size"] and len(tok.bpe_ranks) == s["n_ranks"]
    for c in s["calls"]:
        assert tok(c["text"], max_len=12) == c["out"], c["text"]
    for b in s["bpe"]:
        assert tok.bpe(b["w"]) == b["out"], b["w"]
    for d in s["decode"]:
        assert tok.decode(d["ids"]) == d["out"], d["ids"]
    import glob, tempfile
    assert not glob.glob(os.path.join(tempfile.gettempdir(), "genztok_vocab_*.txt")), "merged vocab copies must not pile up"


def test_decode_batch_ids_beyond_int32_are_unknown(tok):
    ids = np.array([2 ** 32 + 5, 770, -(2 ** 40), 5], dtype=np.int64)
    assert tok.decode_batch(ids, np.array([0, 4], dtype=np.int64)) == [tok.decode(ids.tolist())] == ["<unk> sinh_viên <unk> " + tok.decoder[5]]
    be = tok.encode_batch(["xin chào"], ["hello"], max_len=8, sequence_id=False)
    with pytest.raises(ValueError):
        be.row(0)
