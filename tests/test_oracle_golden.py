"""Pin the CPU oracle (oracle/genztok_oracle.c) to vectors generated from the unmodified reference."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from golden_util import check_cases


def test_readme_vector(oracle, golden):
    # README.md:11-15 (ids + mask reproduce; sequence_id follows the code)
    check_cases(oracle, golden["readme"], "readme")
    assert oracle.decode(golden["readme_decode"]["ids"]) == golden["readme_decode"]["out"] == "<s> sinh_viên </s>"
    out = oracle("sinh_viên công_nghệ", "hello", max_len=10, padding=True, truncation=True)
    assert out["input_ids"] == [1, 770, 1444, 2, 2, 30469, 2, 0, 0, 0]
    assert out["attention_mask"] == [1, 1, 1, 1, 1, 1, 1, 0, 0, 0]


def test_bundled_tables(oracle, golden):
    assert oracle.vocab_size() == golden["meta"]["vocab_size"] == 48423
    assert oracle.special_ids() == [0, 1, 2, 3, 4]
    assert oracle.rank_get("#version:", "0.2") == 0
    assert oracle.rank_get("n", "g</w>") == 1


def test_corner_calls(oracle, golden):
    check_cases(oracle, golden["calls"], "calls")


def test_random_rows(oracle, golden):
    for blk in golden["random"]:
        pairs = blk["pairs"] or [None] * len(blk["texts"])
        cases = [{"text": t, "pair": p, "kw": blk["kw"], "out": o} for t, p, o in zip(blk["texts"], pairs, blk["out"])]
        check_cases(oracle, cases, "random seed %d" % blk["gen"]["seed"])


def test_random_rows_batch_threads(oracle, golden):
    # the batch entry (what bench.py's cpu_baseline times) must equal the per-row results, at any thread count
    blk = golden["random"][1]
    r1 = oracle.encode_batch(blk["texts"], blk["pairs"], threads=1, **blk["kw"])
    r4 = oracle.encode_batch(blk["texts"], blk["pairs"], threads=4, **blk["kw"])
    for k in ("ids", "ids_off", "mask", "seq", "seq_off", "tt", "tt_off", "status"):
        assert (r1[k] == r4[k]).all()
    for i, exp in enumerate(blk["out"]):
        if "raises" in exp:
            assert r1["status"][i] == 1
            continue
        assert r1["status"][i] == 0
        assert r1["ids"][r1["ids_off"][i]:r1["ids_off"][i + 1]].tolist() == exp["input_ids"]
        assert r1["mask"][r1["ids_off"][i]:r1["ids_off"][i + 1]].tolist() == exp["attention_mask"]
        assert [None if v == -1 else v for v in r1["tt"][r1["tt_off"][i]:r1["tt_off"][i + 1]].tolist()] == exp["token_type_ids"]


def test_bpe_strings(oracle, golden):
    for c in golden["bpe"]:
        assert oracle.bpe(c["w"]) == c["out"], c["w"]


def test_bpe_digest_all_vocab(oracle, golden):
    from genz_tokenize_b200.data import bundled_paths
    vocab, codes = bundled_paths()
    # rebuild the word list exactly as oracle/gen_golden.py did (encoder dict order, then merges order)
    enc = {"<pad>": 0, "<s>": 1, "</s>": 2, "<mask>": 3, "<unk>": 4}
    for line in open(vocab, encoding="utf-8").readlines():
        line = line.strip()
        enc[line[:line.rfind(" ")]] = len(enc)
    merges = [tuple(m.split()) for m in open(codes, encoding="utf-8").read().split("\n")[:-1]]
    ranks = dict(zip(merges, range(len(merges))))
    allw = [w[:-2] if w.endswith("@@") else w for w in enc.keys()]
    allw += ["".join(k).replace("</w>", "") for k in ranks.keys()]
    allw = [w for w in allw if w and not any(c.isspace() for c in w)]
    assert len(allw) == golden["bpe_digest"]["n"]
    hs = hashlib.sha256()
    for w in allw:
        hs.update(oracle.bpe(w).encode("utf-8"))
        hs.update(b"\n")
    assert hs.hexdigest() == golden["bpe_digest"]["sha256"]


def test_sequence_id_state_machine(oracle, golden):
    for ids, raw, tt in golden["seqid"]:
        assert oracle.sequence_id(ids) == raw, ids
        if tt == "ValueError":
            with pytest.raises(ValueError):
                oracle.sequence_id(ids, apply_token_type=True)
        else:
            assert oracle.sequence_id(ids, apply_token_type=True) == tt, ids


def test_decode(oracle, golden):
    for c in golden["decode"]:
        assert oracle.decode(c["ids"]) == c["out"], c["ids"]


def test_decode_pad_run_vectors():
    # reference decode() of rows with trailing / inner / leading pad runs, for pad tokens of every text length
    from golden_util import load_decode_golden
    from oracle.oracle import Oracle
    cache = {}
    for b in load_decode_golden():
        sp = [b["pad_token"], "<s>", "</s>", "<mask>", "<unk>"]
        o = cache.setdefault(b["pad_token"], Oracle(specials=sp))
        ids = np.array(b["ids"], dtype=np.int64).astype(np.int32)
        n, w = ids.shape
        assert o.decode_batch(ids.reshape(-1), np.arange(0, n * w + 1, w, dtype=np.int64)) == b["out"], (b["pad_token"], w)


def _oracle_digest(orc, paired):
    from golden_util import row_bytes
    h = hashlib.sha256()
    io, so, to = orc["ids_off"], orc.get("seq_off"), orc.get("tt_off")
    for i in range(orc["n"]):
        if orc["status"][i]:
            h.update(row_bytes(1))
            continue
        a, b = io[i], io[i + 1]
        h.update(row_bytes(0, orc["ids"][a:b], orc["mask"][a:b],
                           orc["seq"][so[i]:so[i + 1]] if paired else None, orc["tt"][to[i]:to[i + 1]] if paired else None))
    return h.hexdigest()


def test_encode_digests_85k_reference_rows(oracle):
    # SHA-256 over the reference's outputs on 85,000 seeded rows (oracle/gen_golden_digest.py): BASELINE-shaped batches plus
    # noisy, heavily truncated (4,001 ValueError rows), ragged and unpadded regimes
    from genz_tokenize_b200 import workload
    from golden_util import load_encode_digests
    for c in load_encode_digests():
        t = workload.generate(c["seed"], c["n"], c["lo"], c["hi"], c["noise"])
        p = workload.generate(c["seed"] + 1000, c["n"], c["lo"], c["hi"], c["noise"]) if c["paired"] else None
        orc = oracle.encode_batch(t, p, threads=8, **c["kw"])
        assert int(orc["status"].sum()) == c["value_errors"], c
        assert _oracle_digest(orc, c["paired"]) == c["sha256"], c
        # decode() of every row the reference returned (rows it raised on have no ids to decode)
        texts = oracle.decode_batch(orc["ids"], orc["ids_off"], threads=8)
        hd = hashlib.sha256()
        for i, s in enumerate(texts):
            if not orc["status"][i]:
                hd.update(s.encode("utf-8", "surrogatepass") + b"\n")
        assert hd.hexdigest() == c["decode_sha256"], c


def test_loader_quirks(golden):
    from oracle.oracle import Oracle
    with tempfile.TemporaryDirectory() as td:
        for i, L in enumerate(golden["loaders"]):
            vp, mp = os.path.join(td, "v.txt"), os.path.join(td, "m.codes")
            open(vp, "wb").write(L["vocab"].encode("utf-8"))
            open(mp, "wb").write(L["merges"].encode("utf-8"))
            o = Oracle(vp, mp)
            assert o.vocab_size() == L["vocab_size"], i
            assert o.special_ids() == L["special_ids"], i
            for k, v in L["encoder"].items():
                assert o.encoder_get(k) == v, (i, k)
            for k in range(-1, L["vocab_size"] + 3):
                assert o.decoder_get(k) == L["decoder"].get(str(k)), (i, k)
            for a, b, r in L["ranks2"]:
                assert o.rank_get(a, b) == r, (i, a, b)
            check_cases(o, L["calls"], "loader %d" % i)
            for d in L["decode"]:
                assert o.decode(d["ids"]) == d["out"], (i, d["ids"])


def test_custom_specials(golden):
    from oracle.oracle import Oracle
    from genz_tokenize_b200.data import bundled_paths
    v, b = bundled_paths()
    cs = golden["custom_specials"]
    o = Oracle(v, b, cs["specials"])
    assert o.vocab_size() == cs["vocab_size"]
    check_cases(o, cs["calls"], "custom specials")
    for d in cs["decode"]:
        assert o.decode(d["ids"]) == d["out"]


def test_missing_file():
    from oracle.oracle import Oracle
    with pytest.raises(FileNotFoundError):
        Oracle("/nonexistent/vocab.txt", "/nonexistent/bpe.codes")


def test_token_type_padding_is_the_pad_id():
    """tokenize.py:256-258: token_type_ids is padded by __padding, i.e. with encoder[pad_token] -- reference vectors with pad ids 5 ... 15117."""
    from golden_util import load_padtt_golden, check_cases
    from oracle.oracle import Oracle
    from genz_tokenize_b200.data import bundled_paths
    v, b = bundled_paths()
    for blk in load_padtt_golden():
        o = Oracle(v, b, [blk["pad_token"], None, None, None, None])
        check_cases(o, blk["calls"], "pad token %r" % blk["pad_token"])
