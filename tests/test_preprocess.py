"""preprocess.py normalisers: oracle pinned to reference-generated vectors (CPU), CUDA kernels
against the same vectors and against the oracle on large random batches (GPU)."""
import gzip
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["remove_html", "convert_unicode", "remove_punctuations", "remove_emoji", "remove_URL"]


@pytest.fixture(scope="module")
def pgold():
    with gzip.open(os.path.join(HERE, "golden", "preprocess_v1.json.gz"), "rb") as f:
        return json.loads(f.read().decode("ascii"))


def test_oracle_matches_reference_vectors(pgold):
    from oracle import oracle as O
    assert pgold["meta"]["ops"] == NAMES
    for op in range(5):
        got = O.preprocess(op, pgold["texts"])
        for t, g, e in zip(pgold["texts"], got, pgold["out"][op]):
            assert g == e, (NAMES[op], t, g, e)


@pytest.mark.gpu
def test_device_matches_reference_vectors(pgold):
    from genz_tokenize_b200 import preprocess as P
    fns = [P.remove_html, P.convert_unicode, P.remove_punctuations, P.remove_emoji, P.remove_URL]
    for op in range(5):
        got = P.run_batch(op, pgold["texts"])
        for t, g, e in zip(pgold["texts"], got, pgold["out"][op]):
            assert g == e, (NAMES[op], t, g, e)
        for t, e in list(zip(pgold["texts"], pgold["out"][op]))[:25]:      # the single-string drop-in functions
            assert fns[op](t) == e
    with pytest.raises(TypeError):
        P.remove_html(b"bytes")


@pytest.mark.gpu
def test_device_matches_oracle_on_large_batches():
    from genz_tokenize_b200 import preprocess as P, workload
    from oracle import oracle as O
    from genz_tokenize_b200.tokenizer import Tokenize
    tok = Tokenize()
    tok.set_option("chunk_rows", 3000)          # several chunks
    t = workload.generate(77, 20000, 0, 40, 0.25)
    docs = workload.unpack(*t)
    rng = np.random.default_rng(3)
    deco = ["<p>", "</p>", "http://a.b/c ", "https://x ", " 😀 ", "!", "?!", "à", "ế", "　", " <br/> ", "ợ"]
    docs = [d + deco[int(rng.integers(0, len(deco)))] + d[: int(rng.integers(0, 30))] for d in docs]
    for op in range(5):
        assert P.run_batch(op, docs, tok=tok) == O.preprocess(op, docs), NAMES[op]


@pytest.mark.gpu
def test_device_resident_chain_normalise_then_tokenise(pgold):
    """genztok_preprocess_device: text in, text out, both on the GPU -- the same bytes as the host form -- and straight into
    encode_device on the same stream: normalise -> tokenise without a trip through host memory."""
    import torch
    from genz_tokenize_b200 import preprocess as P, workload
    from genz_tokenize_b200.tokenizer import Tokenize, pack_strings
    from oracle import oracle as O
    from oracle.oracle import Oracle
    tok = Tokenize(devices=[0])
    dev = torch.device("cuda:0")
    docs = workload.unpack(*workload.generate_hashed(9, 0, 6000, 0, 0, 30, 0.1))
    rng = np.random.default_rng(4)
    deco = ["<p>", "</p>", "http://a.b/c ", " 😀 ", "!", "à", "　", " <br/> "]
    docs = [d + deco[int(rng.integers(0, len(deco)))] + d[: int(rng.integers(0, 20))] for d in docs] + list(pgold["texts"])
    tb, to = pack_strings(docs)
    d_t = torch.from_numpy(np.concatenate([tb, np.zeros(32, dtype=np.uint8)])).to(dev)
    d_o = torch.from_numpy(to).to(dev)
    for op in (P.REMOVE_HTML, P.REMOVE_URL, P.REMOVE_EMOJI):
        out, off, nb = P.run_device(op, d_t, d_o, tok=tok)
        want = O.preprocess(op, docs)
        wb, wo = pack_strings(want)
        assert nb == len(wb) and np.array_equal(off.cpu().numpy(), wo) and np.array_equal(out[:nb].cpu().numpy(), wb), NAMES[op]
        enc = tok.encode_device(out, off, max_len=64, text_bytes=nb)        # chained on the device
        torch.cuda.synchronize()
        ref = Oracle().encode_batch((wb, wo), None, max_len=64, threads=8)
        assert np.array_equal(enc["input_ids"].cpu().numpy().reshape(-1), ref["ids"]), NAMES[op]
        assert np.array_equal(enc["attention_mask"].cpu().numpy().reshape(-1), ref["mask"]), NAMES[op]
