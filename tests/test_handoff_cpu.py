"""Host-side glue on either side of the tokenizer: the streaming line reader and the DataCollection
bundle.  No GPU: the reader is numpy, the bundle is checked on CPU tensors (it never leaves the tensors' device)."""
import os
import tempfile

import numpy as np
import pytest


def _docs(b, o):
    raw = b.tobytes()
    return [raw[o[i]:o[i + 1]] for i in range(len(o) - 1)]


@pytest.mark.parametrize("read_bytes", [7, 64, 1 << 20])
@pytest.mark.parametrize("tail_newline", [True, False])
def test_line_reader_batches(read_bytes, tail_newline):
    from genz_tokenize_b200 import iter_line_batches
    lines = ["xin chào", "", "sinh_viên công_nghệ", "a\r", "  x  ", "ộ" * 40, "", "", "cuối"] * 7
    text = "\n".join(lines) + ("\n" if tail_newline else "")
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "docs.txt")
        open(p, "wb").write(text.encode("utf-8"))
        got, sizes = [], []
        for b, o in iter_line_batches(p, docs_per_batch=5, read_bytes=read_bytes):
            assert b.dtype == np.uint8 and o.dtype == np.int64 and o[0] == 0 and o[-1] == len(b)
            d = _docs(b, o)
            sizes.append(len(d))
            got += d
        assert max(sizes) <= 5
        assert [g.decode("utf-8") for g in got] == [l + " " for l in lines]          # the terminator became a space
        open(p, "wb").write(b"")
        assert list(iter_line_batches(p)) == []
        open(p, "wb").write(b"\n")
        assert [_docs(b, o) for b, o in iter_line_batches(p)] == [[b" "]]


def test_data_collection_like_the_reference():
    import torch
    from genz_tokenize_b200 import DataCollection
    with pytest.raises(Exception, match="y \\(label\\) is required"):            # dataset.py:25-26
        DataCollection(input_ids=np.zeros((4, 8), dtype=np.int32))
    with pytest.raises(ValueError):
        DataCollection(input_ids=np.zeros((4, 8), dtype=np.int32), y=np.zeros(5))
    n = 103
    ids = torch.arange(n * 8, dtype=torch.int32).reshape(n, 8)
    enc = {"input_ids": ids, "attention_mask": (ids % 3 != 0).to(torch.uint8), "row_len": torch.zeros(n)}
    y = np.arange(n, dtype=np.int64)
    dc = DataCollection.from_encoding(enc, y)
    assert list(dc.fields()) == ["input_ids", "attention_mask", "y"] and len(dc) == n
    seen = []
    for feats, yy in dc.to_torch_batches(batch_size=32, seed=5):
        assert set(feats) == {"input_ids", "attention_mask"}                       # y split off, like to_dict (dataset.py:42-49)
        assert feats["input_ids"].shape[0] == yy.shape[0] <= 32
        assert torch.equal(feats["input_ids"][:, 0].to(torch.int64), yy * 8)       # rows travel with their labels
        seen += yy.tolist()
    assert sorted(seen) == list(range(n)) and seen != list(range(n))               # a permutation, shuffled
    assert [yy.tolist() for _, yy in dc.to_torch_batches(batch_size=50, shuffle=False)] == [list(range(50)), list(range(50, 100)), [100, 101, 102]]
    caps = dc.to_dlpack()
    back = torch.utils.dlpack.from_dlpack(caps["input_ids"])
    assert back.data_ptr() == ids.data_ptr()
    both = DataCollection.from_encoding(enc, y, dec=enc)
    assert list(both.fields()) == ["input_ids", "attention_mask", "dec_input_ids", "dec_attention_mask", "y"]


def test_device_reader_needs_a_gpu(tmp_path):
    # iter_device_batches has no host fallback: without CUDA it fails at once (and leaves no reader thread behind)
    import threading
    import torch
    from genz_tokenize_b200 import iter_device_batches
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_file_to_device_batches_handoff")
    p = tmp_path / "docs.txt"
    p.write_bytes(b"mot hai ba\nbon nam\n")
    before = threading.active_count()
    with pytest.raises(Exception):
        next(iter_device_batches(str(p), docs_per_batch=1))
    assert threading.active_count() == before
