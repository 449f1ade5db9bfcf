"""Compare a genz_tokenize_b200 BatchEncoding with the oracle's ragged result, bit for bit."""
import numpy as np


def ragged_take(flat, starts, lens):
    """Concatenate flat[starts[r] : starts[r]+lens[r]] for all r (vectorised)."""
    lens = np.asarray(lens, dtype=np.int64)
    total = int(lens.sum())
    if total == 0:
        return flat[:0]
    first = np.cumsum(lens) - lens
    idx = np.arange(total, dtype=np.int64) - np.repeat(first, lens) + np.repeat(np.asarray(starts, dtype=np.int64), lens)
    return flat[idx]


def assert_matches_oracle(be, orc, eos_id=2, what="", pad_id=0):
    n = orc["n"]
    assert be._n == n
    olen = np.diff(orc["ids_off"])
    if be._width > 0:
        starts = np.arange(n, dtype=np.int64) * be._width
        assert (olen == be._width).all(), "%s: oracle rows are not all %d long" % (what, be._width)
    else:
        starts = be._row_off[:-1]
        assert np.array_equal(be._row_off, orc["ids_off"]), "%s: row offsets differ" % what
    ours_ids = be._ids if be._width == 0 else be._ids
    assert len(ours_ids) == len(orc["ids"]), (what, len(ours_ids), len(orc["ids"]))
    bad = np.nonzero(ours_ids != orc["ids"])[0]
    if len(bad):
        r = int(np.searchsorted(orc["ids_off"], bad[0], side="right") - 1)
        raise AssertionError("%s: input_ids differ first at flat %d (row %d):\n ours=%s\n ref =%s" % (
            what, bad[0], r, ours_ids[orc["ids_off"][r]:orc["ids_off"][r + 1]].tolist(), orc["ids"][orc["ids_off"][r]:orc["ids_off"][r + 1]].tolist()))
    assert np.array_equal(be._mask, orc["mask"]), "%s: attention_mask differs" % what
    assert int(be["real_tokens"]) == int(orc["mask"].sum()), "%s: real_tokens" % what
    if not be._has_pair:
        return
    assert np.array_equal(be._status, orc["status"]), "%s: ValueError rows differ: ours %s ref %s" % (
        what, np.nonzero(be._status)[0][:10], np.nonzero(orc["status"])[0][:10])
    ok = orc["status"] == 0
    if be._seq is not None:
        sl = np.where(ok, be._seq_len, 0)
        assert np.array_equal(sl, np.diff(orc["seq_off"])), "%s: sequence_id lengths differ" % what
        ours = ragged_take(be._seq, starts, sl).astype(np.int32)
        bad = np.nonzero(ours != orc["seq"])[0]
        if len(bad):
            r = int(np.searchsorted(orc["seq_off"], bad[0], side="right") - 1)
            raise AssertionError("%s: sequence_id differs in row %d:\n ours=%s\n ref =%s\n ids =%s" % (
                what, r, be._seq[starts[r]:starts[r] + sl[r]].tolist(), orc["seq"][orc["seq_off"][r]:orc["seq_off"][r + 1]].tolist(),
                orc["ids"][orc["ids_off"][r]:orc["ids_off"][r + 1]].tolist()))
    if be._tt is not None:
        if be._pad_mode:
            tl = np.where(ok, be._tt_len, 0)
            ours = ragged_take(be._tt, starts, tl).astype(np.int32)
            ours = np.where(ours == -3, eos_id, np.where(ours == -4, pad_id, ours))
            assert np.array_equal(tl, np.diff(orc["tt_off"])), "%s: token_type_ids lengths differ" % what
            bad = np.nonzero(ours != orc["tt"])[0]
            if len(bad):
                r = int(np.searchsorted(orc["tt_off"], bad[0], side="right") - 1)
                raise AssertionError("%s: token_type_ids differ in row %d:\n ours=%s\n ref =%s\n ids =%s" % (
                    what, r, be._tt[starts[r]:starts[r] + tl[r]].tolist(), orc["tt"][orc["tt_off"][r]:orc["tt_off"][r + 1]].tolist(),
                    orc["ids"][orc["ids_off"][r]:orc["ids_off"][r + 1]].tolist()))
        else:
            # without padding token_type_ids is the sequence_id list itself (tokenize.py:254-255)
            assert np.array_equal(orc["tt"], orc["seq"])


def pad_run_rows(rng, n_rows, width, pad, vocab, exotic):
    """Fixed-width id rows shaped like encoder output and unlike it: real ids, pad runs at the end / in the middle / at the
    start, a non-pad last id behind a run, whole rows of pads, out-of-range ids."""
    ids = np.full((n_rows, width), pad, dtype=np.int32)
    for r in range(n_rows):
        kind = rng.integers(0, 8)
        nl = int(rng.integers(0, width + 1))
        if kind == 0:
            nl = 0
        elif kind == 1:
            nl = width
        ids[r, :nl] = rng.integers(0, vocab, nl)
        if kind == 2 and nl:                                  # pads inside the real part
            k = rng.integers(0, nl, max(1, nl // 3))
            ids[r, k] = pad
        if kind == 3 and width:                               # something behind the run
            ids[r, width - 1] = rng.integers(0, vocab)
        if kind == 4 and width >= 3 and nl < width:                          # a second island in the run
            ids[r, int(rng.integers(nl, width))] = rng.integers(0, vocab)
        if kind == 5 and nl:
            ids[r, rng.integers(0, nl)] = exotic[rng.integers(0, len(exotic))]
    return ids
