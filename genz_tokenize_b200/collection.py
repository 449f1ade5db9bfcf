"""The step after the tokenizer: the reference's `DataCollection` field bundle
(`genz_tokenize/models/bert/dataset.py:6-55`) for tensors that stay on the GPU.

The reference bundles numpy / TensorFlow arrays and turns them into a shuffled, batched `tf.data.Dataset` of
`({field: batch}, y)`; here the fields are whatever `Tokenize.encode_device` / `encode_batch` returned (torch tensors on
any device, or numpy arrays); for fields on one GPU a batch of ALL fields is cut out by one launch of the library's gather kernel
(`genztok_gather_rows`, csrc/gather.cuh), and every field can be handed to another framework through DLPack.  Nothing is copied
to the host.  Fields on the host (numpy, CPU tensors) are batched there with index_select, like the reference does in TensorFlow.
"""
import ctypes as C

import numpy as np

_FIELDS = ("input_ids", "attention_mask", "token_type_ids", "dec_input_ids", "dec_attention_mask", "dec_token_type_ids", "y")


class DataCollection:
    def __init__(self, input_ids=None, attention_mask=None, token_type_ids=None, dec_input_ids=None,
                 dec_attention_mask=None, dec_token_type_ids=None, y=None):
        self.input_ids = input_ids
        self.attention_mask = attention_mask
        self.token_type_ids = token_type_ids
        self.dec_input_ids = dec_input_ids
        self.dec_attention_mask = dec_attention_mask
        self.dec_token_type_ids = dec_token_type_ids
        self.y = y
        if y is None:
            raise Exception('y (label) is required')                  # dataset.py:25-26
        n = {len(v) for v in self.fields().values()}
        if len(n) > 1:
            raise ValueError("fields differ in their number of rows: %r" % {k: len(v) for k, v in self.fields().items()})

    @classmethod
    def from_encoding(cls, enc, y, dec=None):
        """`enc` (and `dec` for the decoder side): what `Tokenize.encode_device` / `encode_batch` returned."""
        def pick(e, k):
            try:
                return e[k]
            except (KeyError, TypeError):
                return None
        kw = {k: pick(enc, k) for k in ("input_ids", "attention_mask", "token_type_ids")}
        if dec is not None:
            kw.update({"dec_" + k: pick(dec, k) for k in ("input_ids", "attention_mask", "token_type_ids")})
        return cls(y=y, **kw)

    def fields(self):
        """The fields that are set, in the reference's order (dataset.py:34-38 walks `__dict__`)."""
        return {k: getattr(self, k) for k in _FIELDS if getattr(self, k) is not None}

    def __len__(self):
        return len(self.y)

    def to_torch_batches(self, batch_size=32, shuffle=True, seed=None, tokenizer=None):
        """`to_tf_dataset` (dataset.py:28-55) without TensorFlow: shuffle over the whole collection, batches of `batch_size`
        (the last one may be short), each yielded as `({field: batch}, y)`.  Tensors stay on their device.  When every field is a
        contiguous tensor on one CUDA device the batches are gathered by `genztok_gather_rows` (one launch per batch for all the
        fields, on torch's current stream; `tokenizer`: the `Tokenize` whose library handle to use, default: a private one)."""
        import torch
        f = {k: (v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))) for k, v in self.fields().items()}
        n = len(self)
        dev = f["y"].device
        if shuffle:
            g = torch.Generator(device="cpu")
            if seed is not None:
                g.manual_seed(int(seed))
            order = torch.randperm(n, generator=g)
        else:
            order = torch.arange(n)
        bs = int(batch_size)
        on_gpu = dev.type == "cuda" and all(v.device == dev for v in f.values()) and len(f) <= 8
        if on_gpu:
            yield from self._gathered_batches(f, order.to(dev), n, bs, dev, tokenizer)
            return
        for b in range(0, n, bs):
            idx = order[b:b + bs]
            out = {k: v.index_select(0, idx.to(v.device)) for k, v in f.items()}
            y = out.pop("y")
            yield out, y.to(dev)

    _own_tok = None

    def _gathered_batches(self, f, order, n, bs, dev, tokenizer):
        import torch
        from . import _lib as L
        if tokenizer is None:
            if DataCollection._own_tok is None or DataCollection._own_tok[0] != dev.index:
                from .tokenizer import Tokenize
                DataCollection._own_tok = (dev.index, Tokenize(devices=[dev.index if dev.index is not None else torch.cuda.current_device()]))
            tokenizer = DataCollection._own_tok[1]
        lib, h = L.load(), tokenizer._h
        names = list(f)
        src = [f[k].contiguous() for k in names]
        nf = len(names)
        row_bytes = (C.c_int64 * nf)(*[(v.numel() // max(n, 1)) * v.element_size() for v in src])
        srcp = (C.c_void_p * nf)(*[v.data_ptr() for v in src])
        st = tokenizer._torch_stream(dev)
        for b in range(0, n, bs):
            m = min(bs, n - b)
            out = [torch.empty((m,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for v in src]
            dstp = (C.c_void_p * nf)(*[o.data_ptr() for o in out])
            rc = lib.genztok_gather_rows(h, 0, nf, srcp, row_bytes, n, order[b:b + m].data_ptr(), m, dstp, st)
            if rc:
                tokenizer._err(rc, "genztok_gather_rows")
            d = dict(zip(names, out))
            y = d.pop("y")
            yield d, y

    def to_dlpack(self):
        """Every field as a DLPack capsule (zero copy) for a consumer that is not torch."""
        import torch
        from torch.utils import dlpack
        return {k: dlpack.to_dlpack(v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))) for k, v in self.fields().items()}
