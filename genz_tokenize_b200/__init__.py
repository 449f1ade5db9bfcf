"""genz_tokenize_b200 -- B200-native encode/decode engine behind the genz_tokenize `Tokenize` API.

    from genz_tokenize_b200 import Tokenize      # drop-in for genz_tokenize.Tokenize

All tokenisation runs in hand-written sm_100a CUDA kernels inside libgenztok.so (C ABI in
include/genztok.h); this package is the thin ctypes wrapper plus the synthetic-workload
generator used by the benchmark.  There is no CPU fallback.
"""
from . import preprocess
from .collection import DataCollection
from .stream import iter_device_batches, iter_line_batches
from .tokenizer import BatchEncoding, GenztokError, Tokenize, pack_strings

__all__ = ["Tokenize", "BatchEncoding", "GenztokError", "pack_strings", "preprocess", "DataCollection", "iter_line_batches", "iter_device_batches"]
