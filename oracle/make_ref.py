#!/usr/bin/env python3
"""Recipe for oracle/_ref/: the UNMODIFIED reference tokenizer, so that it travels to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, but the
working tree is snapshotted there (git-ignored files included): this script copies the three files the hot path
consists of -- genz_tokenize/tokenize.py, data/vocab.txt, data/bpe.codes -- byte for byte from the reference checkout
into the git-ignored directory oracle/_ref/genz_tokenize/ and writes a one-line __init__.py beside them (the
reference's own __init__ also imports its TensorFlow model zoo, which is neither needed nor installed).  Nothing under
oracle/_ref/ is ever committed; nothing in the product imports it.  Users: bench.py (`cpu_baseline.python_*`,
`--impl reference`) and tests/ (direct CUDA-vs-reference comparison when the directory exists).

    python oracle/make_ref.py [--ref /root/reference]

Import it with  sys.path.insert(0, "oracle/_ref"); from genz_tokenize import Tokenize  -- never put the package
directory itself on sys.path (its tokenize.py would shadow the stdlib module of that name).
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["genz_tokenize/tokenize.py", "genz_tokenize/data/vocab.txt", "genz_tokenize/data/bpe.codes"]


def make(ref="/root/reference", quiet=False):
    """Returns the path of oracle/_ref (created or refreshed), or None when there is no reference checkout here."""
    if not all(os.path.isfile(os.path.join(ref, f)) for f in FILES):
        return DEST if available() else None
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(ref, f), os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
    with open(os.path.join(DEST, "genz_tokenize", "__init__.py"), "w") as fh:
        fh.write("from .tokenize import Tokenize  # written by oracle/make_ref.py (the reference's __init__ also imports its TF models)\n")
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "sha256": manifest}, fh, indent=1)
    if not quiet:
        print("oracle/_ref ready:", ", ".join(FILES))
    return DEST


def available():
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES) and os.path.isfile(os.path.join(DEST, "genz_tokenize", "__init__.py"))


def load_reference():
    """The reference's Tokenize class from oracle/_ref (ImportError when the directory was never made)."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    from genz_tokenize import Tokenize
    return Tokenize


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if make(a.ref) else 1)
