"""Bundled tokenizer model files.

`bundled.pack` holds the reference's vocab.txt / bpe.codes (the files `Tokenize()` loads by
default, reference tokenize.py:15-23) zlib-compressed; `bundled_paths()` unpacks them once into
`_unpacked/` beside this file (or a temp dir when the package is read-only) and returns real
paths, because both the C loader and the drop-in `vocab_file` / `bpe_file` attributes want paths.
"""
import os
import struct
import tempfile
import zlib

_MAGIC = b"GZTPACK1"
_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = None


def _read_pack():
    blob = open(os.path.join(_HERE, "bundled.pack"), "rb").read()
    if blob[:8] != _MAGIC:
        raise RuntimeError("bundled.pack: bad magic")
    (count,) = struct.unpack_from("<I", blob, 8)
    pos, files = 12, {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", blob, pos)
        pos += 4
        name = blob[pos:pos + nl].decode()
        pos += nl
        raw_len, comp_len, crc = struct.unpack_from("<QQI", blob, pos)
        pos += 20
        raw = zlib.decompress(blob[pos:pos + comp_len])
        pos += comp_len
        if len(raw) != raw_len or zlib.crc32(raw) != crc:
            raise RuntimeError("bundled.pack: %s is corrupt" % name)
        files[name] = raw
    return files


def bundled_paths():
    """Return (vocab_path, bpe_path) of the bundled model, unpacking on first use."""
    global _cache
    if _cache is not None and all(os.path.exists(p) for p in _cache):
        return _cache
    files = _read_pack()
    for root in (os.path.join(_HERE, "_unpacked"), os.path.join(tempfile.gettempdir(), "genztok_b200_data_%d" % os.getuid())):
        try:
            os.makedirs(root, exist_ok=True)
            paths = []
            for name in ("vocab.txt", "bpe.codes"):
                p = os.path.join(root, name)
                if not (os.path.exists(p) and os.path.getsize(p) == len(files[name])):
                    tmp = "%s.%d.tmp" % (p, os.getpid())
                    with open(tmp, "wb") as f:
                        f.write(files[name])
                    os.replace(tmp, p)
                paths.append(p)
            _cache = tuple(paths)
            return _cache
        except OSError:
            continue
    raise RuntimeError("cannot unpack the bundled vocab/bpe.codes anywhere writable")
