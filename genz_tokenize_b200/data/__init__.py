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

    def intact(p, raw):
        """An unpacked file is trusted only when its bytes are the pack's (a same-size file somebody else put there is not)."""
        try:
            with open(p, "rb") as f:
                return f.read() == raw
        except OSError:
            return False

    def unpack_into(root):
        paths = []
        for name in ("vocab.txt", "bpe.codes"):
            p = os.path.join(root, name)
            if not intact(p, files[name]):
                tmp = "%s.%d.tmp" % (p, os.getpid())
                with open(tmp, "wb") as f:
                    f.write(files[name])
                os.replace(tmp, p)
            paths.append(p)
        return tuple(paths)

    try:                                                       # beside the package (the usual case: an in-tree or user-owned install)
        root = os.path.join(_HERE, "_unpacked")
        os.makedirs(root, exist_ok=True)
        _cache = unpack_into(root)
        return _cache
    except OSError:
        pass
    # read-only package: a private directory under the temp dir -- created by this user with mode 0700, or a fresh one
    # (a predictable name that another user of a shared host could have prepared is not used)
    root = os.path.join(tempfile.gettempdir(), "genztok_b200_data_%d" % os.getuid())
    try:
        os.mkdir(root, 0o700)
    except FileExistsError:
        st = os.lstat(root)
        import stat
        if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
            root = tempfile.mkdtemp(prefix="genztok_b200_data_")
    except OSError:
        root = tempfile.mkdtemp(prefix="genztok_b200_data_")
    try:
        _cache = unpack_into(root)
        return _cache
    except OSError:
        pass
    raise RuntimeError("cannot unpack the bundled vocab/bpe.codes anywhere writable")
