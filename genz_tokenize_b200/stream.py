"""The step before the tokenizer: a streaming reader that turns a text file with one document per
line into the C ABI's packed batches -- (uint8 bytes, int64 offsets[n+1]) -- without building Python strings.

Every document keeps its line terminator as a trailing ASCII space (the `\\n` becomes `0x20`; a `\\r` before it is
whitespace anyway), so the documents of a batch are adjacent in one buffer and tokenise exactly like the stripped lines:
trailing whitespace produces no word (`tokenize.py:106`), whereas a kept `\\n` would attach to the last word and turn it
into `<unk>`.  An empty line is an empty document (`[<s>, </s>]`).
"""
import numpy as np


def iter_line_batches(path, docs_per_batch=1 << 20, read_bytes=64 << 20):
    """Yield (bytes uint8[...], offsets int64[n+1]) with at most `docs_per_batch` documents (lines of `path`) each."""
    docs_per_batch = int(docs_per_batch)
    if docs_per_batch < 1:
        raise ValueError("docs_per_batch must be positive")
    carry = np.empty(0, dtype=np.uint8)          # bytes behind the last complete line
    with open(path, "rb") as f:
        while True:
            raw = f.read(int(read_bytes))
            eof = not raw
            if eof:
                if len(carry) == 0:
                    return
                buf = np.concatenate([carry, np.array([10], dtype=np.uint8)])      # last line without a terminator
                carry = np.empty(0, dtype=np.uint8)
            else:
                buf = np.concatenate([carry, np.frombuffer(raw, dtype=np.uint8)]) if len(carry) else np.frombuffer(raw, dtype=np.uint8).copy()
            ends = np.flatnonzero(buf == 10) + 1                                   # one past every '\n'
            if len(ends) == 0:
                carry = buf
                if eof:
                    return
                continue
            carry = buf[ends[-1]:].copy()
            buf = buf[:ends[-1]]
            buf[ends - 1] = 32
            starts = np.concatenate([[0], ends]).astype(np.int64)
            for b in range(0, len(ends), docs_per_batch):
                e = min(b + docs_per_batch, len(ends))
                lo, hi = starts[b], starts[e]
                yield buf[lo:hi], starts[b:e + 1] - lo
            if eof:
                return


def iter_device_batches(path, docs_per_batch=1 << 20, read_bytes=64 << 20, device=None, depth=2):
    """`iter_line_batches` delivered on the GPU, with the three stages overlapped: a reader thread reads the file and scans it for
    line ends into pinned staging buffers (`depth` slots), the host->device copies run on their own CUDA stream, and the consumer
    -- which tokenises batch i on its stream meanwhile -- only waits on the copy's event.  Yields
    `(d_bytes uint8, d_off int64[n+1], nbytes)`: `d_bytes` is padded so that the kernels' 16-byte loads stay inside it
    (`Tokenize.encode_device(d_bytes, d_off, ..., text_bytes=nbytes)`).  The tensors of a batch belong to its slot: they are
    valid until the generator is asked for the next batch (with `depth` > 2: for `depth - 2` more batches)."""
    import queue
    import threading
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    depth = max(2, int(depth))
    copy_stream = torch.cuda.Stream(device=dev)

    class Slot:
        def __init__(self):
            self.h_bytes = self.h_off = self.d_bytes = self.d_off = None
            self.copied = torch.cuda.Event()           # recorded on the copy stream: the batch is on the device
            self.released = None                       # recorded on the consumer's stream when it is done with the slot

        def fit(self, nb, no):
            if self.h_bytes is None or self.h_bytes.numel() < nb:
                cap = int(nb * 1.25) + 4096
                self.h_bytes = torch.empty((cap,), dtype=torch.uint8, pin_memory=True)
                self.d_bytes = torch.empty((cap,), dtype=torch.uint8, device=dev)
            if self.h_off is None or self.h_off.numel() < no:
                cap = int(no * 1.25) + 16
                self.h_off = torch.empty((cap,), dtype=torch.int64, pin_memory=True)
                self.d_off = torch.empty((cap,), dtype=torch.int64, device=dev)

    free, ready = queue.Queue(), queue.Queue(maxsize=depth)
    for _ in range(depth):
        free.put(Slot())
    stop = threading.Event()

    def reader():
        try:
            torch.cuda.set_device(dev)
            for b, o in iter_line_batches(path, docs_per_batch, read_bytes):
                slot = free.get()
                if stop.is_set():
                    return
                nb, no = len(b), len(o)
                pad = nb + (-nb) % 16 + 32
                slot.fit(pad, no)
                hb = slot.h_bytes.numpy()
                hb[:nb] = b
                hb[nb:pad] = 0
                slot.h_off.numpy()[:no] = o
                with torch.cuda.stream(copy_stream):
                    if slot.released is not None:
                        copy_stream.wait_event(slot.released)          # the consumer's kernels on the slot's old contents
                    slot.d_bytes[:pad].copy_(slot.h_bytes[:pad], non_blocking=True)
                    slot.d_off[:no].copy_(slot.h_off[:no], non_blocking=True)
                    slot.copied.record(copy_stream)
                ready.put((slot, nb, pad, no))
            ready.put(None)
        except BaseException as e:                                    # hand the failure to the consumer
            ready.put(e)

    th = threading.Thread(target=reader, name="genztok-line-reader", daemon=True)
    th.start()
    held = []
    try:
        while True:
            item = ready.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            slot, nb, pad, no = item
            torch.cuda.current_stream(dev).wait_event(slot.copied)
            held.append(slot)
            yield slot.d_bytes[:pad], slot.d_off[:no], nb
            # the consumer is back: what it launched on the oldest slot is in its stream by now -> the reader may refill it
            while len(held) >= depth - 1 and held:
                s = held.pop(0)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                s.released = ev
                free.put(s)
    finally:
        stop.set()
        while not ready.empty():                                      # a reader blocked on a full queue
            try:
                ready.get_nowait()
            except queue.Empty:
                break
        for _ in range(depth):
            free.put(Slot.__new__(Slot))                             # wake the reader if it waits for a slot
        th.join(timeout=5)
