"""Offline evaluator with the reference's riser/test.py semantics: for every
read, optionally trim the adapter + poly(A) (dynamic cut-off, else the fixed trim), then
classify a ladder of prefix lengths (min length, +1 s steps, up to max_sec) re-normalising at
every length, and write the reference's TSV line (test.py:226).  Reads come from any iterable
of (read_id, int16 array) -- fast5 I/O (ont_fast5_api) is out of scope.

Differences from test.py, both documented there as TODOs: a constant read (MAD == 0) is
normalised to zeros like the live path (preprocess.py:123) instead of dividing by zero;
the dynamic trim runs at test.py's defaults (resolution 500, MAD threshold 20).
"""
import math

import numpy as np
import torch

from . import _lib
from .preprocess import RaggedBatch

# riser/test.py:16-26
SAMPLING_HZ = {"RNA002": 3012, "RNA004": 4000}
MIN_SIGNAL_LENGTH = 4096
MAX_SIGNAL_SEC = {"RNA002": 4, "RNA004": 2.15}
FIXED_TRIM = {"RNA002": 6481, "RNA004": 4634}


def ladder(kit):
    """Prefix lengths test.py:202-224 visits: ceil(min_sec*hz), += hz, while <= floor(max_sec*hz)."""
    hz = SAMPLING_HZ[kit]
    n = math.ceil(MIN_SIGNAL_LENGTH / hz * hz)
    top = math.floor(MAX_SIGNAL_SEC[kit] * hz)
    out = []
    while n <= top:
        out.append(n)
        n += hz
    return out


def evaluate(reads, model, processor, kit, already_trimmed=True, model_id="model", dataset="dataset",
             filename="reads", out_path=None, batch_reads=256):
    """-> list of TSV lines (also written to out_path if given)."""
    device = _lib.require_device()
    steps = ladder(kit)
    lines = []
    reads = list(reads)
    for lo in range(0, len(reads), batch_reads):
        chunk = reads[lo:lo + batch_reads]
        ids = [r for r, _ in chunk]
        sigs = [np.ascontiguousarray(s, dtype=np.int16) for _, s in chunk]
        batch = RaggedBatch(sigs, device)
        n = batch.n_host
        if already_trimmed:
            start = np.zeros(len(chunk), dtype=np.int32)
            pa_start = pa_end = ["boostnano"] * len(chunk)
        else:
            starts_d = torch.empty(len(chunk), dtype=torch.int32, device=device)
            ends = processor.polya_end_device(batch, starts=starts_d).cpu().numpy()
            starts = starts_d.cpu().numpy()
            start = np.where(ends > 0, ends + 1, FIXED_TRIM[kit]).astype(np.int32)      # test.py:193-198
            pa_start = [None if v <= 0 else int(v) for v in starts]
            pa_end = [None if v <= 0 else int(v) for v in ends]
        avail = np.maximum(n - start, 0)
        preds = [[] for _ in chunk]
        for L in steps:
            length = np.where(avail >= L, L, 0).astype(np.int32)                        # test.py:206-208
            if not length.any():
                continue
            x, len_t = processor.mad_normalise_batch(batch, start=start, length=length)
            if x.shape[1] < L or x.shape[1] % 2:
                xx = torch.zeros(x.shape[0], (L + 3) & ~3, dtype=torch.float32, device=device)
                xx[:, :x.shape[1]] = x
                x = xx
            probs = model.classify_batch(x, len_t, max_len=L).cpu().numpy()
            for i in np.flatnonzero(length):
                preds[i].append(f"{L}:{float(probs[i, 0])},{float(probs[i, 1])}")
        for i, rid in enumerate(ids):
            lines.append(f"{model_id}\t{dataset}\t{filename}\t{rid}\t{pa_start[i]}\t{pa_end[i]}\t{';'.join(preds[i])}\n")
    if out_path:
        with open(out_path, "w") as f:
            f.writelines(lines)
    return lines
