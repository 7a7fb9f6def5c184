"""Device timing of the preprocessing kernels on raw prefixes: poly(A) detection
(riser_polya_end, 2*L algorithmic bytes per read), window selection and normalisation.
usage: python tools/time_preprocess.py [B] [prefix_len]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, RaggedBatch, synth, _lib   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18000
proc = SignalProcessor(Kit.create_from_version("RNA002"))
pool = synth.raw_reads(5, 256, min_body=n, max_body=n + 1000)
sigs = [pool[i % len(pool)][1][:n] for i in range(B)]
batch = RaggedBatch(sigs, torch.device("cuda"))


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


ms = timed(lambda: proc.polya_end_device(batch))
ends = proc.polya_end_device(batch).cpu().numpy()
print(f"polya_end: {ms:.3f} ms for {B} x {n} samples -> {B / ms * 1e3:.0f} reads/s, "
      f"{2 * n * B / ms / 1e6:.0f} GB/s algorithmic (found in {(ends > 0).mean() * 100:.0f}% of reads)")
start = torch.from_numpy(np.where(ends > 0, ends + 1, 0).astype(np.int32)).cuda()
length = torch.from_numpy(np.minimum(n - np.where(ends > 0, ends + 1, 0), 12048).astype(np.int32)).cuda()
out = torch.zeros(B, 12048, device="cuda")
L = _lib.lib()
ms = timed(lambda: _lib.check(L.riser_normalise(_lib.ptr(batch.sig), _lib.ptr(batch.off), _lib.ptr(start),
                                                _lib.ptr(length), B, 12048, _lib.ptr(out), out.stride(0), None,
                                                _lib.stream_ptr()), "normalise"))
tot = int(length.sum().item())
print(f"normalise (trimmed windows, odd alignment): {ms:.3f} ms -> {6 * tot / ms / 1e6:.0f} GB/s algorithmic")
