"""Host / device breakdown of one live poll (BatchedClassifier.classify_batch) for C channels: packing into the
pinned arena, H2D, kernels, D2H, bookkeeping.  usage: python tools/live_breakdown.py [channels] [n_models]"""
import json
import logging
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, RaggedBatch, synth   # noqa: E402
from riser_b200.config import shipped_config                                                # noqa: E402
from riser_b200 import pipeline                                                             # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2
log = logging.getLogger("lb")
targets = ["mRNA", "mtRNA", "globin"][:M]
models = [Model(synth.state_dict(synth.TARGET_SEEDS[t]), shipped_config(), log, t) for t in targets]
proc = SignalProcessor(Kit.create_from_version("RNA002"))
clf = BatchedClassifier(models, proc)
reads = synth.raw_reads(7, min(C, 1024), min_body=14000, max_body=20000, frac_no_polya=0.05)
rng = np.random.default_rng(0)
# a poll in steady state: prefixes of 1..6 s (3012 Hz), as the accumulating cache hands them over
sigs = [reads[i % len(reads)][1][:int(rng.integers(1, 7)) * 3012] for i in range(C)]
ids = [f"r{i}" for i in range(C)]
for _ in range(5):
    clf.classify_batch(sigs, ids, {}, 0.9, "deplete")
T = {}


def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3


N = 30
for _ in range(N):
    cache = {}
    t = time.perf_counter()
    B = pipeline.bucket_size(C)
    cached = np.full(B, -1, dtype=np.int32)
    ss = list(sigs) + [pipeline._EMPTY] * (B - C)
    clf._arena.reserve(B * (clf.fixed_trim + clf.max_len + 8192), B)
    tick("prep", t)
    t = time.perf_counter()
    batch = RaggedBatch(ss, clf.device, arena=clf._arena, trusted=True)
    tick("pack+h2d", t)
    t = time.perf_counter()
    start, length, det = clf.select_windows(batch, cached)
    tick("polya+select", t)
    t = time.perf_counter()
    dec, probs = clf.run_windows(batch, start, length, 0.9, "deplete")
    tick("normalise+forward+decide", t)
    t = time.perf_counter()
    h = dec.cpu(), probs.cpu(), length.cpu(), det.cpu()
    tick("d2h", t)
t = time.perf_counter()
for _ in range(N):
    clf.classify_batch(sigs, ids, {}, 0.9, "deplete")
total = (time.perf_counter() - t) * 1e3 / N
print(json.dumps({"channels": C, "models": M, "classify_batch_ms": round(total, 3),
                  "stages_ms_with_sync_each": {k: round(v / N, 3) for k, v in T.items()},
                  "assessed": int((h[2] > 0).sum())}))
