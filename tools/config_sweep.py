"""Throughput of the other BASELINE.json configurations (they are parity-test shapes, not the
bench line): config 1 (RNA002, 512 x 12,048 already trimmed), config 3 (ragged raw prefixes,
three targets, fused trim + normalise + classify + decide).  Device-timed, inputs resident.

usage: python tools/config_sweep.py [precision]
"""
import json
import logging
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, RaggedBatch, synth   # noqa: E402
from riser_b200.config import shipped_config                                                # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
log = logging.getLogger("sweep")
dev = torch.device("cuda")


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


out = {"precision": prec}
# ---- config 1: RNA002 mRNA model, 512 already-trimmed 12,048-sample chunks
proc = SignalProcessor(Kit.create_from_version("RNA002"))
m = Model(synth.state_dict(0), shipped_config(), log, "mRNA", precision=prec)
clf = BatchedClassifier([m], proc)
X = synth.body_batch(1234, 512, 12048)
batch = RaggedBatch([X[i] for i in range(512)], dev)
start = torch.zeros(512, dtype=torch.int32, device=dev)
length = torch.full((512,), 12048, dtype=torch.int32, device=dev)
ms = timed(lambda: clf.run_windows(batch, start, length, 0.9, "deplete"))
out["config1_512x12048"] = {"ms": round(ms, 3), "reads_per_s": round(512 / ms * 1e3)}

# ---- config 3: ragged raw prefixes (adapter + poly(A) + 2-4 s of body), mRNA / mtRNA / globin
models = [Model(synth.state_dict(synth.TARGET_SEEDS[t]), shipped_config(), log, t, precision=prec)
          for t in ("mRNA", "mtRNA", "globin")]
clf3 = BatchedClassifier(models, proc)
reads = synth.raw_reads(3, 2048, min_body=6024, max_body=12048, frac_no_polya=0.02)
rb = RaggedBatch([s for _, s in reads], dev)
cached = np.full(len(reads), -1, dtype=np.int32)


def step3():
    st, ln, _ = clf3.select_windows(rb, cached)
    return clf3.run_windows(rb, st, ln, 0.9, "deplete")


ms3 = timed(step3)
dec, _ = step3()
n_assessed = int((dec != 4).sum().item())
out["config3_ragged_2048_reads_3_models"] = {"ms": round(ms3, 3), "reads_per_s": round(2048 / ms3 * 1e3),
                                              "assessed": n_assessed,
                                              "model_classifications_per_s": round(3 * n_assessed / ms3 * 1e3)}
print(json.dumps(out))
