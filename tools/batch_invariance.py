"""Are a read's probabilities independent of the batch it is classified in?  Classifies a batch of raw reads whole and
as two halves (r mod 2) on ONE GPU through BatchedClassifier.classify_batch and compares bit for bit -- what
tools/shard_check.py asserts across ranks.  usage: python tools/batch_invariance.py [n_reads] [env=val ...]"""
import logging
import os
import sys

import numpy as np

for kv in sys.argv[2:]:
    k, v = kv.split("=")
    os.environ[k] = v
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, synth      # noqa: E402
from riser_b200.config import shipped_config                                      # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 96
log = logging.getLogger("inv")
models = [Model(synth.state_dict(synth.TARGET_SEEDS[t]), shipped_config(), log, t) for t in ["mRNA", "mtRNA"]]
clf = BatchedClassifier(models, SignalProcessor(Kit.create_from_version("RNA002")))
reads = synth.raw_reads(21, n_reads, min_body=3000, max_body=15000, frac_no_polya=0.2)
sigs, ids = [s for _, s in reads], [r for r, _ in reads]
whole = clf.classify_batch(sigs, ids, {}, 0.9, "deplete")
again = clf.classify_batch(sigs, ids, {}, 0.9, "deplete")
print("same batch twice: identical" if np.array_equal(whole.p_on, again.p_on, equal_nan=True) else "same batch twice: DIFFERENT")
for half in (0, 1):
    idx = np.arange(half, n_reads, 2)
    part = clf.classify_batch([sigs[i] for i in idx], [ids[i] for i in idx], {}, 0.9, "deplete")
    ok = whole.sig_len[idx] > 0
    d = np.abs(part.p_on.astype(np.float64) - whole.p_on[idx].astype(np.float64))
    d[~ok] = 0
    bad = np.flatnonzero(d.max(axis=1) > 0)
    print(f"half {half}: {len(idx)} reads, {int(ok.sum())} assessed, {len(bad)} differ, max |dp| {d.max():.3e}; "
          f"lengths equal {np.array_equal(part.sig_len, whole.sig_len[idx])}; decisions equal "
          f"{np.array_equal(part.decisions, whole.decisions[idx])}")
    for b in bad[:8]:
        print(f"   read {idx[b]} (position {b} in the half): window {whole.sig_len[idx[b]]}, |dp| {d[b].tolist()}")
