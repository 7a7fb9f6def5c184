"""One-off randomised stress of the preprocessing kernels against the oracle (bit-exact normalise, identical
poly(A) ends) on a few thousand windows of mixed kinds.  usage: python tools/stress_preprocess.py [n_reads] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import preprocess_oracle as pp                       # noqa: E402  (checker only)
from riser_b200 import Kit, SignalProcessor, synth               # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
proc = SignalProcessor(Kit.create_from_version("RNA002"))


def make(k):
    kind = k % 6
    n = int(rng.integers(1, 20000))
    if kind == 0:
        return synth.body(rng, n)
    if kind == 1:
        return synth.body(rng, n, spike_frac=float(rng.choice([0.0, 0.02, 0.1])), mean_dwell=float(rng.choice([3, 9, 40])))
    if kind == 2:
        x = rng.normal(rng.integers(-500, 1500), rng.choice([0.4, 3, 30, 300]), size=n)
    elif kind == 3:
        x = rng.integers(0, int(rng.choice([2, 5, 50])), size=n) * int(rng.choice([1, 7, 400]))
    elif kind == 4:
        x = rng.normal(500, 40, size=n)
        x[rng.random(n) < 0.01] += rng.choice([-3000, 2500, 20000])
    else:
        x = np.cumsum(rng.normal(0, 6, size=n)) + 400
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


sigs = [make(k) for k in range(N)]
start = np.array([int(rng.integers(0, max(1, len(s) // 3))) for s in sigs], dtype=np.int32)
length = np.array([int(rng.integers(0, len(s) - st + 1)) for s, st in zip(sigs, start)], dtype=np.int32)
t0 = time.time()
out, _, stats = proc.mad_normalise_batch(sigs, start=start, length=length, return_stats=True)
out, stats = out.cpu().numpy(), stats.cpu().numpy()
bad = 0
for b, s in enumerate(sigs):
    n = int(length[b])
    if n == 0:
        continue
    w = s[start[b]:start[b] + n]
    want = np.asarray(pp.mad_normalise(w), dtype=np.float64).astype(np.float32)
    if not np.array_equal(out[b, :n], want):
        bad += 1
        print("normalise mismatch", b, n, int(np.flatnonzero(out[b, :n] != want)[0]))
print(f"normalise: {N} windows, {bad} mismatches, {time.time() - t0:.1f}s")
raws = [synth.raw_read(rng, int(rng.integers(2000, 16000)), polya=bool(rng.random() < 0.85))[0][:int(rng.integers(500, 24000))]
        for _ in range(N // 2)] + [s for s in sigs[:N // 2]]
ends = proc.get_polyA_end_batch(raws)
want = np.array([(-1 if pp.polya_end(s) is None else pp.polya_end(s)) for s in raws], dtype=np.int32)
print(f"poly(A): {len(raws)} prefixes, {int((ends != want).sum())} mismatches, found in {(want > 0).mean() * 100:.0f}%")
sys.exit(1 if bad or (ends != want).any() else 0)
