"""Decision / probability agreement of the whole path with the oracle's serial restatement of control.py:31-97
on N raw synthetic reads and the three targets (mRNA, mtRNA, globin), both modes.  The oracle leg runs on the
CPU (~25 ms per read and model).  usage: python tools/parity_sweep.py [n_reads] [seed]"""
import json
import logging
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import control_oracle as ctl                                                   # noqa: E402 (checker)
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, synth              # noqa: E402
from riser_b200.config import shipped_config                                               # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 77
log = logging.getLogger("parity")
targets = ["mRNA", "mtRNA", "globin"]
states = [synth.state_dict(synth.TARGET_SEEDS[t]) for t in targets]
models = [Model(s, shipped_config(), log, t) for s, t in zip(states, targets)]
proc = SignalProcessor(Kit.create_from_version("RNA002"))
clf = BatchedClassifier(models, proc)
reads = synth.raw_reads(seed, N, min_body=3000, max_body=16000, frac_no_polya=0.1)
out = {"reads": N, "models": targets}
for mode in ("deplete", "enrich"):
    res = clf.classify_batch([s for _, s in reads], [r for r, _ in reads], {}, 0.9, mode)
    t0 = time.time()
    dec, p_on, p_off, sig_len, _ = ctl.run_batch(reads, states, "RNA002", {}, 0.9, mode)
    ok = sig_len > 0
    dp = np.abs(res.p_on[ok] - p_on[ok])
    near = (np.abs(p_on - 0.9).min(axis=1) <= 1e-3) | (np.abs(p_off - 0.9).min(axis=1) <= 1e-3)
    out[mode] = {"assessed": int(ok.sum()), "skipped": int((~ok).sum()), "sig_len_equal": bool(np.array_equal(res.sig_len, sig_len)),
                 "max_abs_dp": float(dp.max()), "mean_abs_dp": float(dp.mean()),
                 "decisions_equal": int((res.decisions == dec).sum()), "decisions_differ": int((res.decisions != dec).sum()),
                 "differ_outside_1e-3_of_threshold": int(((res.decisions != dec) & ~near).sum()),
                 "reads_within_1e-3_of_threshold": int((near & ok).sum()),
                 "decision_histogram": np.bincount(dec, minlength=5).tolist(), "oracle_seconds": round(time.time() - t0, 1)}
print(json.dumps(out))
