"""Live-stream simulation (BASELINE config 5): C channels each carrying a read whose prefix
grows by 1 s of samples per poll (AccumulatingCache semantics, riser/client.py:29-31,44); the
batched SequencerControl classifies every poll's batch.  Reports p50 / p99 of the time from
"batch in hand" to "decisions on host" (the ReadUntil decision budget is ~1 s).

usage: python tools/live_sim.py [channels] [polls] [kit] [n_models]
"""
import json
import logging
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, Model, SequencerControl, synth, sim     # noqa: E402
from riser_b200.config import shipped_config                                          # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
polls = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kit = sys.argv[3] if len(sys.argv) > 3 else "RNA002"
n_models = int(sys.argv[4]) if len(sys.argv) > 4 else 1
log = logging.getLogger("live")
hz = Kit.create_from_version(kit).sampling_hz
targets = ["mRNA", "mtRNA", "globin"][:n_models]
models = [Model(synth.state_dict(synth.TARGET_SEEDS[t]), shipped_config(), log, t) for t in targets]
proc = SignalProcessor(Kit.create_from_version(kit))
t0 = time.time()
reads = synth.raw_reads(7, min(2 * C, 2048), min_body=14000, max_body=20000, frac_no_polya=0.05)
print(f"generated {len(reads)} reads in {time.time() - t0:.1f}s", file=sys.stderr)
client = sim.LiveSimClient(reads, C, chunk=hz, n_polls=polls, first_len=hz)
out = tempfile.mkdtemp() + "/live"
control = SequencerControl(client, models, proc, log, out, warm_up_batches=(C,))
control.start()
control.target("deplete", 1, 0.9)
control.finish()
lat = np.array(control.batch_latencies) * 1e3
sizes = control.batch_sizes if hasattr(control, "batch_sizes") else []
print(json.dumps({"channels": C, "kit": kit, "models": targets, "polls": len(lat),
                  "latency_ms_first10": [round(float(x), 2) for x in lat[:10]],
                  "batch_size_median": int(np.median(sizes)) if sizes else 0,
                  "assessed_rows": sum(1 for _ in open(out + ".csv")) - 1,
                  "p50_ms_after_warmup": round(float(np.median(lat[5:])), 2) if len(lat) > 8 else None,
                  "p99_ms_after_warmup": round(float(np.percentile(lat[5:], 99)), 2) if len(lat) > 8 else None,
                  "max_ms_after_warmup": round(float(lat[5:].max()), 2) if len(lat) > 8 else None,
                  "rejected": len(client.unblocked), "finished": len(client.finished)}))
