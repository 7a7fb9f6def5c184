"""Sustained (power-capped) throughput of the captured step: the CUDA graph of one batch replayed back to back
for a few seconds, ms/step over the second half, SM clock and board power sampled through NVML meanwhile.
usage: python tools/sustained.py [seconds] [precision]      (RISER_B200_LIB selects a kernel build)"""
import json
import logging
import os
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, FixedBatchPipeline, synth, model as rmodel  # noqa: E402
from riser_b200.config import shipped_config                                                                       # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
prec = int(sys.argv[2]) if len(sys.argv) > 2 else rmodel.DEFAULT_PRECISION
B, L = 4096, 16000
mdl = Model(synth.state_dict(0), shipped_config(), logging.getLogger("p"), "mRNA", precision=prec)
clf = BatchedClassifier([mdl], SignalProcessor(Kit.create_from_version("RNA004")))
clf.max_len, clf.ld = L, (L + 3) & ~3
pool = synth.body_batch(100, 256, L)
host = torch.empty(B, L, dtype=torch.int16).pin_memory()
for i in range(B):
    host.numpy()[i] = pool[i % len(pool)]
pipe = FixedBatchPipeline(clf, B, L, 0.9, "deplete")
pipe.result(pipe.submit(host))
g = pipe.slots[0]["graph"]

samples, stop = [], False


def sample():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.02)


th = threading.Thread(target=sample, daemon=True)
th.start()
n = max(int(secs / 0.0075), 20)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for k in range(n):
    if k == n // 2:
        ev[1].record()
    g.replay()
ev[2].record()
ev[2].synchronize()
stop = True
th.join(timeout=1)
half = samples[len(samples) // 2:]
print(json.dumps({"lib": os.path.basename(os.environ.get("RISER_B200_LIB", "libriser_b200.so")), "precision": prec,
                  "steps": n, "ms_per_step_first_half": ev[0].elapsed_time(ev[1]) / (n // 2),
                  "ms_per_step_second_half": ev[1].elapsed_time(ev[2]) / (n - n // 2),
                  "reads_per_s_sustained": B / (ev[1].elapsed_time(ev[2]) / (n - n // 2)) * 1e3,
                  "sm_mhz_median": sorted(s[0] for s in half)[len(half) // 2] if half else None,
                  "power_w_median": sorted(s[1] for s in half)[len(half) // 2] if half else None}))
