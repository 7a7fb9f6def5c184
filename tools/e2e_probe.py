"""Where the gap between bench.py's HBM-resident `value` and its host-buffer `e2e` comes from: the same
FixedBatchPipeline run for several step counts (pipeline fill amortises), with and without the H2D copy.
usage: python tools/e2e_probe.py [B] [L]"""
import json
import logging
import sys

import torch

sys.path.insert(0, ".")
from riser_b200 import Kit, SignalProcessor, Model, BatchedClassifier, FixedBatchPipeline, synth   # noqa: E402
from riser_b200.config import shipped_config                                                       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
mdl = Model(synth.state_dict(0), shipped_config(), logging.getLogger("p"), "mRNA")
clf = BatchedClassifier([mdl], SignalProcessor(Kit.create_from_version("RNA004")))
clf.max_len, clf.ld = L, (L + 3) & ~3
pool = synth.body_batch(100, 256, L)
host = torch.empty(B, L, dtype=torch.int16).pin_memory()
for i in range(B):
    host.numpy()[i] = pool[i % len(pool)]
hosts = [host, host.clone().pin_memory()]
pipe = FixedBatchPipeline(clf, B, L, 0.9, "deplete")
for k in range(3):
    pipe.result(pipe.submit(hosts[k % 2]))
out = {}
for steps in (10, 30, 100):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pipe.copy_stream.wait_event(a)
    tickets = []
    for k in range(steps):
        tickets.append(pipe.submit(hosts[k % 2]))
        if k >= 1:
            pipe.result(tickets[k - 1])
    pipe.result(tickets[-1])
    b.record()
    b.synchronize()
    out[f"e2e_{steps}_steps_ms_per_step"] = a.elapsed_time(b) / steps
# the captured graph alone, back to back, inputs resident
slot = pipe.slots[0]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30):
    slot["graph"].replay()
b.record()
b.synchronize()
out["graph_only_ms_per_step"] = a.elapsed_time(b) / 30
# H2D alone
a.record()
for k in range(10):
    slot["batch"].sig[:B * L].view(B, L).copy_(hosts[k % 2], non_blocking=True)
b.record()
b.synchronize()
out["h2d_only_ms"] = a.elapsed_time(b) / 10
print(json.dumps(out))
