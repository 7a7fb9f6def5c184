"""In-situ per-layer times of the conv stack: every layer of a real forward bracketed by CUDA events on the launching
stream (riser_forward_stage 16 + i), the forwards replayed back to back like bench.py's steps (L2 flushed between
them), so the clocks / power state are those of the real step -- unlike an ncu launch list, whose launches are
serialised and cold.  Prints the median per layer over the timed forwards.
usage: python tools/layer_events.py [B=4096] [L=16000] [precision=3] [iters=12] [tag]"""
import json
import logging
import os
import sys

import torch

sys.path.insert(0, ".")
from riser_b200.config import AttrDict           # noqa: E402
from riser_b200 import Model, SignalProcessor, Kit, RaggedBatch, synth, _lib   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 3
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 12
tag = sys.argv[5] if len(sys.argv) > 5 else ""
CFG = AttrDict({"cnn": {"n_layers": 12, "depth": 1, "channels": synth.CHANNELS, "kernels": [3] * 12,
                        "n_classes": 2, "classifier": "gap_fc"}})
model = Model(synth.state_dict(0), CFG, logging.getLogger("t"), "mRNA", precision=prec)
proc = SignalProcessor(Kit.create_from_version("RNA004"))
pool = synth.body_batch(1, 256, L)
batch = RaggedBatch([pool[i % 256] for i in range(B)], torch.device("cuda"))
x = torch.zeros(B, (L + 3) & ~3, device="cuda")
x, lens = proc.mad_normalise_batch(batch, out=x)
probs = torch.empty(B, 2, device="cuda")
model.classify_batch(x, lens, max_len=L, probs=probs)
plan = model.plan(B, L)
lib = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = 12


def stage(st):
    _lib.check(lib.riser_forward_stage(plan._handle, st, _lib.ptr(x), x.stride(0), _lib.ptr(lens), _lib.ptr(probs),
                                       None, _lib.stream_ptr()), "stage")


first = 2 if plan.fused_layer0 else 1        # fused: layer 1's launch computes layer 0 too
rows = []
for it in range(iters + 3):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 2)]
    ev[0].record()
    stage(0)
    for i in range(1, n):
        ev[i].record()
        stage(16 + i)
    ev[n].record()
    stage(2)
    ev[n + 1].record()
    ev[n + 1].synchronize()
    if it >= 3:
        rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(1, n)] + [ev[0].elapsed_time(ev[n + 1])])
med = [sorted(c)[len(c) // 2] for c in zip(*rows)]
names = [plan.layer_kernel(i).replace("conv_", "").replace("_kernel", "") for i in range(1, n)]
out = {"tag": tag, "env": {k: v for k, v in os.environ.items() if k.startswith("RISER_")}, "B": B, "L": L,
       "precision": prec, "layer_ms": dict(zip([f"{i}:{nm}" for i, nm in zip(range(1, n), names)], [round(v, 4) for v in med[:-1]])),
       "conv_ms": round(sum(med[:-1]), 4), "forward_ms": round(med[-1], 4)}
print(json.dumps(out))
