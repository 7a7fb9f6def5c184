"""Quick device timing of riser_forward (and normalise) for a fixed-length batch.
usage: python tools/time_forward.py [B] [L] [precision] [iters]"""
import logging
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from riser_b200.config import AttrDict           # noqa: E402
from riser_b200 import Model, SignalProcessor, Kit, RaggedBatch, synth   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = int(sys.argv[2]) if len(sys.argv) > 2 else 12048
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
CFG = AttrDict({"cnn": {"n_layers": 12, "depth": 1, "channels": synth.CHANNELS, "kernels": [3] * 12,
                        "n_classes": 2, "classifier": "gap_fc"}})
model = Model(synth.state_dict(0), CFG, logging.getLogger("t"), "mRNA", precision=prec)
proc = SignalProcessor(Kit.create_from_version("RNA002"))
pool = synth.body_batch(1, min(B, 256), L)
sigs = [pool[i % len(pool)] for i in range(B)]
batch = RaggedBatch(sigs, torch.device("cuda"))
ld = (L + 3) & ~3
x = torch.zeros(B, ld, device="cuda")
x, lens = proc.mad_normalise_batch(batch, out=x)
probs = torch.empty(B, 2, device="cuda")
model.classify_batch(x, lens, max_len=L, probs=probs)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for name, fn in (("normalise", lambda: proc.mad_normalise_batch(batch, out=x)),
                 ("forward", lambda: model.classify_batch(x, lens, max_len=L, probs=probs))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    extra = ""
    if name == "normalise":
        extra = f"  {6 * L * B / ms / 1e6:.0f} GB/s algorithmic"
    else:
        flops = 435.1e6 if L == 12048 else None
        if flops:
            extra = f"  {flops * B / ms / 1e9:.1f} TFLOP/s algorithmic"
    print(f"{name}: {ms:.3f} ms / batch of {B} x {L}  -> {B / ms * 1e3:.0f} reads/s{extra}")
print("p_on[:4]", probs[:4, 1].tolist(), "layer0 ms", model.time_layer0(x, lens, L))
