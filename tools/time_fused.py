"""CUDA-event timing of the fused layers 0+1 launch alone (riser_forward_stage ... via Model.time_* hooks are
not exposed, so this times stage 0+1 of the forward with the conv stack cut after layer 1 by a 2-layer model).
usage: python tools/time_fused.py [precision] [B] [L]"""
import logging
import sys

import torch

sys.path.insert(0, ".")
from riser_b200.config import AttrDict           # noqa: E402
from riser_b200 import Model, SignalProcessor, Kit, RaggedBatch, synth, _lib   # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(sys.argv[3]) if len(sys.argv) > 3 else 16000
sd = synth.state_dict(0)
CFG = AttrDict({"cnn": {"n_layers": 12, "depth": 1, "channels": synth.CHANNELS, "kernels": [3] * 12,
                        "n_classes": 2, "classifier": "gap_fc"}})
model = Model(sd, CFG, logging.getLogger("t"), "mRNA", precision=prec)
proc = SignalProcessor(Kit.create_from_version("RNA004"))
pool = synth.body_batch(1, 256, L)
batch = RaggedBatch([pool[i % 256] for i in range(B)], torch.device("cuda"))
x = torch.zeros(B, (L + 3) & ~3, device="cuda")
x, lens = proc.mad_normalise_batch(batch, out=x)
probs = torch.empty(B, 2, device="cuda")
model.classify_batch(x, lens, max_len=L, probs=probs)
plan = model.plan(B, L)
lib = _lib.lib()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def stage(st):
    _lib.check(lib.riser_forward_stage(plan._handle, st, _lib.ptr(x), x.stride(0), _lib.ptr(lens), _lib.ptr(probs),
                                       None, _lib.stream_ptr()), "stage")


# stage 1 launches layers 1..11; time the whole stage and subtract nothing: use ncu for the split.  Here we
# time the full forward (stage 0 + 1) -- the fused launch is the first kernel of stage 1.
ts = []
for _ in range(6):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    stage(0)
    stage(1)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
print(f"forward stages 0+1: min {min(ts[1:]):.3f} ms  median {sorted(ts[1:])[len(ts[1:]) // 2]:.3f} ms")
