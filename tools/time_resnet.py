"""Device timing of the ResNet variant (riser_b200/resnet.py) on a fixed-length batch.
usage: python tools/time_resnet.py [B] [L] [basic|bottleneck]"""
import json
import logging
import sys

import torch

sys.path.insert(0, ".")
from riser_b200 import synth                       # noqa: E402
from riser_b200.config import AttrDict             # noqa: E402
from riser_b200.resnet import ResNetModel          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = int(sys.argv[2]) if len(sys.argv) > 2 else 12048
name = sys.argv[3] if len(sys.argv) > 3 else "basic"
cfg = synth.RESNET_CONFIGS[name]
model = ResNetModel(synth.resnet_state_dict(cfg, 0), AttrDict({"model": "resnet", "resnet": cfg}), logging.getLogger("t"), "mRNA")
x = torch.randn(B, L, device="cuda")
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
for _ in range(2):
    p = model.classify_batch(x, lens, max_len=L)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    p = model.classify_batch(x, lens, max_len=L)
b.record()
b.synchronize()
ms = a.elapsed_time(b) / 3
print(json.dumps({"net": "resnet-" + name, "B": B, "L": L, "ms": ms, "reads_per_s": B / ms * 1e3,
                  "p_on_first": p[:3, 1].tolist()}))
