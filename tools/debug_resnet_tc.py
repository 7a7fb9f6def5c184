"""Block-by-block comparison of the tensor-core ResNet path (csrc/resnet_tc.cu) with the fp32 CUDA-core path
(csrc/resnet.cu) on the same ragged batch: max |difference| of every block's output over the valid rows.
usage: python tools/debug_resnet_tc.py [basic|bottleneck] [B]"""
import logging
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from riser_b200 import synth                       # noqa: E402
from riser_b200.config import AttrDict             # noqa: E402
from riser_b200.resnet import ResNetModel          # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "basic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.RESNET_CONFIGS[name]
sd = synth.resnet_state_dict(cfg, 0)
log = logging.getLogger("dbg")


def build(tc):
    os.environ["RISER_RESNET_TC"] = tc
    return ResNetModel(sd, AttrDict({"model": "resnet", "resnet": cfg}), log, "mRNA")


m_tc, m_cc = build("1"), build("0")
print(f"tc model: {m_tc.n_tc_fused} fused blocks, {m_tc.n_tc_convs} tc convs, {m_tc.n_cuda_core_convs} CUDA-core convs")
rng = np.random.default_rng(0)
lens = rng.integers(4096, 12049, B).astype(np.int32)
lens[0] = 12048
x = torch.randn(B, 12048, device="cuda")
lt = torch.from_numpy(lens).cuda()
p_tc = m_tc.classify_batch(x, lt, max_len=12048)
p_cc = m_cc.classify_batch(x, lt, max_len=12048)
torch.cuda.synchronize()
n_tc = m_tc._acts[("len", B)].cpu().numpy()


def block_out(model, convs, fused):
    key = id(fused) if fused is not None else id(convs[-1])
    for k, v in model._acts.items():
        if k[0] == key:
            return v
    raise KeyError


j = 2
for bi, ((cv_t, sc_t, f_t), (cv_c, sc_c, f_c)) in enumerate(zip(m_tc.blocks, m_cc.blocks)):
    j += len(cv_t)
    n_valid = n_tc[j - 1]
    a, b = block_out(m_tc, cv_t, f_t), block_out(m_cc, cv_c, f_c)
    worst, scale = 0.0, 0.0
    for r in range(B):
        d = (a[r, :n_valid[r]] - b[r, :n_valid[r]]).abs()
        worst = max(worst, float(d.max()) if d.numel() else 0.0)
        scale = max(scale, float(b[r, :n_valid[r]].abs().max()) if d.numel() else 0.0)
    kind = "fused" if f_t is not None else ("tc convs" if cv_t[0].tc is not None else "cuda cores")
    print(f"block {bi:2d} ({kind:10s}) stride {cv_t[0].stride if len(cv_t) == 2 else cv_t[1].stride} "
          f"C {cv_t[-1].cout:4d} L {a.shape[1]:5d}: max |diff| {worst:.3e} (|act| max {scale:.3e})")
print("probs max |diff|", float((p_tc - p_cc).abs().max()))
