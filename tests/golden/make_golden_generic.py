#!/usr/bin/env python
"""Golden probabilities for ConvNet shapes beyond the shipped one (riser/nets/cnn.py:8-65: depth > 1, other kernel
sizes, 'gap' head, n_classes != 2), from the REAL reference module (imported read-only through oracle/refshim.py).

Run once in the build container:   python tests/golden/make_golden_generic.py
Writes tests/golden/convnet_generic.npz (seeds, lengths and the reference's outputs; weights and inputs are
regenerated from seeds by riser_b200/synth.py)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim                      # noqa: E402
from oracle import preprocess_oracle as pp      # noqa: E402
from riser_b200 import synth                    # noqa: E402

ref = refshim.load()
SEED_READS, N_READS = 4242, 10
bodies = synth.ragged_bodies(SEED_READS, N_READS, 700, 5000)
normed = [np.asarray(pp.mad_normalise(b), dtype=np.float64) for b in bodies]
out = {"seed_reads": np.array(SEED_READS), "lengths": np.array([len(b) for b in bodies])}
torch.set_num_threads(4)
for name, cfg in synth.GENERIC_CNN_CONFIGS.items():
    m = ref.ConvNet(refshim.AttrDict(cfg))
    m.load_state_dict(synth.generic_cnn_state_dict(cfg, 0))
    m.eval()
    with torch.no_grad():      # riser/model.py:22-28, one read at a time at its own length
        probs = np.stack([F.softmax(m(torch.from_numpy(x).float()[None]).reshape(1, -1), dim=1)[0].numpy() for x in normed])
    out[f"probs_{name}"] = probs.astype(np.float32)
    print(name, "p range", probs.min(), probs.max())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "convnet_generic.npz"), **out)
