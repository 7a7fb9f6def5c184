"""ConvNet shapes beyond the shipped one (riser/nets/cnn.py:8-65: depth > 1, kernel sizes other than 3, 'gap' head,
n_classes != 2) through riser_b200.Model, against golden probabilities of the reference's own ConvNet module
(tests/golden/make_golden_generic.py)."""
import logging
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as pp
from riser_b200 import Model, synth
from riser_b200.config import AttrDict

pytestmark = pytest.mark.gpu
LOG = logging.getLogger("test")


def _inputs(g):
    bodies = synth.ragged_bodies(int(g["seed_reads"]), len(g["lengths"]), 700, 5000)
    assert [len(b) for b in bodies] == g["lengths"].tolist()
    return [np.asarray(pp.mad_normalise(b), dtype=np.float64) for b in bodies]


@pytest.mark.parametrize("name", sorted(synth.GENERIC_CNN_CONFIGS))
def test_generic_convnet_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "convnet_generic.npz"))
    cfg = synth.GENERIC_CNN_CONFIGS[name]
    normed = _inputs(g)
    model = Model(synth.generic_cnn_state_dict(cfg, 0), AttrDict({"model": "cnn", "cnn": cfg}), LOG, "mRNA")
    assert model._generic is not None
    # k = 3 convolutions past the first run on tcgen05 (riser_res_tc), the others on the fp32 CUDA-core kernel
    n_k3 = sum(cfg["depth"] - (1 if i == 0 else 0) for i, k in enumerate(cfg["kernels"]) if k == 3)
    assert model._generic.n_tc == n_k3 and model._generic.n_tc + model._generic.n_cuda_core == cfg["depth"] * cfg["n_layers"]
    ld = 5000
    x = torch.full((len(normed), ld), 7.0)               # poison the padding: must never be read as signal
    for b, v in enumerate(normed):
        x[b, :len(v)] = torch.from_numpy(v).float()
    lens = torch.tensor([len(v) for v in normed], dtype=torch.int32)
    probs = model.classify_batch(x.cuda(), lens.cuda(), max_len=ld).cpu().numpy()
    want = g[f"probs_{name}"]
    assert probs.shape == want.shape == (len(normed), cfg["n_classes"])
    err = np.abs(probs - want).max()
    assert err < 1e-3, err                                # north star: 1e-3 absolute on probabilities
    assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-5)
    # riser/model.py:22-28, one read at its own length
    one = model.classify(normed[3]).cpu().numpy()
    assert np.abs(one - want[3]).max() < 1e-3
    print(f"{name}: max |dp| {err:.2e} over {probs.size} probabilities")


def test_generic_convnet_short_reads_and_refusals():
    cfg = synth.GENERIC_CNN_CONFIGS["k5_fc"]
    model = Model(synth.generic_cnn_state_dict(cfg, 0), AttrDict({"model": "cnn", "cnn": cfg}), LOG, "mRNA")
    x = torch.zeros(3, 64, device="cuda")
    lens = torch.tensor([64, 16, 0], dtype=torch.int32, device="cuda")    # 5 pools: 16 samples leave nothing
    p = model.classify_batch(x, lens, max_len=64).cpu().numpy()
    assert np.isfinite(p[0]).all() and np.isnan(p[1]).all() and np.isnan(p[2]).all()
    assert model.classify_batch(x[:0], lens[:0], max_len=64).shape == (0, 2)
    with pytest.raises(RuntimeError):
        model.classify(np.zeros(8))
    bad = dict(cfg, kernels=[3, 4, 3, 5, 3])
    with pytest.raises(NotImplementedError):
        Model(synth.generic_cnn_state_dict(bad, 0), AttrDict({"model": "cnn", "cnn": bad}), LOG, "mRNA")
    with pytest.raises(NotImplementedError):
        Model(synth.generic_cnn_state_dict(cfg, 0), AttrDict({"model": "cnn", "cnn": dict(cfg, classifier="fc")}), LOG, "mRNA")
    sd = synth.generic_cnn_state_dict(cfg, 0)
    sd.pop("layers.2.0.bias")
    with pytest.raises(RuntimeError, match="missing keys"):
        Model(sd, AttrDict({"model": "cnn", "cnn": cfg}), LOG, "mRNA")
