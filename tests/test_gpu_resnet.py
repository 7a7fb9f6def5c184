"""GPU parity of the ResNet variant (csrc/resnet.cu + riser_b200/resnet.py) against golden
probabilities produced by the reference's own nets/resnet.py ResNet module, and the oracle."""
import logging
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as pp
from oracle import resnet_oracle as ro
from riser_b200 import synth
from riser_b200.config import AttrDict
from riser_b200.resnet import ResNetModel

pytestmark = pytest.mark.gpu
LOG = logging.getLogger("test")


@pytest.mark.parametrize("tc", ["1", "0"])
@pytest.mark.parametrize("name", ["basic", "bottleneck"])
def test_resnet_ragged_batch_matches_reference(golden_dir, monkeypatch, name, tc):
    """tc = "1": residual blocks on tcgen05 (csrc/resnet_tc.cu; BasicBlocks fused, bottleneck convs one launch each);
    tc = "0": everything on the fp32 CUDA-core kernels (csrc/resnet.cu).  Same goldens: the reference's own module."""
    monkeypatch.setenv("RISER_RESNET_TC", tc)
    g = np.load(os.path.join(golden_dir, "resnet_probs.npz"))
    cfg = synth.RESNET_CONFIGS[name]
    bodies = synth.ragged_bodies(int(g["seed"]), 24, 4096, 12048)
    normed = [pp.mad_normalise(b) for b in bodies]
    model = ResNetModel(synth.resnet_state_dict(cfg, 0), AttrDict({"model": "resnet", "resnet": cfg}), LOG, "mRNA")
    if tc == "1":
        # (the bottleneck configuration's 128 -> 256 stride-2 shortcut needs 270 KB of shared memory: CUDA cores)
        assert model.n_cuda_core_convs <= (1 if name == "bottleneck" else 0) and model.n_tc_fused + model.n_tc_convs > 0
        assert (model.n_tc_fused > 0) == (name == "basic")
    else:
        assert model.n_tc_fused == model.n_tc_convs == 0
    n = np.array([len(x) for x in normed], dtype=np.int32)
    x = torch.full((len(normed), 12048), 9.0)            # poison the padding: must never be read as signal
    for b, v in enumerate(normed):
        x[b, :len(v)] = torch.from_numpy(np.asarray(v, dtype=np.float64)).float()
    probs = model.classify_batch(x.cuda(), torch.from_numpy(n).cuda(), max_len=12048).cpu().numpy()
    want = g[f"probs_{name}"]
    err = np.abs(probs - want).max()
    print(f"{name} tc={tc}: max |dp| = {err:.2e}")
    assert err < 1e-3, err                                # the north star's bar
    assert err < (2e-4 if tc == "1" else 5e-5), err       # hi + lo operand planes / fp32 CUDA cores: far inside it
    # a second call reuses the activation buffers: same result (stale rows beyond a read's length are never read)
    again = model.classify_batch(x.cuda(), torch.from_numpy(n).cuda(), max_len=12048).cpu().numpy()
    assert np.array_equal(again, probs)
    # single-read drop-in call
    p = model.classify(normed[3])
    assert abs(p[1].item() - want[3, 1]) < (2e-4 if tc == "1" else 5e-5)
    assert abs(p[1].item() - ro.classify(synth.resnet_state_dict(cfg, 0), cfg, normed[3])[1].item()) < 2e-4


def test_resnet_behind_the_batched_pipeline():
    """A ResNetModel is interchangeable with Model in BatchedClassifier (trim -> normalise ->
    classify -> decide); decisions follow control.py:75-82 on the oracle's probabilities."""
    from oracle import control_oracle as ctl
    from riser_b200 import Kit, SignalProcessor, BatchedClassifier
    cfg = synth.RESNET_CONFIGS["basic"]
    sd = synth.resnet_state_dict(cfg, 0)
    model = ResNetModel(sd, AttrDict({"model": "resnet", "resnet": cfg}), LOG, "mRNA")
    proc = SignalProcessor(Kit.create_from_version("RNA002"))
    clf = BatchedClassifier([model], proc)
    reads = synth.raw_reads(17, 16, min_body=9000, max_body=14000, frac_no_polya=0.2)
    res = clf.classify_batch([s for _, s in reads], [r for r, _ in reads], {}, 0.9, "enrich")
    cache = {}
    for i, (rid, sig) in enumerate(reads):
        window, _ = pp.select_window(sig, rid, cache, "RNA002")
        if window is None:
            assert res.decisions[i] == ctl.SKIPPED
            continue
        p = ro.classify(sd, cfg, pp.mad_normalise(window))
        assert abs(res.p_on[i, 0] - p[1].item()) < 3e-4 and res.sig_len[i] == len(window)
        want = ctl.decide([p[1]], [p[0]], len(window), 12048, 0.9, "enrich")
        assert res.decisions[i] == want or abs(p[1].item() - 0.9) < 1e-3 or abs(p[0].item() - 0.9) < 1e-3
