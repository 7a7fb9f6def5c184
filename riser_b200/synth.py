"""Seeded synthetic nanopore squiggles and synthetic ConvNet weights.

The reference ships neither data nor trained weights (its six ``.pth`` files are
absent), so tests and ``bench.py`` run on synthetic inputs built here.  The
recipe is the one fixed in SURVEY.md 8(d): adapter, poly(A) and body segments
with ADC-like levels, sparse spikes to exercise outlier smoothing, a few reads
without a detectable poly(A) and a few constant reads (MAD == 0).  Everything
is drawn from ``numpy.random.Generator(PCG64(seed))`` so the same seed gives
the same bytes on every machine.
"""
import os

import numpy as np

CHANNELS = [20, 30, 45, 67, 100, 150, 225, 337, 505, 757, 1135, 1702]  # riser/model/*.yaml:9
TARGET_SEEDS = {"mRNA": 0, "mtRNA": 1, "globin": 2}
_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _levels(rng, n, mean_dwell=9.0, level_mu=500.0, level_sd=60.0):
    """Piece-wise constant current levels with geometric dwell times."""
    n_seg = int(n / mean_dwell * 1.3) + 16
    while True:
        dwell = rng.geometric(1.0 / mean_dwell, size=n_seg)
        if int(dwell.sum()) >= n:
            break
        n_seg *= 2
    lev = rng.normal(level_mu, level_sd, size=n_seg)
    return np.repeat(lev, dwell)[:n]


def body(rng, n, spike_frac=0.005, mean_dwell=9.0, level_mu=500.0, level_sd=60.0):
    """Post-poly(A) RNA body signal, int16, n samples, with sparse +-400 spikes
    (singletons and runs of 2..5) that become |z| > 3.5 outliers."""
    x = _levels(rng, n, mean_dwell, level_mu, level_sd) + rng.normal(0.0, 8.0, size=n)
    n_spikes = int(n * spike_frac / 2)
    if n_spikes and n > 8:
        starts = rng.integers(0, n, size=n_spikes)
        runs = np.where(rng.random(n_spikes) < 0.7, 1, rng.integers(2, 6, size=n_spikes))
        signs = np.where(rng.random(n_spikes) < 0.5, -400.0, 400.0)
        for s, r, g in zip(starts, runs, signs):
            x[s:s + r] += g
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def raw_read(rng, n_body, polya=True, mean_dwell=9.0, level_mu=500.0, level_sd=60.0):
    """A whole read prefix as the sequencer delivers it: adapter, poly(A), body.
    Returns (int16 signal, index of the first body sample)."""
    n_ad = int(rng.integers(1500, 3501))
    adapter = rng.normal(420.0, 45.0, size=n_ad)
    if polya:
        n_pa = int(rng.integers(1500, 3001))
        tail = rng.normal(620.0, 8.0, size=n_pa)
    else:  # no raised, quiet segment: get_polyA_end never fires -> fixed trim path
        n_pa = int(rng.integers(1500, 3001))
        tail = rng.normal(430.0, 45.0, size=n_pa)
    head = np.clip(np.rint(np.concatenate([adapter, tail])), -32768, 32767).astype(np.int16)
    return np.concatenate([head, body(rng, n_body, mean_dwell=mean_dwell,
                                      level_mu=level_mu, level_sd=level_sd)]), n_ad + n_pa


def body_batch(seed, n_reads, length, two_class=True):
    """[n_reads, length] int16 already-trimmed bodies (BASELINE config 1 / 2 shape).
    Reads alternate between two signal 'classes' (different dwell / level spread)
    so a fitted head has something to separate."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_reads, length), dtype=np.int16)
    for r in range(n_reads):
        if two_class and (r & 1):
            out[r] = body(rng, length, mean_dwell=14.0, level_sd=75.0)
        else:
            out[r] = body(rng, length)
    return out


def ragged_bodies(seed, n_reads, min_len=4096, max_len=12048, frac_max=0.25):
    """List of int16 bodies with lengths U{min_len..max_len} plus a point mass at
    max_len (BASELINE config 3 shape)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for r in range(n_reads):
        n = max_len if rng.random() < frac_max else int(rng.integers(min_len, max_len + 1))
        if r & 1:
            out.append(body(rng, n, mean_dwell=14.0, level_sd=75.0))
        else:
            out.append(body(rng, n))
    return out


def raw_reads(seed, n_reads, min_body=2000, max_body=16000, frac_no_polya=0.02, frac_const=0.0):
    """List of (read_id, int16 raw prefix) for the full trim+normalise+classify path."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for r in range(n_reads):
        nb = int(rng.integers(min_body, max_body + 1))
        u = rng.random()
        if u < frac_const:
            sig = np.full(nb + 4000, 500, dtype=np.int16)
        else:
            sig, _ = raw_read(rng, nb, polya=not (u < frac_const + frac_no_polya),
                              mean_dwell=14.0 if (r & 1) else 9.0,
                              level_sd=75.0 if (r & 1) else 60.0)
        out.append((f"read-{seed}-{r:06d}", sig))
    return out


def conv_weights(seed, channels=CHANNELS, kernel=3):
    """He-normal (fan-in) conv weights and N(0, 0.05) biases with the reference's
    state-dict key names / shapes (layers.{i}.0.weight [Cout, Cin, 3], .bias [Cout])."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    state = {}
    cin = 1
    for i, cout in enumerate(channels):
        std = np.sqrt(2.0 / (cin * kernel))
        state[f"layers.{i}.0.weight"] = rng.normal(0.0, std, size=(cout, cin, kernel)).astype(np.float32)
        state[f"layers.{i}.0.bias"] = rng.normal(0.0, 0.05, size=(cout,)).astype(np.float32)
        cin = cout
    return state


def head_weights(seed):
    """The fitted 2 x 1702 linear head for ``conv_weights(seed)``; produced once by
    tests/golden/make_golden.py (logistic fit on reference features) and committed
    under riser_b200/data/."""
    path = os.path.join(_DATA_DIR, f"synth_head_seed{seed}.npz")
    with np.load(path) as z:
        return {"classifier.2.weight": z["weight"].astype(np.float32),
                "classifier.2.bias": z["bias"].astype(np.float32)}


def state_dict(seed, as_torch=True):
    """Full synthetic state-dict (26 tensors, 10,447,564 parameters) loadable by the
    reference's ``ConvNet.load_state_dict`` (riser/model.py:19)."""
    sd = conv_weights(seed)
    sd.update(head_weights(seed))
    if as_torch:
        import torch
        sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    return sd


def save_state_dict(seed, path):
    import torch
    torch.save(state_dict(seed), path)
    return path


# ---------------------------------------------------------------- ConvNet shapes beyond the shipped one
GENERIC_CNN_CONFIGS = {
    # riser/nets/cnn.py in full: depth > 1, kernel sizes other than 3, 'gap' head, n_classes != 2 (the reference
    # ships only the 12 x k3 depth-1 'gap_fc' two-class configuration; these exercise the rest of its constructor)
    "deep_gap": dict(n_layers=4, depth=2, channels=[8, 12, 18, 27], kernels=[5, 3, 7, 3], n_classes=3, classifier="gap"),
    "k5_fc": dict(n_layers=5, depth=1, channels=[10, 15, 22, 33, 50], kernels=[3, 5, 3, 5, 3], n_classes=2, classifier="gap_fc"),
    "wide_d3": dict(n_layers=3, depth=3, channels=[16, 40, 96], kernels=[3, 3, 3], n_classes=4, classifier="gap_fc"),
}


def generic_cnn_state_dict(cfg, seed, as_torch=True):
    """Seeded state-dict with the key names and shapes ConvNet(cfg) has (riser/nets/cnn.py:13-41,52-65): He-normal
    conv weights, small biases, a head scaled so that the class probabilities spread instead of saturating."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    sd = {}
    cin = 1
    for i in range(cfg["n_layers"]):
        cout, k = cfg["channels"][i], cfg["kernels"][i]
        for d in range(cfg["depth"]):
            ci = cin if d == 0 else cout
            sd[f"layers.{i}.{2 * d}.weight"] = rng.normal(0.0, np.sqrt(2.0 / (ci * k)), size=(cout, ci, k)).astype(np.float32)
            sd[f"layers.{i}.{2 * d}.bias"] = rng.normal(0.0, 0.05, size=(cout,)).astype(np.float32)
        cin = cout
    nc = cfg["n_classes"]
    w = rng.normal(0.0, 2.0 / np.sqrt(cin), size=(nc, cin)).astype(np.float32)
    b = rng.normal(0.0, 0.3, size=(nc,)).astype(np.float32)
    if cfg["classifier"] == "gap":
        sd["classifier.0.weight"], sd["classifier.0.bias"] = w[:, :, None].copy(), b
    else:
        sd["classifier.2.weight"], sd["classifier.2.bias"] = w, b
    if as_torch:
        import torch
        sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    return sd


# ---------------------------------------------------------------- ResNet variant (riser/nets/resnet.py)
RESNET_CONFIGS = {
    # the reference ships no resnet config; these two exercise both block types.
    # decoder_scale / decoder_centre: calibrated once with the reference (seed 0, 48 ragged
    # reads) so that the logit difference is centred with a spread of ~2 -> probabilities
    # span (0, 1) instead of saturating.
    "basic": dict(channels=[20, 30, 45, 67], kernel=19, stride=3, padding=5, block="basic", n_layers=4,
                  blocks=[2, 2, 2, 2], n_classes=2, decoder_scale=0.5, decoder_centre=102.62 / 3.0),
    "bottleneck": dict(channels=[32, 64, 128, 256], kernel=19, stride=3, padding=5, block="bottleneck",
                       n_layers=4, blocks=[1, 2, 2, 1], n_classes=2, decoder_scale=3.0,
                       decoder_centre=-14.91 * 2.0),
}


def resnet_state_dict(cfg, seed, as_torch=True):
    """Seeded state-dict with the reference ResNet's key names and shapes (resnet.py:7-131):
    kaiming-like conv weights, non-trivial BatchNorm affine parameters and running statistics
    (so that BN folding is really tested), unused shortcut parameters included exactly as the
    reference constructs them (resnet.py:21-24)."""
    rng = np.random.Generator(np.random.PCG64(5000 + seed))
    sd = {}

    def conv(name, cout, cin, k, bias=False):
        sd[name + ".weight"] = rng.normal(0.0, np.sqrt(2.0 / (cin * k)), size=(cout, cin, k)).astype(np.float32)
        if bias:
            sd[name + ".bias"] = rng.normal(0.0, 0.05, size=(cout,)).astype(np.float32)

    def bn(name, c):
        sd[name + ".weight"] = rng.uniform(0.6, 1.4, size=(c,)).astype(np.float32)
        sd[name + ".bias"] = rng.normal(0.0, 0.1, size=(c,)).astype(np.float32)
        sd[name + ".running_mean"] = rng.normal(0.0, 0.2, size=(c,)).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
        sd[name + ".num_batches_tracked"] = np.array(100, dtype=np.int64)

    ch = cfg["channels"]
    conv("conv_block.0", ch[0], 1, cfg["kernel"], bias=True)
    bn("conv_block.1", ch[0])
    cin = ch[0]
    for i in range(cfg["n_layers"]):
        cout = ch[i]
        for j in range(cfg["blocks"][i]):
            p = f"layers.{i}.{j}"
            if cfg["block"] == "bottleneck":
                mid = cout // 4
                conv(p + ".blocks.0.0", mid, cin, 1); bn(p + ".blocks.0.1", mid)
                conv(p + ".blocks.1.0", mid, mid, 3); bn(p + ".blocks.1.1", mid)
                conv(p + ".blocks.2.0", cout, mid, 1); bn(p + ".blocks.2.1", cout)
            else:
                conv(p + ".blocks.0.0", cout, cin, 3); bn(p + ".blocks.0.1", cout)
                conv(p + ".blocks.1.0", cout, cout, 3); bn(p + ".blocks.1.1", cout)
            conv(p + ".shortcut.0", cout, cin, 1); bn(p + ".shortcut.1", cout)
            cin = cout
    # zero-sum rows and a small scale keep the logits O(1) (post-ReLU features have a large common
    # mean), so that probabilities are not saturated and parity tests mean something
    wdec = rng.normal(0.0, 1.0, size=(cfg["n_classes"], ch[-1]))
    wdec -= wdec.mean(axis=1, keepdims=True)
    sd["decoder.2.weight"] = (wdec * cfg.get("decoder_scale", 0.3) / np.sqrt(ch[-1])).astype(np.float32)
    bias = rng.normal(0.0, 0.1, size=(cfg["n_classes"],))
    bias[1] -= cfg.get("decoder_centre", 0.0)
    sd["decoder.2.bias"] = bias.astype(np.float32)
    if as_torch:
        import torch
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    return sd
