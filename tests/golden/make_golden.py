#!/usr/bin/env python
"""Generate the committed golden fixtures by running the REAL reference
(comprna/riser, imported read-only from /root/reference via oracle/refshim.py).

Run once in the build container:   python tests/golden/make_golden.py
Writes tests/golden/*.npz and riser_b200/data/synth_head_seed{0,1,2}.npz.
Nothing here runs on the GPU box; the fixtures travel instead.

What the reference functions are fed is always regenerated from seeds by
riser_b200/synth.py, so the fixtures only store seeds, shapes, small inputs and
the reference's outputs (or their SHA-256 where the output is large).
"""
import hashlib
import io
import logging
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim                      # noqa: E402
from riser_b200 import synth, sim               # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "riser_b200", "data")
ref = refshim.load()
LOG = logging.getLogger("golden")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def proc_for(kit):
    return ref.preprocess.SignalProcessor(ref.preprocess.Kit.create_from_version(kit))


# ---------------------------------------------------------------- A. edge cases
def edge_cases():
    rng = np.random.Generator(np.random.PCG64(5))
    proc = proc_for("RNA002")
    out = {}
    cases = {
        "odd": rng.integers(300, 700, size=4097).astype(np.int16),
        "even": rng.integers(300, 700, size=4096).astype(np.int16),
        "tiny3": np.array([500, 520, 480], dtype=np.int16),
        "tiny2": np.array([500, 520], dtype=np.int16),
        "one": np.array([512], dtype=np.int16),
        "constant": np.full(5000, 431, dtype=np.int16),
        "mad0_with_spike": np.concatenate([np.full(4999, 431), [900]]).astype(np.int16),
        "negatives": rng.integers(-600, 600, size=6001).astype(np.int16),
        "full_range": np.concatenate([rng.integers(-32768, 32768, size=8190), [-32768, 32767]]).astype(np.int16),
        "two_level": np.where(rng.random(5001) < 0.5, 100, 101).astype(np.int16),
    }
    # outliers at both ends and in runs
    base = rng.normal(500, 30, size=4200)
    base[0] += 900; base[1] += 900; base[-1] -= 800; base[-2] -= 800
    base[100:105] += 700; base[200] -= 650; base[201] += 650; base[300:302] += 500
    cases["end_outliers_runs"] = np.rint(base).astype(np.int16)
    base = rng.normal(500, 30, size=4099)
    base[0] -= 900; base[-1] += 900
    cases["single_end_outliers"] = np.rint(base).astype(np.int16)
    base = rng.normal(500, 30, size=5000)
    base[1000:1400] += 600          # one very long run
    cases["long_run"] = np.rint(base).astype(np.int16)
    names = sorted(cases)
    out["norm_names"] = np.array(names)
    for n in names:
        x = cases[n]
        y = proc.mad_normalise(x)
        out[f"norm_in_{n}"] = x
        out[f"norm_out_{n}"] = np.asarray(y)            # float64, or int64 when mad == 0
    # _smooth_outliers called directly
    for k, v in {"a": [9., 9., 0.], "b": [0., 10., 10., 10., 1., -9., 0.5, 8.],
                 "c": [-4., 0., 4.], "d": [3.5, -3.5, 3.6, 0.]}.items():
        out[f"smooth_in_{k}"] = np.array(v)
        out[f"smooth_out_{k}"] = proc._smooth_outliers(np.array(v))
    # empty -> ValueError
    try:
        proc.mad_normalise(np.array([], dtype=np.int16))
        out["empty_raises"] = np.array(0)
    except ValueError:
        out["empty_raises"] = np.array(1)
    # kit constants / length gates (preprocess.py:20-40,81-85)
    for kit in ("RNA002", "RNA004"):
        p = proc_for(kit)
        out[f"kit_{kit}"] = np.array([p.kit.sampling_hz, p.kit.transloc_rate, p.get_min_length(),
                                      p.get_max_length(), p.get_fixed_trim_length()])
    np.savez_compressed(os.path.join(GOLD, "preprocess_edge.npz"), **out)
    print("edge cases:", len(names))


# ---------------------------------------------------------------- B. polyA + realistic reads
from tests.golden.make_golden_params import (POLYA_SEED, POLYA_READS, POLYA_PREFIXES, NORM_SEED,   # noqa: E402
                                             NORM_READS, NORM_EXTRA_LENGTHS, SCEN, polya_reads,
                                             norm_inputs, scenario_reads)


def polya_cases():
    proc = proc_for("RNA002")
    reads = polya_reads()
    ends = np.full((POLYA_READS, len(POLYA_PREFIXES)), -1, dtype=np.int64)
    for r, (_, sig) in enumerate(reads):
        for j, n in enumerate(POLYA_PREFIXES):
            e = proc.get_polyA_end(sig[:n])
            ends[r, j] = -1 if e is None else e
    # hand-made traces for the truthiness corner: a quiet raised window at index 0
    rng = np.random.Generator(np.random.PCG64(6))
    hand = {}
    t = np.concatenate([rng.normal(620, 5, 1500), rng.normal(500, 60, 3000)])
    hand["quiet_from_zero"] = np.rint(t).astype(np.int16)
    t = np.concatenate([rng.normal(400, 40, 2000), rng.normal(620, 5, 2000), rng.normal(500, 60, 3000)])
    hand["clean_step"] = np.rint(t).astype(np.int16)
    t = np.concatenate([rng.normal(400, 40, 1000), rng.normal(620, 5, 700), rng.normal(500, 60, 3000)])
    hand["early_step_no_history"] = np.rint(t).astype(np.int16)
    t = np.concatenate([rng.normal(-400, 40, 2000), rng.normal(-300, 5, 2000), rng.normal(-500, 60, 3000)])
    hand["negative_levels"] = np.rint(t).astype(np.int16)
    out = {"seed": np.array(POLYA_SEED), "n_reads": np.array(POLYA_READS),
           "prefixes": np.array(POLYA_PREFIXES), "ends": ends, "hand_names": np.array(sorted(hand))}
    for k in sorted(hand):
        e = proc.get_polyA_end(hand[k])
        out[f"hand_in_{k}"] = hand[k]
        out[f"hand_end_{k}"] = np.array(-1 if e is None else e)
    # trim_polyA cache semantics (preprocess.py:87-102): found -> cached; not found -> not cached;
    # hit short-circuits even if the signal changed
    cache = {}
    rid, sig = reads[0]
    log = []
    for n in (3000, 9000, 12000, 600):
        s, trimmed = proc.trim_polyA(sig[:n], rid, cache)
        log.append([n, len(s), int(trimmed), cache.get(rid, -1)])
    out["cache_log"] = np.array(log)
    np.savez_compressed(os.path.join(GOLD, "polya.npz"), **out)
    print("polyA found in", int((ends >= 0).any(axis=1).sum()), "of", POLYA_READS, "reads;",
          "hand:", {k: int(out[f'hand_end_{k}']) for k in sorted(hand)})


def realistic_norm():
    proc = proc_for("RNA002")
    bodies = norm_inputs()
    digests64, digests32, sums = [], [], []
    normed = []
    for x in bodies:
        y = proc.mad_normalise(x)
        normed.append(y)
        digests64.append(sha(y.astype(np.float64)))
        digests32.append(sha(y.astype(np.float32)))
        sums.append(float(np.sum(y)))
    np.savez_compressed(os.path.join(GOLD, "normalise_reads.npz"),
                        seed=np.array(NORM_SEED), lengths=np.array([len(x) for x in bodies]),
                        sha_f64=np.stack(digests64), sha_f32=np.stack(digests32), sums=np.array(sums),
                        first8=np.stack([y[:8] for y in normed]), last8=np.stack([y[-8:] for y in normed]))
    print("normalise reads:", len(bodies))
    return bodies, normed


# ---------------------------------------------------------------- C. fitted heads
def ref_net(state):
    m = ref.ConvNet(refshim.cnn_config().cnn).eval()
    m.load_state_dict(state)
    return m


def fit_heads():
    os.makedirs(DATA, exist_ok=True)
    proc = proc_for("RNA002")
    bodies = synth.ragged_bodies(77, 192, 4096, 12048)
    xs = [torch.tensor(np.asarray(proc.mad_normalise(x)), dtype=torch.float) for x in bodies]
    y = torch.tensor([r & 1 for r in range(len(xs))], dtype=torch.float)
    for target, seed in synth.TARGET_SEEDS.items():
        sd = {k: torch.from_numpy(v) for k, v in synth.conv_weights(seed).items()}
        sd["classifier.2.weight"] = torch.zeros(2, 1702)
        sd["classifier.2.bias"] = torch.zeros(2)
        m = ref_net(sd)
        with torch.no_grad():
            feats = []
            for x in xs:
                h = x.view(1, 1, -1)
                for layer in m.layers:
                    h = layer(h)
                feats.append(h.mean(dim=2)[0])
            f = torch.stack(feats)
        mu = f.mean(0)
        torch.manual_seed(seed)
        w = torch.zeros(1702, requires_grad=True)
        b = torch.zeros(1, requires_grad=True)
        opt = torch.optim.Adam([w, b], lr=0.01)
        for _ in range(400):
            opt.zero_grad()
            z = (f - mu) @ w + b
            loss = F.binary_cross_entropy_with_logits(z, y) + 0.1 * (w * w).sum()
            loss.backward()
            opt.step()
        w, b = w.detach(), b.detach()
        W = torch.stack([-w / 2, w / 2]).numpy().astype(np.float32)
        c = float(mu @ w - b[0])
        Bv = np.array([c / 2, -c / 2], dtype=np.float32)
        np.savez(os.path.join(DATA, f"synth_head_seed{seed}.npz"), weight=W, bias=Bv)
        p = F.softmax(f @ torch.from_numpy(W).T + torch.from_numpy(Bv), dim=1)[:, 1].numpy()
        print(f"head {target}: p_on quantiles", np.quantile(p, [.05, .25, .5, .75, .95]).round(3),
              "frac>0.9 %.2f frac<0.1 %.2f" % ((p > .9).mean(), (p < .1).mean()))


# ---------------------------------------------------------------- D. ConvNet / Model goldens
def model_goldens(bodies, normed):
    tmp = tempfile.mkdtemp()
    cfg = refshim.cnn_config()
    out = {"seed": np.array(NORM_SEED), "lengths": np.array([len(x) for x in bodies])}
    names = []
    for target, seed in synth.TARGET_SEEDS.items():
        path = synth.save_state_dict(seed, os.path.join(tmp, f"{target}.pth"))
        mdl = ref.model.Model(path, cfg, LOG, target)          # the reference's own load path
        assert str(mdl.device) == "cpu"
        probs = np.stack([mdl.classify(y).numpy() for y in normed])
        out[f"probs_{target}"] = probs.astype(np.float32)
        names.append(target)
        print(f"model {target}: p_on range {probs[:,1].min():.3f}..{probs[:,1].max():.3f}")
    out["targets"] = np.array(names)
    # BASELINE config 1: already-trimmed 12,048-sample bodies, test.py ladder 4096/7108/10120
    # (riser/test.py:202-224) and the full 12,048 window the live path shows (control.py:46)
    ladder = [4096, 7108, 10120, 12048]
    X = synth.body_batch(1234, 48, 12048)
    proc = proc_for("RNA002")
    mdl = ref.model.Model(os.path.join(tmp, "mRNA.pth"), cfg, LOG, "mRNA")
    lad = np.zeros((X.shape[0], len(ladder), 2), dtype=np.float32)
    for r in range(X.shape[0]):
        for j, n in enumerate(ladder):
            lad[r, j] = mdl.classify(proc.mad_normalise(X[r, :n])).numpy()
    out["cfg1_seed"] = np.array(1234)
    out["cfg1_ladder"] = np.array(ladder)
    out["cfg1_probs"] = lad
    np.savez_compressed(os.path.join(GOLD, "convnet_probs.npz"), **out)


# ---------------------------------------------------------------- E. control loop scenario
def control_scenario():
    tmp = tempfile.mkdtemp()
    cfg = refshim.cnn_config()
    models = []
    for t in SCEN["targets"]:
        path = synth.save_state_dict(synth.TARGET_SEEDS[t], os.path.join(tmp, f"{t}.pth"))
        models.append(ref.model.Model(path, cfg, LOG, t))
    reads = scenario_reads()
    out = {k: np.array(v) for k, v in SCEN.items()}
    for mode in ("deplete", "enrich"):
        client = sim.SimClient(reads, SCEN["chunk"], SCEN["n_polls"], first_len=SCEN["first_len"])
        base = os.path.join(tmp, f"run_{mode}")
        ctl = ref.control.SequencerControl(client, models, proc_for(SCEN["kit"]), LOG, base)
        ctl.start()
        ctl.target(mode, 1, SCEN["threshold"])
        ctl.finish()
        with open(base + ".csv") as f:
            rows = [ln.rstrip("\n") for ln in f]
        header, rows = rows[0], [ln.split(",", 1)[1] for ln in rows[1:]]   # drop batch_start
        out[f"header"] = np.array(header)
        out[f"rows_{mode}"] = np.array(rows)
        out[f"unblocked_{mode}"] = np.array(client.unblocked, dtype=np.int64).reshape(-1, 2)
        out[f"finished_{mode}"] = np.array(client.finished, dtype=np.int64).reshape(-1, 2)
        dec = [r.rsplit(",", 1)[1] for r in rows]
        print(f"control {mode}: {len(rows)} assessed rows;", {d: dec.count(d) for d in sorted(set(dec))})
    np.savez_compressed(os.path.join(GOLD, "control_scenario.npz"), **out)


# ---------------------------------------------------------------- F. ResNet variant (nets/resnet.py)
def resnet_goldens():
    """The reference's own ResNet module (eval mode) on seeded weights with non-trivial BatchNorm
    statistics; the reference ships no config for it, so both block types are exercised with
    the configurations in riser_b200.synth.RESNET_CONFIGS."""
    proc = proc_for("RNA002")
    bodies = synth.ragged_bodies(321, 24, 4096, 12048)
    normed = [np.asarray(proc.mad_normalise(b)) for b in bodies]
    out = {"seed": np.array(321), "lengths": np.array([len(b) for b in bodies])}
    for name, cfg in synth.RESNET_CONFIGS.items():
        m = ref.ResNet(refshim.AttrDict(cfg)).eval()
        sd = synth.resnet_state_dict(cfg, 0)
        assert set(sd) == set(m.state_dict())
        m.load_state_dict(sd)
        with torch.no_grad():
            probs = np.stack([F.softmax(m(torch.from_numpy(x).float()[None]), dim=1)[0].numpy() for x in normed])
        out[f"probs_{name}"] = probs.astype(np.float32)
        print(f"resnet {name}: p_on range {probs[:, 1].min():.3f}..{probs[:, 1].max():.3f}")
    np.savez_compressed(os.path.join(GOLD, "resnet_probs.npz"), **out)


# ---------------------------------------------------------------- G. retrain/preprocess.py (float32 path)
def retrain_goldens():
    from oracle import retrain_oracle as rt
    rp = refshim.load_retrain_preprocess()
    rng = np.random.Generator(np.random.PCG64(77))
    digests, lengths = [], []
    edge_in, edge_out = {}, {}
    bodies = synth.ragged_bodies(555, 16, 4096, 12048)
    for k, raw in enumerate(bodies):
        pa = rt.pa_signal(raw, scale=0.1 + 0.01 * k, offset=3.0 * k - 10)
        digests.append(sha(rp.mad_normalise(pa.copy(), 3.5)))
        lengths.append(len(pa))
    cases = {"even_small": rng.normal(80, 9, size=4096), "odd_small": rng.normal(80, 9, size=4097),
             "ties": np.round(rng.normal(80, 2, size=5000)), "constant": np.full(4500, 71.25),
             "negative": rng.normal(-30, 4, size=6000), "tiny": np.array([1.5, 2.5, 9.0, 2.0, 1.0])}
    spiky = rng.normal(80, 5, size=5000)
    spiky[0] += 90; spiky[-1] -= 90; spiky[100:104] += 70; spiky[200] -= 60; spiky[201] += 60
    cases["end_outliers_runs"] = spiky
    for name, v in cases.items():
        x = np.asarray(v, dtype=np.float32)
        with np.errstate(all="ignore"):
            y = rp.mad_normalise(x.copy(), 3.5)
        edge_in[name], edge_out[name] = x, np.asarray(y)
    np.savez_compressed(os.path.join(GOLD, "retrain_norm.npz"), seed=np.array(555), lengths=np.array(lengths),
                        sha=np.stack(digests), names=np.array(sorted(cases)),
                        **{f"in_{k}": v for k, v in edge_in.items()}, **{f"out_{k}": v for k, v in edge_out.items()})
    print("retrain goldens:", len(digests), "reads +", len(cases), "edge cases")


if __name__ == "__main__":
    logging.basicConfig(level=logging.WARNING)
    torch.set_num_threads(8)
    edge_cases()
    polya_cases()
    bodies, normed = realistic_norm()
    if "--skip-heads" not in sys.argv:
        fit_heads()
    model_goldens(bodies, normed)
    control_scenario()
    resnet_goldens()
    retrain_goldens()
    print("done")
