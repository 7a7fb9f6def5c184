"""The oracle (oracle/) against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as pp
from oracle import convnet_oracle as net
from oracle import control_oracle as ctl
from riser_b200 import synth, sim

from tests.golden import make_golden_params as P


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def edge(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess_edge.npz"))


def test_kit_constants(edge):
    for kit in ("RNA002", "RNA004"):
        hz, rate = pp.kit_constants(kit)
        got = [hz, rate, pp.MIN_INPUT_SIGNALS, pp.max_length(kit), pp.fixed_trim_length(kit)]
        assert got == list(edge[f"kit_{kit}"])
    assert pp.max_length("RNA002") == 12048 and pp.max_length("RNA004") == 8615
    with pytest.raises(Exception):
        pp.kit_constants("RNA003")


def test_normalise_edge_cases_bit_exact(edge):
    for name in edge["norm_names"]:
        want = edge[f"norm_out_{name}"]
        got = pp.mad_normalise(edge[f"norm_in_{name}"])
        assert got.dtype == want.dtype, name          # int64 zeros when MAD == 0
        assert np.array_equal(got, want), name


def test_smooth_outliers_direct(edge):
    for k in "abcd":
        got = pp.smooth_outliers(edge[f"smooth_in_{k}"].copy())
        assert np.array_equal(got, edge[f"smooth_out_{k}"]), k
    assert list(pp.smooth_outliers(np.array([9., 9., 0.]))) == [9., 3.5, 0.]


def test_empty_raises(edge):
    assert int(edge["empty_raises"]) == 1
    with pytest.raises(ValueError):
        pp.mad_normalise(np.array([], dtype=np.int16))


def test_normalise_realistic_reads_sha(golden_dir):
    g = np.load(os.path.join(golden_dir, "normalise_reads.npz"))
    bodies = P.norm_inputs()
    assert [len(b) for b in bodies] == list(g["lengths"])
    for i, x in enumerate(bodies):
        y = pp.mad_normalise(x)
        assert np.array_equal(sha(y.astype(np.float64)), g["sha_f64"][i]), i
        assert np.array_equal(sha(y.astype(np.float32)), g["sha_f32"][i]), i


def test_polya_end(golden_dir):
    g = np.load(os.path.join(golden_dir, "polya.npz"))
    reads = P.polya_reads()
    for r, (_, sig) in enumerate(reads):
        for j, n in enumerate(g["prefixes"]):
            e = pp.polya_end(sig[:n])
            assert (-1 if e is None else e) == g["ends"][r, j], (r, n)
    for k in g["hand_names"]:
        e = pp.polya_end(g[f"hand_in_{k}"])
        assert (-1 if e is None else e) == int(g[f"hand_end_{k}"]), k
    # cache semantics
    cache, log = {}, []
    rid, sig = reads[0]
    for n in (3000, 9000, 12000, 600):
        s, trimmed = pp.trim_polya(sig[:n], rid, cache)
        log.append([n, len(s), int(trimmed), cache.get(rid, -1)])
    assert np.array_equal(np.array(log), g["cache_log"])


def test_convnet_probs(golden_dir):
    g = np.load(os.path.join(golden_dir, "convnet_probs.npz"))
    bodies = P.norm_inputs()
    normed = [pp.mad_normalise(x) for x in bodies]
    for target in g["targets"]:
        state = synth.state_dict(synth.TARGET_SEEDS[str(target)])
        got = net.classify_ragged(state, normed[:10] + normed[-8:])
        want = np.concatenate([g[f"probs_{target}"][:10], g[f"probs_{target}"][-8:]])
        assert np.abs(got - want).max() < 1e-6, target


def test_config1_ladder(golden_dir):
    g = np.load(os.path.join(golden_dir, "convnet_probs.npz"))
    X = synth.body_batch(int(g["cfg1_seed"]), 48, 12048)
    state = synth.state_dict(0)
    for r in (0, 1, 17):
        for j, n in enumerate(g["cfg1_ladder"]):
            p = net.classify(state, pp.mad_normalise(X[r, :n])).numpy()
            assert np.abs(p - g["cfg1_probs"][r, j]).max() < 1e-6


def test_flops_table():
    # SURVEY.md 8(d) / appendix A.2
    assert round(net.flops_per_read(12048) / 1e6, 1) == 435.1
    assert round(net.flops_per_read(8615) / 1e6, 1) == 314.9
    assert round(net.flops_per_read(16000) / 1e6, 1) == 584.9
    assert round(net.flops_per_read(4096) / 1e6, 1) == 153.2


@pytest.mark.parametrize("mode", ["deplete", "enrich"])
def test_control_scenario(golden_dir, mode):
    """oracle/control_oracle.run_batch driven by the simulated client reproduces
    the rows the reference's SequencerControl.target wrote (control.py:31-106)."""
    g = np.load(os.path.join(golden_dir, "control_scenario.npz"))
    reads = P.scenario_reads()
    states = [synth.state_dict(synth.TARGET_SEEDS[str(t)]) for t in g["targets"]]
    client = sim.SimClient(reads, int(g["chunk"]), int(g["n_polls"]), first_len=int(g["first_len"]))
    client.start_streaming_reads()
    thr = float(g["threshold"])
    cache, rows = {}, []
    while client.is_running():
        batch = client.get_read_batch()
        items = [(rd.id, client.get_raw_signal(rd)) for _, rd in batch]
        dec, p_on, _, sig_len, cache = ctl.run_batch(items, states, str(g["kit"]), cache, thr, mode)
        rej, fin = [], []
        for (ch, rd), d, po, n in zip(batch, dec, p_on, sig_len):
            if d == ctl.SKIPPED:
                continue
            rows.append((rd.id, ch, int(n), [float(x) for x in po], ctl.NAMES[int(d)]))
            if d == ctl.REJECT:
                rej.append((ch, rd.number))
            if d in (ctl.REJECT, ctl.ACCEPT, ctl.NO_DECISION):
                fin.append((ch, rd.number))
        client.reject_reads(rej, 0.1)
        # control.py:104-106 orders done = reject + accept + unclassified; compare as sets per batch
        client.finish_processing_reads(fin)
    want = [r.split(",") for r in g[f"rows_{mode}"]]
    assert len(rows) == len(want)
    for got, w in zip(rows, want):
        rid, ch, n, po, d = got
        assert [rid, str(ch), str(n)] == w[:3]
        assert d == w[-1]
        wp = [float(x) for x in w[4].split(";")]
        assert np.abs(np.array(po) - np.array(wp)).max() < 1e-6
    assert sorted(map(tuple, g[f"unblocked_{mode}"])) == sorted(client.unblocked)
    assert sorted(map(tuple, g[f"finished_{mode}"])) == sorted(client.finished)


def test_resnet_oracle_against_reference_golden(golden_dir):
    from oracle import resnet_oracle as ro
    g = np.load(os.path.join(golden_dir, "resnet_probs.npz"))
    bodies = synth.ragged_bodies(int(g["seed"]), 24, 4096, 12048)
    assert [len(b) for b in bodies] == list(g["lengths"])
    for name, cfg in synth.RESNET_CONFIGS.items():
        sd = synth.resnet_state_dict(cfg, 0)
        for i in (0, 5, 11, 23):
            p = ro.classify(sd, cfg, pp.mad_normalise(bodies[i])).numpy()
            assert np.abs(p - g[f"probs_{name}"][i]).max() < 1e-6, (name, i)


def test_retrain_oracle_bit_exact_against_reference(golden_dir):
    from oracle import retrain_oracle as rt
    g = np.load(os.path.join(golden_dir, "retrain_norm.npz"))
    bodies = synth.ragged_bodies(int(g["seed"]), 16, 4096, 12048)
    for k, raw in enumerate(bodies):
        pa = rt.pa_signal(raw, scale=0.1 + 0.01 * k, offset=3.0 * k - 10)
        assert np.array_equal(sha(rt.mad_normalise(pa)), g["sha"][k]), k
    for name in g["names"]:
        got, want = rt.mad_normalise(g[f"in_{name}"]), g[f"out_{name}"]
        assert got.dtype == want.dtype == np.float32
        assert np.array_equal(got, want, equal_nan=True), name
