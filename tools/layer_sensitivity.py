"""Per-layer precision sensitivity of the ConvNet's softmax output (CPU emulation, no GPU needed).

Which conv layers could run as ONE fp16 tensor pass (operands rounded to fp16, no hi/lo correction terms) inside the
north star's |dp| <= 1e-3 bar?  The emulation rounds the operands of selected layers exactly as the CUDA path's
single-pass mode does (activations and weights to fp16, round-to-nearest-even) and accumulates in fp32 on the CPU
(torch conv1d); every other layer is exact fp32.  What it leaves out is the tensor cores' truncating fp32
accumulation, which is the 1.5e-4 floor all modes share on the GPU, so the numbers below are
lower bounds on what the GPU would show.

For every layer i = 1..11 (layer 0 runs in fp32-equivalent arithmetic on every path):
  only_i      layer i alone single-pass, the rest exact            -> that layer's own contribution
  w_only_i    only the WEIGHTS of layer i rounded (activations exact): the "drop the W_lo term" half
  a_only_i    only the ACTIVATIONS of layer i rounded:               the "drop the a_lo term" half
  from_i      layers i..11 single-pass, 1..i-1 exact                 -> cumulative, what a mode switch at i costs
  upto_i      layers 1..i single-pass, i+1..11 exact

usage: python tools/layer_sensitivity.py [reads_per_model=48] [L=12048] > profiles/<round>_layer_sensitivity.json
The oracle is used as the checker here (tools/ is measurement infrastructure, not the product path).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from riser_b200 import synth                      # noqa: E402
from oracle import preprocess_oracle as pre       # noqa: E402


def r16(t):
    return t.to(torch.float16).to(torch.float32)


def layer(state, i, h, round_w=False, round_a=False):
    w, b = state[f"layers.{i}.0.weight"], state[f"layers.{i}.0.bias"]
    if round_w:
        w = r16(w)
    if round_a:
        h = r16(h)
    return F.max_pool1d(F.relu(F.conv1d(h, w, b, stride=1, padding="same")), 2, 2)


def head(state, h):
    return F.softmax(F.linear(h.mean(dim=2), state["classifier.2.weight"], state["classifier.2.bias"]), dim=1)


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 12048
    torch.set_grad_enabled(False)
    sig = synth.body_batch(7, n_reads, L)
    x = torch.from_numpy(np.stack([np.asarray(pre.mad_normalise(s), dtype=np.float32) for s in sig]))
    n_layers = len(synth.CHANNELS)
    out = {"what": "max / mean |dp| of the softmax output against the exact fp32 forward when selected layers take "
                   "fp16-rounded operands (CPU emulation, fp32 accumulate; tools/layer_sensitivity.py)",
           "reads_per_model": n_reads, "samples_per_read": L, "models": list(synth.TARGET_SEEDS),
           "bar": 1e-3, "rows": []}
    acc = {}

    def note(key, p, base):
        d = (p - base).abs()
        m = acc.setdefault(key, [0.0, 0.0, 0])
        m[0] = max(m[0], float(d.max()))
        m[1] += float(d.sum())
        m[2] += d.numel()

    for name, seed in synth.TARGET_SEEDS.items():
        state = synth.state_dict(seed)
        acts = [x.unsqueeze(1)]                   # acts[i] = exact input of layer i
        for i in range(n_layers):
            acts.append(layer(state, i, acts[i]))
        base = head(state, acts[n_layers])

        def run(first, rounded, round_w=True, round_a=True):
            h = acts[first]
            for i in range(first, n_layers):
                on = i in rounded
                h = layer(state, i, h, on and round_w, on and round_a)
            return head(state, h)

        for i in range(1, n_layers):
            note(("only", i), run(i, {i}), base)
            note(("w_only", i), run(i, {i}, True, False), base)
            note(("a_only", i), run(i, {i}, False, True), base)
            note(("from", i), run(i, set(range(i, n_layers))), base)
            note(("upto", i), run(1, set(range(1, i + 1))), base)
        print(f"{name} done", file=sys.stderr)
    for i in range(1, n_layers):
        row = {"layer": i, "cin": synth.CHANNELS[i - 1], "cout": synth.CHANNELS[i]}
        for kind in ("only", "w_only", "a_only", "from", "upto"):
            m = acc[(kind, i)]
            row[kind + "_max"] = m[0]
            row[kind + "_mean"] = m[1] / m[2]
        out["rows"].append(row)
    out["all_single_pass_max"] = acc[("from", 1)][0]
    out["layers_single_pass_alone_within_bar"] = [r["layer"] for r in out["rows"] if r["only_max"] <= 1e-3]
    out["layers_single_pass_alone_within_half_bar"] = [r["layer"] for r in out["rows"] if r["only_max"] <= 5e-4]
    print(json.dumps(out, indent=1))
    print("layer  only_max  w_only   a_only   from_max  upto_max", file=sys.stderr)
    for r in out["rows"]:
        print(f"{r['layer']:5d}  {r['only_max']:.2e} {r['w_only_max']:.2e} {r['a_only_max']:.2e} "
              f"{r['from_max']:.2e}  {r['upto_max']:.2e}", file=sys.stderr)


if __name__ == "__main__":
    main()
