"""torch-CPU fp32 restatement of riser/nets/cnn.py (ConvNet, 'gap_fc' head) and
riser/model.py:22-28 (classify).  Test infrastructure, see oracle/__init__.

Works directly on a state-dict with the reference's key names
(``layers.{i}.0.weight|bias``, ``classifier.2.weight|bias``; ConvNet with
depth 1 -- the only shape the shipped configs use, riser/model/*.yaml:6-12).
"""
import numpy as np
import torch
import torch.nn.functional as F

# riser/model/*_config_*.yaml:6-12 (identical in all six shipped configs)
CHANNELS = [20, 30, 45, 67, 100, 150, 225, 337, 505, 757, 1135, 1702]
KERNELS = [3] * 12
N_CLASSES = 2


def n_conv_layers(state):
    n = 0
    while f"layers.{n}.0.weight" in state:
        n += 1
    return n


def features(state, x):
    """Conv trunk + global average pool: x [B, L] fp32 -> [B, C_last].
    riser/nets/cnn.py:43-47 with _make_layer (cnn.py:52-65): Conv1d(k, stride 1,
    padding='same', bias) -> ReLU -> MaxPool1d(2, 2); then AdaptiveAvgPool1d(1)
    + Flatten (cnn.py:29-31)."""
    h = x.unsqueeze(1)
    for i in range(n_conv_layers(state)):
        w = state[f"layers.{i}.0.weight"]
        b = state[f"layers.{i}.0.bias"]
        h = F.max_pool1d(F.relu(F.conv1d(h, w, b, stride=1, padding="same")), 2, 2)
    return h.mean(dim=2)


def logits(state, x):
    """riser/nets/cnn.py:43-50, 'gap_fc' classifier (cnn.py:28-33)."""
    return F.linear(features(state, x), state["classifier.2.weight"], state["classifier.2.bias"])


def classify(state, signal):
    """riser/model.py:22-28: numpy (float64 / int64) -> fp32 -> net -> softmax -> row 0."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(signal)).unsqueeze(0).to(dtype=torch.float)
        return F.softmax(logits(state, x), dim=1)[0]


def classify_ragged(state, signals):
    """Per-read classify over a ragged list (the reference has no batched ragged
    path; riser/control.py:68-71 calls classify once per read).  -> [B, 2] fp32."""
    return torch.stack([classify(state, s) for s in signals]).numpy()


def flops_per_read(length, channels=CHANNELS):
    """Algorithmic FLOPs, SURVEY.md 8(d): sum 2*3*Cin*Cout*L_i + 2*C_last*2."""
    total, cin, l = 0, 1, int(length)
    for c in channels:
        total += 2 * 3 * cin * c * l
        cin, l = c, l // 2
    return total + 2 * channels[-1] * N_CLASSES
