"""torch-CPU fp32 restatement of riser/nets/resnet.py (ResNet, BasicBlock, BottleneckBlock)
in eval mode, working directly on a state-dict with the reference's key names.  Test
infrastructure, see oracle/__init__.  Pinned against the reference's own ResNet module in
tests/golden/make_golden.py (fixture tests/golden/resnet_probs.npz)."""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5   # nn.BatchNorm1d default (resnet.py:23,29,80)


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=EPS)


def _conv_block(x, sd, prefix, relu, stride=1, padding=0):
    """resnet.py:26-37: Conv1d (no bias) -> BatchNorm1d [-> ReLU]."""
    x = F.conv1d(x, sd[prefix + ".0.weight"], None, stride=stride, padding=padding)
    x = _bn(x, sd, prefix + ".1")
    return F.relu(x) if relu else x


def _residual_block(x, sd, prefix, cin, cout, stride, kind):
    """resnet.py:39-47 forward, with BasicBlock (:50-57) or BottleneckBlock (:60-70)."""
    if cin != cout or stride != 1:                         # should_apply_shortcut, :45-47
        res = _bn(F.conv1d(x, sd[prefix + ".shortcut.0.weight"], None, stride=stride), sd, prefix + ".shortcut.1")
    else:
        res = x
    if kind == "bottleneck":
        h = _conv_block(x, sd, prefix + ".blocks.0", True)
        h = _conv_block(h, sd, prefix + ".blocks.1", True, stride=stride, padding=1)
        h = _conv_block(h, sd, prefix + ".blocks.2", False)
    else:
        h = _conv_block(x, sd, prefix + ".blocks.0", True, stride=stride, padding=1)
        h = _conv_block(h, sd, prefix + ".blocks.1", False, padding=1)
    return F.relu(h + res)


def logits(sd, c, x):
    """resnet.py:104-110.  c: config with channels, kernel, padding, stride, block, n_layers, blocks."""
    h = x.unsqueeze(1)
    h = F.conv1d(h, sd["conv_block.0.weight"], sd["conv_block.0.bias"], stride=c["stride"], padding=c["padding"])
    h = F.relu(_bn(h, sd, "conv_block.1"))
    h = F.max_pool1d(h, 2, stride=2, padding=1)
    cin = c["channels"][0]
    for i in range(c["n_layers"]):
        cout = c["channels"][i]
        for j in range(c["blocks"][i]):
            stride = 2 if (i > 0 and j == 0) else 1        # resnet.py:90,114-121
            h = _residual_block(h, sd, f"layers.{i}.{j}", cin, cout, stride, c["block"])
            cin = cout
    return F.linear(h.mean(dim=2), sd["decoder.2.weight"], sd["decoder.2.bias"])


def classify(sd, c, signal):
    """model.py:22-28 applied to the ResNet."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(signal)).unsqueeze(0).to(dtype=torch.float)
        return F.softmax(logits(sd, c, x), dim=1)[0]
