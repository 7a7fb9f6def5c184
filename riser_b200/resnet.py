"""ResNet variant of the classifier (riser/nets/resnet.py) behind the same Model call surface.

The reference can only select it in train.py:177-178 (``config.model == 'resnet'``,
``ResNet(config.resnet)``) and ships no config or weights for it, so the hyper-parameters are
whatever the caller's config says: ``channels, kernel, padding, stride, block, n_layers,
blocks, n_classes`` (resnet.py:73-102).  BatchNorm (eval mode) is folded into the convolution
weights here on the host; the convolutions, the stem max-pool and the head run in the fp32
CUDA-core kernels of csrc/resnet.cu (see the note there about tensor cores).
"""
import numpy as np
import torch

from . import _lib

EPS = 1e-5


def _fold(w, b, bn):
    """Conv weight [Cout, Cin, K] (+ optional bias) followed by eval-mode BatchNorm1d -> one
    conv with w' = w * g / sqrt(var + eps), b' = beta + (b - mean) * g / sqrt(var + eps)."""
    g, beta, mean, var = bn
    scale = g / torch.sqrt(var + EPS)
    wf = w * scale[:, None, None]
    bf = beta + ((b if b is not None else 0.0) - mean) * scale
    return wf, bf


class _Conv:
    def __init__(self, w, b, stride, pad, device):
        self.cout, self.cin, self.k = w.shape
        self.stride, self.pad = stride, pad
        self.w = w.permute(2, 1, 0).contiguous().to(device)        # [K][Cin][Cout]
        self.b = b.contiguous().to(device)

    def out_len(self, n):
        return torch.clamp((n + 2 * self.pad - self.k) // self.stride + 1, min=0).to(torch.int32)


class ResNetModel():
    def __init__(self, state, config, logger, target):
        self.target = target
        self.logger = logger
        self.device = _lib.require_device()
        self.logger.info('Using %s device', self.device)
        c = config.resnet if hasattr(config, "resnet") else config
        sd = state if isinstance(state, dict) else torch.load(state, map_location=torch.device('cpu'))
        sd = {k: torch.as_tensor(v).detach().to("cpu", torch.float32) for k, v in sd.items()
              if not k.endswith("num_batches_tracked")}
        self.c = c
        self.n_classes = int(c.n_classes)
        dev = self.device

        def bn(prefix):
            return tuple(sd[prefix + s] for s in (".weight", ".bias", ".running_mean", ".running_var"))

        def conv_bn(prefix, stride, pad, bias=None):
            w, b = _fold(sd[prefix + ".0.weight"], bias, bn(prefix + ".1"))
            return _Conv(w, b, stride, pad, dev)

        self.stem = conv_bn("conv_block", int(c.stride), int(c.padding), bias=sd["conv_block.0.bias"])
        self.blocks = []
        cin = int(c.channels[0])
        for i in range(int(c.n_layers)):
            cout = int(c.channels[i])
            for j in range(int(c.blocks[i])):
                stride = 2 if (i > 0 and j == 0) else 1
                p = f"layers.{i}.{j}"
                shortcut = None
                if cin != cout or stride != 1:                              # resnet.py:45-47
                    shortcut = conv_bn(p + ".shortcut", stride, 0)
                if c.block == "bottleneck":
                    convs = [conv_bn(p + ".blocks.0", 1, 0), conv_bn(p + ".blocks.1", stride, 1),
                             conv_bn(p + ".blocks.2", 1, 0)]
                else:
                    convs = [conv_bn(p + ".blocks.0", stride, 1), conv_bn(p + ".blocks.1", 1, 1)]
                self.blocks.append((convs, shortcut))
                cin = cout
        self.fc_w = sd["decoder.2.weight"].contiguous().to(dev)
        self.fc_b = sd["decoder.2.bias"].contiguous().to(dev)
        self.c_last = cin

    # ------------------------------------------------------------------ launches
    def _conv(self, cv, x, n_in, L_in, residual=None, relu=True):
        B = x.shape[0]
        n_out = cv.out_len(n_in)
        L_out = max(1, (L_in + 2 * cv.pad - cv.k) // cv.stride + 1)
        out = torch.zeros(B, L_out, cv.cout, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().riser_conv1d_cl(_lib.ptr(x), _lib.ptr(n_in), _lib.ptr(cv.w), _lib.ptr(cv.b),
                                              _lib.ptr(residual), _lib.ptr(out), _lib.ptr(n_out), B, L_in, L_out,
                                              cv.cin, cv.cout, cv.k, cv.stride, cv.pad, 1 if relu else 0,
                                              _lib.stream_ptr()), "riser_conv1d_cl")
        return out, n_out, L_out

    def classify_batch(self, x, lens, max_len=None, probs=None, **_unused):
        """x: fp32 [B, ld] normalised signals on the device, lens int32 [B].  -> probs [B, n_classes]."""
        B = x.shape[0]
        L0 = int(max_len if max_len is not None else x.shape[1])
        xin = x[:, :L0].contiguous().view(B, L0, 1)
        h, n, L = self._conv(self.stem, xin, lens.to(torch.int32), L0)
        n_p = (n // 2 + 1).to(torch.int32)                                 # MaxPool1d(2, 2, padding=1)
        n_p = torch.where(n > 0, n_p, torch.zeros_like(n_p))
        L_p = L // 2 + 1
        pooled = torch.zeros(B, L_p, self.stem.cout, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().riser_maxpool1d_cl(_lib.ptr(h), _lib.ptr(n), _lib.ptr(pooled), _lib.ptr(n_p), B, L,
                                                 L_p, self.stem.cout, _lib.stream_ptr()), "riser_maxpool1d_cl")
        h, n, L = pooled, n_p, L_p
        for convs, shortcut in self.blocks:
            res = h
            if shortcut is not None:
                res, _, _ = self._conv(shortcut, h, n, L, relu=False)
            y, ny, Ly = h, n, L
            for k, cv in enumerate(convs):
                last = k == len(convs) - 1
                y, ny, Ly = self._conv(cv, y, ny, Ly, residual=res if last else None, relu=True)
            h, n, L = y, ny, Ly
        if probs is None:
            probs = torch.empty(B, self.n_classes, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().riser_gap_linear_softmax(_lib.ptr(h), _lib.ptr(n), _lib.ptr(self.fc_w),
                                                       _lib.ptr(self.fc_b), _lib.ptr(probs), B, L, self.c_last,
                                                       self.n_classes, _lib.stream_ptr()),
                   "riser_gap_linear_softmax")
        return probs

    def classify(self, signal):
        """riser/model.py:22-28 for the ResNet: 1-D numpy -> Tensor[n_classes] on the device."""
        signal = np.asarray(signal)
        n = signal.shape[0]
        x = torch.from_numpy(signal).to(self.device, dtype=torch.float).view(1, n)
        lens = torch.tensor([n], dtype=torch.int32, device=self.device)
        return self.classify_batch(x, lens, max_len=n)[0]
