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
        self._acts = {}
        # (kernel, stride, padding) of every op of the main chain, kernel < 0 = the stem's max-pool (riser_len_chain)
        chain = [(self.stem.k, self.stem.stride, self.stem.pad), (-1, 2, 1)]
        for convs, _ in self.blocks:
            chain += [(cv.k, cv.stride, cv.pad) for cv in convs]
        self.n_chain = len(chain)
        self.chain = torch.tensor(chain, dtype=torch.int32).contiguous().to(dev)

    # ------------------------------------------------------------------ launches
    def _buf(self, key, *shape):
        # rows at or beyond a read's length are never read by any kernel (each takes the per-read lengths), so the
        # activation buffers need no clearing and are reused from call to call
        t = self._acts.get((key,) + shape)
        if t is None:
            t = self._acts[(key,) + shape] = torch.empty(*shape, dtype=torch.float32, device=self.device)
        return t

    def _conv(self, cv, x, n_in, L_in, n_out, residual=None, relu=True):
        B = x.shape[0]
        L_out = max(1, (L_in + 2 * cv.pad - cv.k) // cv.stride + 1)
        out = self._buf(id(cv), B, L_out, cv.cout)
        _lib.check(_lib.lib().riser_conv1d_cl(_lib.ptr(x), _lib.ptr(n_in), _lib.ptr(cv.w), _lib.ptr(cv.b),
                                              _lib.ptr(residual), _lib.ptr(out), _lib.ptr(n_out), B, L_in, L_out,
                                              cv.cin, cv.cout, cv.k, cv.stride, cv.pad, 1 if relu else 0,
                                              _lib.stream_ptr()), "riser_conv1d_cl")
        return out, L_out

    def classify_batch(self, x, lens, max_len=None, probs=None, **_unused):
        """x: fp32 [B, ld] normalised signals on the device, lens int32 [B].  -> probs [B, n_classes]."""
        B = x.shape[0]
        L0 = int(max_len if max_len is not None else x.shape[1])
        xin = x[:, :L0].contiguous().view(B, L0, 1)
        lens = lens.to(torch.int32)
        # valid lengths after every op of the main chain, one launch (stem conv, stem pool, every block conv)
        n_all = self._acts.get(("len", B))
        if n_all is None:
            n_all = self._acts[("len", B)] = torch.empty(self.n_chain, B, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().riser_len_chain(_lib.ptr(lens), B, _lib.ptr(self.chain), self.n_chain, _lib.ptr(n_all),
                                              _lib.stream_ptr()), "riser_len_chain")
        h, L = self._conv(self.stem, xin, lens, L0, n_all[0])
        L_p = L // 2 + 1                                                   # MaxPool1d(2, 2, padding=1)
        pooled = self._buf("pool", B, L_p, self.stem.cout)
        _lib.check(_lib.lib().riser_maxpool1d_cl(_lib.ptr(h), _lib.ptr(n_all[0]), _lib.ptr(pooled), _lib.ptr(n_all[1]),
                                                 B, L, L_p, self.stem.cout, _lib.stream_ptr()), "riser_maxpool1d_cl")
        h, n, L, j = pooled, n_all[1], L_p, 2
        for convs, shortcut in self.blocks:
            n_block = n_all[j + len(convs) - 1]             # the shortcut's output is as long as the block's
            res = h
            if shortcut is not None:
                res, _ = self._conv(shortcut, h, n, L, n_block, relu=False)
            y, ny, Ly = h, n, L
            for k, cv in enumerate(convs):
                last = k == len(convs) - 1
                y, Ly = self._conv(cv, y, ny, Ly, n_all[j], residual=res if last else None, relu=True)
                ny = n_all[j]
                j += 1
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
