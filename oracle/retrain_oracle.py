"""numpy float32 restatement of riser/retrain/preprocess.py:8-44 (test infrastructure, see
oracle/__init__).  Input is the pA-scaled float32 signal, so numpy keeps every step in float32;
there is no MAD == 0 guard.  Pinned against the reference's own functions in
tests/golden/retrain_norm.npz (bit-for-bit)."""
import numpy as np


def mad_normalise(signal, outlier_lim=3.5):
    x = np.asarray(signal, dtype=np.float32)
    if x.shape[0] == 0:
        raise ValueError("Signal must not be empty")
    med = np.median(x)                                    # float32 (mean of the middle pair in float32)
    mad = np.median(np.abs(x - med))                      # :36-39
    with np.errstate(divide="ignore", invalid="ignore"):
        arr = (x - med) / (np.float32(1.4826) * mad)      # :42-44, float32 throughout
    lim = np.float32(outlier_lim)
    n = len(arr)
    for i in np.flatnonzero(np.abs(arr) > lim):           # :18-33
        if i == 0:
            arr[0] = arr[1]
        elif i == n - 1:
            arr[i] = arr[i - 1]
        else:
            v = (arr[i - 1] + arr[i + 1]) / np.float32(2)
            arr[i] = lim if v > lim else (-lim if v < -lim else v)
    return arr


def pa_signal(raw, scale=0.1456, offset=12.0):
    """What ont_fast5_api's get_raw_data(scale=True) returns: float32 scale * (raw + offset)."""
    return np.array(np.float32(scale) * (raw + np.float32(offset)), dtype=np.float32)
