"""numpy restatement of riser/preprocess.py (test infrastructure, see oracle/__init__).

All arithmetic is float64 exactly as numpy does it for the reference, so the
outputs are bit-identical to the reference's (checked against
tests/golden/preprocess_*.npz, which were produced by the reference itself).
"""
import numpy as np

# riser/preprocess.py:6-12
OUTLIER_LIMIT = 3.5
SCALING_FACTOR = 1.4826
MIN_INPUT_SIGNALS = 4096
MAX_INPUT_NT = 280
TRIM_RESOLUTION = 500
TRIM_MAD_THRESHOLD = 20
TRIM_FIXED_LENGTH_NT = 150.6

# riser/preprocess.py:20-27
KITS = {"RNA002": (3012, 70), "RNA004": (4000, 130)}


def kit_constants(version):
    """(sampling_hz, transloc_rate) -- riser/preprocess.py:20-27 (raises on unknown kit)."""
    if version not in KITS:
        raise Exception(f"Invalid kit version {version}")
    return KITS[version]


def max_length(version):
    """riser/preprocess.py:36-37."""
    hz, rate = kit_constants(version)
    return int(MAX_INPUT_NT / rate * hz)


def fixed_trim_length(version):
    """riser/preprocess.py:81-82."""
    hz, rate = kit_constants(version)
    return int(TRIM_FIXED_LENGTH_NT / rate * hz)


def window_mad(window, median):
    """riser/preprocess.py:117-120."""
    return np.median(np.abs(window - median))


def polya_end(signal):
    """riser/preprocess.py:42-79.  Returns the window-start index of the first
    high-MAD window after a low-MAD/raised-mean window, or None.

    Python truthiness is part of the algorithm: a start (or end) found at index 0
    counts as "not found" (riser/preprocess.py:62,66)."""
    res = TRIM_RESOLUTION
    start = None
    end = None
    n = len(signal)
    for i in range(0, n - res + 1, res):
        w = signal[i:i + res]
        med = np.median(w)
        mad = window_mad(w, med)
        mean = np.mean(w)
        rolling = np.mean(signal[i - 2 * res:i]) if i > 2 * res else mean
        change = (mean - rolling) / rolling * 100
        if not start and change > 20 and mad <= TRIM_MAD_THRESHOLD:
            start = i
        if start and not end and mad > 20:
            end = i
    return end


def trim_polya(signal, read_id, cache):
    """riser/preprocess.py:87-102: cache hit short-circuits detection; only found
    ends are cached; the cut is at end+1."""
    if read_id in cache:
        end = cache[read_id]
    else:
        end = polya_end(signal)
        if end:
            cache[read_id] = end
    if end:
        return signal[end + 1:], True
    return signal, False


def smooth_outliers(arr):
    """riser/preprocess.py:127-147.  Outlier set is fixed up front (strict > 3.5);
    ascending, in place; interior points are averaged then clipped, end points
    copy their single neighbour unclipped."""
    n = len(arr)
    for i in np.flatnonzero(np.abs(arr) > OUTLIER_LIMIT):
        if i == 0:
            arr[0] = arr[1]
        elif i == n - 1:
            arr[i] = arr[i - 1]
        else:
            v = (arr[i - 1] + arr[i + 1]) / 2
            arr[i] = min(max(v, -OUTLIER_LIMIT), OUTLIER_LIMIT)
    return arr


def mad_normalise(signal):
    """riser/preprocess.py:108-125.  float64 out; when MAD == 0 the reference's
    np.vectorize returns integer zeros (dtype int64) -- reproduced here."""
    signal = np.asarray(signal)
    if signal.shape[0] == 0:
        raise ValueError("Signal must not be empty")
    med = np.median(signal)
    mad = window_mad(signal, med)
    if mad == 0:
        return smooth_outliers(np.zeros(signal.shape[0], dtype=np.int64))
    out = (signal - med) / (SCALING_FACTOR * mad)
    return smooth_outliers(np.asarray(out, dtype=np.float64))


def select_window(signal, read_id, cache, version):
    """Length gating of riser/control.py:36-60.  Returns (window or None, trimmed)."""
    mx = max_length(version)
    sig, trimmed = trim_polya(signal, read_id, cache)
    if not trimmed:
        fixed = fixed_trim_length(version)
        if len(sig) > fixed + mx:                 # preprocess.py:84-85 (strict)
            return sig[fixed:][:mx], False        # control.py:43-46
        return None, False                        # control.py:49-50
    if len(sig) < MIN_INPUT_SIGNALS:              # control.py:55-56
        return None, True
    return sig[:mx], True                         # control.py:59-60
