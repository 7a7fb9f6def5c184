"""Host logic of the live path (CPU): the slice of a read that classify_batch uploads must contain exactly the
window the reference's gating (control.py:36-60, via the oracle) selects for a read with a cached poly(A) end;
batch sizes round up to a small set of buckets."""
import numpy as np

from oracle import preprocess_oracle as pp
from riser_b200.pipeline import upload_ranges, bucket_size


def test_upload_ranges_cover_the_selected_window():
    rng = np.random.default_rng(4)
    mx, mn = pp.max_length("RNA002"), pp.MIN_INPUT_SIGNALS
    for _ in range(400):
        n = int(rng.integers(0, 30000))
        end = int(rng.integers(1, 9000))
        sig = rng.integers(0, 1000, size=n).astype(np.int16)
        skip, take = upload_ranges([n], np.array([end], dtype=np.int32), mn, mx)
        window, trimmed = pp.select_window(sig, "r", {"r": end}, "RNA002")
        assert trimmed
        if window is None:
            assert take[0] == 0
        else:
            assert skip[0] == end + 1 and take[0] == len(window)
            assert np.array_equal(sig[skip[0]:skip[0] + take[0]], window)
    # boundary: exactly min_len samples after the end is assessed, one fewer is not
    for extra, want in ((mn, mn), (mn - 1, 0), (mx + 5, mx)):
        _, take = upload_ranges([5001 + extra], np.array([5000], dtype=np.int32), mn, mx)
        assert take[0] == want
    # no cached end: the whole prefix goes up
    skip, take = upload_ranges([12345, 0], np.array([-1, -1], dtype=np.int32), mn, mx)
    assert skip.tolist() == [0, 0] and take.tolist() == [12345, 0]


def test_bucket_sizes():
    got = [bucket_size(n) for n in (1, 64, 65, 96, 97, 128, 129, 300, 512, 513, 3000)]
    assert got == [64, 64, 96, 96, 128, 128, 192, 384, 512, 768, 3072]
