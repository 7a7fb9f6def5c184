"""riser_b200/_hostpack (csrc/hostpack.c): the native gather of read prefixes into the staging buffer.
CPU only -- compared with the plain numpy slice assignments it replaces."""
import numpy as np
import pytest

from riser_b200 import build
from riser_b200.preprocess import _hostpack


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build_hostpack()


def _reads(rng, B):
    sigs = [rng.integers(-2000, 2000, size=int(n)).astype(np.int16) for n in rng.integers(0, 9000, size=B)]
    sigs[3] = np.zeros(0, dtype=np.int16)
    sigs[5] = sigs[5].tobytes()                    # the client hands over bytes (read.raw_data)
    return sigs


def _as16(s):
    return np.frombuffer(s, np.int16) if isinstance(s, bytes) else s


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_pack_matches_numpy(threads):
    hp = _hostpack()
    rng = np.random.default_rng(1)
    sigs = _reads(rng, 700)
    nbytes = np.zeros(len(sigs), dtype=np.int64)
    hp.lengths(sigs, nbytes)
    assert np.array_equal(nbytes, [2 * len(_as16(s)) for s in sigs])
    n = nbytes >> 1
    pos = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum((n + 7) & ~7, out=pos[1:])
    dst = np.full(int(pos[-1]) + 8, -1, dtype=np.int16)
    want = dst.copy()
    for s, o, k in zip(sigs, pos[:-1], n):
        want[o:o + k] = _as16(s)
    hp.pack(sigs, dst, pos[:-1] << 1, np.zeros_like(n), nbytes, threads)
    assert np.array_equal(dst, want)
    # slices: signal[skip : skip + take]
    skip, take = n // 3, n // 2
    pos = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum((take + 7) & ~7, out=pos[1:])
    dst = np.full(int(pos[-1]) + 8, -1, dtype=np.int16)
    want = dst.copy()
    for s, o, a, k in zip(sigs, pos[:-1], skip, take):
        want[o:o + k] = _as16(s)[a:a + k]
    hp.pack(sigs, dst, pos[:-1] << 1, skip << 1, take << 1, threads)
    assert np.array_equal(dst, want)


def test_pack_rejects_bad_ranges():
    hp = _hostpack()
    sigs = [np.arange(10, dtype=np.int16), np.arange(6, dtype=np.int16)]
    dst = np.zeros(32, dtype=np.int16)
    z = np.zeros(2, dtype=np.int64)
    with pytest.raises(ValueError):
        hp.pack(sigs, dst, z, z, np.array([22, 12], dtype=np.int64), 1)        # beyond the first read
    with pytest.raises(ValueError):
        hp.pack(sigs, dst, np.array([60, 0], dtype=np.int64), z, np.array([20, 12], dtype=np.int64), 1)   # beyond dst
    with pytest.raises(ValueError):
        hp.pack(sigs, dst, z[:1], z, z, 1)                                      # one int64 per item
    with pytest.raises((TypeError, BufferError)):
        hp.pack([object()], dst, z[:1], z[:1], np.array([2], dtype=np.int64), 1)
    hp.pack([object()], dst, z[:1], z[:1], z[:1], 1)                            # take == 0: the item is not touched
