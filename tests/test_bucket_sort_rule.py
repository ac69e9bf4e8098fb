"""The rule of the sweep's bucket sort (swiftest_b200/csrc/encounter_kernels.cu: bucket_map / bucket_sort_kernel), restated
in numpy and held to the stable sort the reference's semantics fix (ties in position order, tests/test_oracle.py): the map
key -> bucket must be monotone -- it is a chain of individually rounded IEEE operations -- so that buckets sorted on
(key, id) and concatenated ARE the stably sorted sequence."""
import numpy as np
import pytest

BUCKET_TARGET, BUCKET_MAXNB, BUCKET_CAP = 128, 1 << 16, 2048


def device_bucket_sort(keys):
    """-> (order, overflow): the permutation the device produces, and whether a bucket exceeds the CTA's capacity."""
    n = len(keys)
    nb = 1
    while nb < BUCKET_MAXNB and nb * BUCKET_TARGET < n:
        nb <<= 1
    kmin, kmax = keys.min(), keys.max()
    scale = np.float64(nb) / (kmax - kmin) if kmax > kmin else np.float64(0.0)
    with np.errstate(invalid="ignore"):
        b = ((keys - kmin) * scale).astype(np.int64)     # cvt.rzi: truncation
    b = np.clip(b, 0, nb - 1)
    assert np.all(np.diff(b[np.argsort(keys, kind="stable")]) >= 0), "bucket map is not monotone in the key"
    order = np.lexsort((np.arange(n), keys, b))            # by bucket, then (key, id) inside the bucket
    return order, np.bincount(b, minlength=nb).max() > BUCKET_CAP


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (216, 2), (20000, 3), (200000, 4)])
def test_bucket_sort_is_the_stable_sort_on_disk_extents(n, seed):
    from swiftest_b200 import workloads as W
    d = W.disk(max(n // 2, 1), seed=seed)
    rmag = np.sqrt((d["rh"] * d["rh"]).sum(1))
    w = 1.1 * d["rhill"] * 6.5
    keys = np.concatenate([rmag - w, rmag + w])
    order, overflow = device_bucket_sort(keys)
    assert not overflow
    assert np.array_equal(order, np.argsort(keys, kind="stable"))


def test_bucket_sort_with_ties_negative_extents_and_extreme_ranges():
    rng = np.random.default_rng(7)
    base = rng.uniform(0.3, 40.0, 5000)
    keys = np.concatenate([base, base[:700], -base[:50], [0.0] * 30, [1e-300, np.nextafter(40.0, 0.0)], base[:700]])
    order, overflow = device_bucket_sort(keys)
    assert not overflow
    assert np.array_equal(order, np.argsort(keys, kind="stable"))
    # one far outlier squeezes everything else into the first bucket: the device must notice (and take the radix sort)
    keys = np.concatenate([rng.uniform(1.0, 1.001, 5000), [1e12]])
    order, overflow = device_bucket_sort(keys)
    assert overflow
    assert np.array_equal(order, np.argsort(keys, kind="stable"))   # the rule itself still sorts
    # all keys equal: scale = 0, one bucket
    keys = np.full(1500, 2.5)
    order, overflow = device_bucket_sort(keys)
    assert not overflow and np.array_equal(order, np.arange(1500))
