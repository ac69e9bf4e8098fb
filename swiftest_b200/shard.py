"""Host-side sharding logic for the multi-GPU path (one process per GPU).

* pl-pl gravity: balanced contiguous i-slices (counts differ by at most one, larger slices first), identical
  to swcu_partition in the C ABI.
* test particles: block partition ceil(ntot/nimages) per image exactly like
  swiftest_coarray_distribute_system (reference swiftest/swiftest_coarray.f90:705-711).
"""


def partition(n, nranks, rank):
    """Balanced contiguous slice [i0, i1) of n units for `rank` of `nranks` (mirrors swcu_partition)."""
    if nranks <= 0 or not (0 <= rank < nranks) or n < 0:
        raise ValueError("bad partition arguments")
    q, r = divmod(n, nranks)
    i0 = rank * q + min(rank, r)
    return i0, i0 + q + (1 if rank < r else 0)


def tp_block_partition(ntot, nimages, image):
    """Coarray-style block partition (0-based image): ntp_per_image = ceil(ntot/nimages); image k gets
    [k*per, min((k+1)*per, ntot))  (swiftest_coarray.f90:705-711)."""
    if nimages <= 0 or not (0 <= image < nimages) or ntot < 0:
        raise ValueError("bad partition arguments")
    per = -(-ntot // nimages)
    i0 = min(image * per, ntot)
    return i0, min(i0 + per, ntot)
