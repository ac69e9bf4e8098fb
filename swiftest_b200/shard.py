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


def needs_rebalance(counts):
    """swiftest_coarray_balance_system (swiftest_coarray.f90:14-60): test particles are collected and redistributed
    when the largest and the smallest image differ by at least the number of images."""
    counts = list(counts)
    if not counts:
        return False
    return max(counts) - min(counts) >= len(counts)


def rebalance_plan(counts):
    """What coarray_collect + coarray_distribute do to the particle blocks, as a transfer plan.

    `counts[k]` = active particles image k holds now (its block stays contiguous and ordered: collect concatenates
    the images in order, distribute cuts the concatenation into ceil(ntot/nimages) blocks).  Returns a list of
    (src_image, src_lo, src_hi, dst_image, dst_lo) tuples: particles [src_lo, src_hi) of src_image's local array go
    to dst_image's new local array starting at dst_lo.  Transfers with src_image == dst_image stay on the device."""
    counts = list(counts)
    nimg = len(counts)
    ntot = sum(counts)
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    plan = []
    for dst in range(nimg):
        d0, d1 = tp_block_partition(ntot, nimg, dst)
        for src in range(nimg):
            lo, hi = max(d0, starts[src]), min(d1, starts[src + 1])
            if lo < hi:
                plan.append((src, lo - starts[src], hi - starts[src], dst, lo - d0))
    return plan
