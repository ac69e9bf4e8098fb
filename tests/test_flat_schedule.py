"""CPU restatement of the integer bookkeeping of the third-law gravity kernel (swiftest_b200/csrc/kick_flat_kernels.cu):
the block-pair enumeration (items_of / item_decode) and the graded claim schedule (split_quanta / range).  A mistake in
either double-counts or drops pair interactions, so both are checked exhaustively over many shapes here; the kernel
itself is compared with the oracle on the GPU (tests/test_gpu_parity.py)."""
import itertools

FIB = 4  # 32-column chunks per block pair


def geometry(nb, nbm):
    Km = (nbm - 1) // 2
    evenm = 1 if nbm % 2 == 0 else 0
    return Km, evenm


def items_of(I, nb, nbm, Km, evenm):
    return 1 + Km + (1 if (evenm and I < nbm // 2) else 0) + (nb - nbm)


def item_decode(t, nb, nbm, Km, evenm):
    base = 1 + Km + (nb - nbm)
    if evenm:
        big = (base + 1) * (nbm // 2)
        if t < big:
            I, k = divmod(t, base + 1)
        else:
            I, k = divmod(t - big, base)
            I += nbm // 2
    else:
        I, k = divmod(t, base)
    kmI = Km + (1 if (evenm and I < nbm // 2) else 0)
    if k == 0:
        return I, I, True
    if k <= kmI:
        return I, (I + k) % nbm, False
    return I, nbm + (k - 1 - kmI), False


def test_block_pairs_are_enumerated_exactly_once():
    for nb in range(1, 34):
        for nbm in range(1, nb + 1):
            Km, evenm = geometry(nb, nbm)
            total = sum(items_of(I, nb, nbm, Km, evenm) for I in range(nbm))
            seen = {}
            for t in range(total):
                I, J, diag = item_decode(t, nb, nbm, Km, evenm)
                assert 0 <= I < nbm and 0 <= J < nb and diag == (I == J)
                key = (min(I, J), max(I, J))
                assert key not in seen, (nb, nbm, key)
                seen[key] = t
            want = {(a, b) for a in range(nb) for b in range(a, nb) if a < nbm}   # at least one owner block
            assert set(seen) == want, (nb, nbm)
            # equal work per owner block up to one item
            per = [items_of(I, nb, nbm, Km, evenm) for I in range(nbm)]
            assert max(per) - min(per) <= 1


def split_quanta(items, warps_all, quantum, fine=(2, 4, 8, 16)):
    U, G = items * FIB, quantum * FIB
    left = U
    sz = [G, 8, 4, 2, 1]
    nq = [0] * 5
    for k in (4, 3, 2, 1):
        if sz[k] >= G:
            continue
        want = min(left, warps_all * fine[4 - k])
        want -= want % sz[k]
        nq[k] = want // sz[k]
        left -= want
    nq[0] = (left + G - 1) // G
    ph_q, ph_u, qacc, uacc = [], [], 0, 0
    for k in range(5):
        ph_q.append(qacc)
        ph_u.append(uacc)
        qacc += nq[k]
        uacc += left if k == 0 else nq[k] * sz[k]
    ph_q.append(qacc)
    ph_u.append(uacc)
    return ph_q, ph_u, sz, qacc


def claim_range(q, ph_q, ph_u, sz, unit0, unit1):
    if q >= ph_q[5]:
        return None
    k = sum(1 for j in range(1, 5) if q >= ph_q[j])
    u0 = unit0 + ph_u[k] + (q - ph_q[k]) * sz[k]
    u1 = min(u0 + sz[k], unit0 + ph_u[k + 1])
    return (u0, u1) if u0 < unit1 else None


def test_graded_schedule_tiles_the_run_exactly():
    shapes = itertools.product((1, 2, 3, 7, 50, 391, 5000, 38269), (1, 4, 12, 1776, 14208), (1, 2, 4, 6),
                               ((2, 4, 8, 16), (0, 4, 8, 16), (0, 0, 0, 0), (4, 4, 8, 0), (1, 1, 1, 1)))
    for items, warps, quantum, fine in shapes:
        for unit0 in (0, 4 * 123):
            ph_q, ph_u, sz, nquanta = split_quanta(items, warps, quantum, fine)
            unit1 = unit0 + items * FIB
            assert ph_u[5] == items * FIB
            cursor = unit0
            sizes = []
            for q in range(nquanta):
                r = claim_range(q, ph_q, ph_u, sz, unit0, unit1)
                assert r is not None and r[0] == cursor and r[0] < r[1] <= unit1, (items, warps, quantum, fine, q, r)
                cursor = r[1]
                sizes.append(r[1] - r[0])
            assert cursor == unit1
            for q in (nquanta, nquanta + 1, nquanta + 10 ** 6):
                assert claim_range(q, ph_q, ph_u, sz, unit0, unit1) is None
            # claims never grow along the run, and whole-item claims start on item boundaries
            G = quantum * FIB
            assert all(a >= b or a < G for a, b in zip(sizes, sizes[1:]))
            n_coarse = ph_q[1]
            assert all((unit0 + i * G) % FIB == 0 for i in range(n_coarse))


def rank_runs(total, nranks):
    """kick_pl_flat: balanced consecutive runs of items per rank (the shape of swcu_partition)."""
    q, r = divmod(total, nranks)
    out = []
    for rk in range(nranks):
        i0 = rk * q + min(rk, r)
        out.append((i0, i0 + q + (1 if rk < r else 0)))
    return out


def test_rank_runs_tile_the_items_and_every_rank_schedules_its_own_run():
    """Several GPUs: every rank owns a consecutive run of items and hands it out with its OWN counter (no shared
    counter over NVLink any more); the runs tile [0, total) with sizes within one item, and the graded schedule of every
    run covers exactly that run."""
    for total, nranks in ((306153, 8), (306153, 4), (306153, 2), (7, 8), (8, 8), (100, 3), (1, 2)):
        runs = rank_runs(total, nranks)
        assert runs[0][0] == 0 and runs[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
        sizes = [b - a for a, b in runs]
        assert max(sizes) - min(sizes) <= 1
        for i0, i1 in runs:
            items = i1 - i0
            ph_q, ph_u, sz, nquanta = split_quanta(items, 1776, 4 if items >= 32 * 1776 else (2 if items >= 16 * 1776 else 1))
            cursor = i0 * FIB
            for q in range(nquanta):
                r = claim_range(q, ph_q, ph_u, sz, i0 * FIB, i1 * FIB)
                assert r is not None and r[0] == cursor
                cursor = r[1]
            assert cursor == i1 * FIB
