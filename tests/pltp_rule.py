"""numpy restatement of the rule the sort-free pl-tp sweep applies on the device
(swiftest_b200/csrc/encounter_kernels.cu::pltp_direct_kernel), used by the CPU suite to hold the RULE to the oracle's
sort-and-sweep (encounter_check.f90:261-326, :795-902) without a GPU, and by the GPU suite as a second opinion."""
import numpy as np


def check_one(xr, yr, zr, vxr, vyr, vzr, renc, dt):
    """encounter_check_one (encounter_check.f90:591-618), element-wise, one IEEE operation per Fortran operation."""
    vsmall = np.sqrt(np.finfo(np.float64).tiny)  # globals_module.f90:135
    r2 = xr * xr + yr * yr + zr * zr
    r2crit = renc * renc
    vdotr = vxr * xr + vyr * yr + vzr * zr
    v2 = vxr * vxr + vyr * vyr + vzr * vzr
    with np.errstate(all="ignore"):
        tmin = -vdotr / v2
        r2min = np.where(tmin < dt, r2 - vdotr * vdotr / v2, r2 + 2 * vdotr * dt + v2 * (dt * dt))
    r2min = np.where((vdotr > 0.0) | (v2 <= vsmall), r2, r2min)
    inside = ~(r2 > r2crit)
    vdotr = np.where(inside, -1.0, vdotr)
    r2min = np.where(inside, r2, r2min)
    return (vdotr < 0.0) & (r2min <= r2crit)


def direct_pltp(rpl, vpl, rtp, vtp, renc, dt):
    """-> (index1, index2, nbox_total, needs_sort_path).  index1/index2 1-based, canonical order."""
    n1 = len(renc)
    rmag = np.sqrt(rpl[:, 0] * rpl[:, 0] + rpl[:, 1] * rpl[:, 1] + rpl[:, 2] * rpl[:, 2])
    w = 1.1 * renc
    rmin, rmax = rmag - w, rmag + w
    K = np.sqrt(rtp[:, 0] * rtp[:, 0] + rtp[:, 1] * rtp[:, 1] + rtp[:, 2] * rtp[:, 2])
    i1, i2, nbox, flag = [], [], 0, bool(np.isnan(rmin).any() or np.isnan(rmax).any())
    for i in range(n1):
        in_b = (K >= rmin[i]) & (K <= rmax[i])
        in_e = in_b & (K < rmax[i])
        flag |= bool((in_b != in_e).any())
        c = int(in_b.sum()) + int(in_e.sum())
        for j in range(n1):
            if j == i:
                continue
            if (rmin[j] > rmin[i] or (rmin[j] == rmin[i] and j > i)) and rmin[j] <= rmax[i]:
                c += 1
            if rmax[j] >= rmin[i] and (rmax[j] < rmax[i] or (rmax[j] == rmax[i] and j < i)):
                c += 1
        if c >= 2:
            nbox += c
        q = np.nonzero(in_b)[0]
        d = rtp[q] - rpl[i]
        dv = vtp[q] - vpl[i]
        hit = check_one(d[:, 0], d[:, 1], d[:, 2], dv[:, 0], dv[:, 1], dv[:, 2], renc[i] + 0.0, dt)
        i1.append(np.full(int(hit.sum()), i + 1, np.int32))
        i2.append((q[hit] + 1).astype(np.int32))
    return np.concatenate(i1) if i1 else np.zeros(0, np.int32), np.concatenate(i2) if i2 else np.zeros(0, np.int32), nbox, flag


def pairless_particle_boxes(rtp):
    """What the reference's nbox sum holds beyond the planets' boxes: each of g >= 3 particles with bit-identical |r| finds
    one endpoint of each of the g - 1 others inside its own degenerate interval (no planet among them, so no pair
    comes of it); the device's statistic leaves these out (swcu_encounter_stats is a diagnostic, not a reference output)."""
    K = np.sqrt(rtp[:, 0] * rtp[:, 0] + rtp[:, 1] * rtp[:, 1] + rtp[:, 2] * rtp[:, 2])
    _, g = np.unique(K, return_counts=True)
    g = g[g >= 3]
    return int((g * (g - 1)).sum())


def tie_case(kind, ntp=4000, seed=5):
    """8 planets + a cloud in which particles sit EXACTLY on a planet's inner extent ('rmin'), outer extent ('rmax'),
    or share |r| with other particles ('dup')."""
    from swiftest_b200 import workloads as W
    p = W.planets8_year_units()
    tp = W.tp_cloud(ntp, seed=seed)
    rtp, vtp = tp["rh"].copy(), tp["vh"].copy()
    renc = p["rhill"] * 6.5
    rpl = p["rh"]
    rmag = np.sqrt(rpl[:, 0] * rpl[:, 0] + rpl[:, 1] * rpl[:, 1] + rpl[:, 2] * rpl[:, 2])
    rmin, rmax = rmag - 1.1 * renc, rmag + 1.1 * renc
    if kind in ("rmin", "rmax"):
        ext = rmin if kind == "rmin" else rmax
        for k, i in enumerate((4, 5, 2)):           # sqrt(x*x) == |x| exactly: |r_tp| equals the extent bit for bit
            rtp[10 + k] = (ext[i], 0.0, 0.0) if k != 1 else (0.0, -ext[i], 0.0)
            vtp[10 + k] = vtp[10 + k] * 0.5
    elif kind == "dup":
        rtp[100:104] = rtp[99]
        vtp[100:104] = vtp[99]
        rtp[200] = rtp[201][[1, 0, 2]] * (1, -1, 1)  # another position, same |r| up to rounding (not forced equal)
    return rpl, p["vh"], rtp, vtp, renc, 0.05
