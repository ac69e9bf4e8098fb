"""The rule of the sort-free pl-tp sweep (tests/pltp_rule.py = numpy restatement of pltp_direct_kernel) against the
oracle's sort-and-sweep: the same pair list in the same order and the same nbox, including particles that sit exactly on
a planet's inner extent; particles exactly on an OUTER extent must raise the "sort path decides" flag."""
import numpy as np
import pytest

from swiftest_b200 import workloads as W
from tests import pltp_rule as R


def _compare(oracle, args, expect_flag=False):
    i1, i2, nbox, flag = R.direct_pltp(*args)
    r1, r2, _ = oracle.encounter_pltp(*args[:4], args[4], args[5])
    assert flag == expect_flag
    if not flag:
        assert np.array_equal(i1, r1) and np.array_equal(i2, r2)
        assert nbox + R.pairless_particle_boxes(args[2]) == oracle.nbox_total()
    return len(r1)


def test_rule_matches_sort_and_sweep_on_the_reference_fixture(oracle):
    f = W.fixture("108pl_50tp")
    for boost in (1.0, 3.0, 10.0):
        n = _compare(oracle, (f["pl_rh"], f["pl_vh"], f["tp_rh"], f["tp_vh"], f["pl_rhill"] * 6.5 * boost, 0.05))
    assert n > 0


@pytest.mark.parametrize("ntp", [1, 2, 50, 20000])
def test_rule_matches_sort_and_sweep_8_planets(oracle, ntp):
    p = W.planets8_year_units()
    tp = W.tp_cloud(ntp, seed=ntp)
    n = _compare(oracle, (p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 6.5, 0.05))
    if ntp == 20000:
        assert n > 100


def test_rule_with_particles_exactly_on_an_inner_extent_and_duplicates(oracle):
    assert _compare(oracle, R.tie_case("rmin")) > 0
    assert _compare(oracle, R.tie_case("dup")) > 0


def test_rule_flags_particles_exactly_on_an_outer_extent(oracle):
    _compare(oracle, R.tie_case("rmax"), expect_flag=True)


def test_rule_planet_extent_ties_and_degenerate_planets(oracle):
    """Two planets with identical extents, a planet with renc = 0, a planet with negative renc (rmin > rmax)."""
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=8)
    rpl, vpl, renc = p["rh"].copy(), p["vh"].copy(), p["rhill"] * 6.5
    rpl[3] = rpl[2][[1, 0, 2]]          # same |r| ...
    rpl[3] = rpl[2]                      # ... exactly: identical position
    renc[3] = renc[2]
    renc[6] = 0.0
    renc[7] = -renc[7]
    _compare(oracle, (rpl, vpl, tp["rh"], tp["vh"], renc, 0.05))
