"""CPU tests of the oracle's plain restatements (oracle/port.py) against the compiled reference's
outputs stored in the golden fixtures -- this is what pins the draw-order bookkeeping of supplied-count
mode (the same counts reproduce the seeded reference bit for bit)."""
import numpy as np
import pytest

from oracle import port
from conftest import rel_err
from _stubs import edges_orb, sort_indices


def _mids(gg):
    return [0.5 * (gg[kk][1:] + gg[kk][:-1]) for kk in ("mtot", "mrat", "redz")]


def test_port_integrate(golden):
    got = port.integrate_differential_number_3dx1d(edges_orb(golden), golden["diff_num"])
    assert rel_err(got, golden["number"]) < 1e-14


def test_port_sam_poisson_gwb(golden):
    got = port.sam_poisson_gwb(golden["number"], golden["h2fdf"], golden["counts_gwb"])
    assert np.array_equal(got, golden["gwb_ref"])


def test_port_loudest_plain(golden):
    hc2ss, hc2bg = port.loudest_hc_from_sorted(golden["number"], golden["h2fdf"], int(golden["nloud"]),
                                               golden["order"], golden["counts_loud"])
    assert np.array_equal(hc2ss, golden["l1_hc2ss"])
    assert np.array_equal(hc2bg, golden["l1_hc2bg"])


def test_port_loudest_par(golden):
    mt, mr, rz = _mids(golden)
    out = port.loudest_hc_and_par_from_sorted(golden["number"], golden["h2fdf"], int(golden["nloud"]), mt, mr, rz,
                                              golden["order"], golden["counts_loud"])
    assert np.array_equal(out["hc2ss"], golden["l2_hc2ss"])
    assert np.array_equal(out["ssidx"], golden["l2_ssidx"])
    assert np.array_equal(out["hc2bg"], golden["l2_hc2bg"])
    assert rel_err(out["bgpar"], golden["l2_bgpar"]) < 1e-14
    assert rel_err(out["lspar"], golden["l2_lspar"]) < 1e-14


def test_port_loudest_redz(golden):
    mt, mr, rz = _mids(golden)
    out = port.loudest_hc_and_par_from_sorted_redz(
        golden["number"], golden["h2fdf"], int(golden["nloud"]), mt, mr, rz, golden["par_redz"], golden["par_dcom"],
        golden["par_sepa"], golden["par_angs"], golden["order"], golden["counts_loud"])
    assert np.array_equal(out["hc2ss"], golden["l3_hc2ss"])
    assert np.array_equal(out["sspar"], golden["l3_sspar"])
    assert np.array_equal(out["hc2bg"], golden["l3_hc2bg"])
    assert np.array_equal(np.isnan(out["bgpar"]), np.isnan(golden["l3_bgpar"]))
    assert rel_err(out["bgpar"], golden["l3_bgpar"]) < 1e-14


def test_port_ss_bg(golden):
    mt, mr, rz = _mids(golden)
    out = port.ss_bg_hc_and_par(golden["number"], golden["h2fdf"], mt, mr, rz, golden["counts_ssbg"])
    assert np.array_equal(out["hc2ss"], golden["s2_hc2ss"])
    assert np.array_equal(out["ssidx"], golden["s2_ssidx"])
    assert np.array_equal(out["sspar"], golden["s2_sspar"])
    assert rel_err(out["hc2bg"], golden["s2_hc2bg"]) < 1e-13
    assert rel_err(out["bgpar"], golden["s2_bgpar"]) < 1e-10
    assert np.array_equal(out["hc2ss"], golden["s1_hc2ss"]) and np.array_equal(out["ssidx"], golden["s1_ssidx"])


def test_port_dbn_tiny(golden):
    """Step-major pure-Python restatement of the 2PL / GW dbn loops on a few (M,q) rows."""
    rows = [(0, 0), (golden["mtot"].size // 2, golden["mrat"].size // 2), (golden["mtot"].size - 1, golden["mrat"].size - 1)]
    rz, dn = port.dynamic_binary_number_rows(golden, rows)
    for (ii, jj), rr, dd in zip(rows, rz, dn):
        assert np.array_equal(rr == -1, golden["redz_final"][ii, jj] == -1)
        assert rel_err(rr, golden["redz_final"][ii, jj]) < 1e-13
        assert rel_err(dd, golden["diff_num"][ii, jj]) < 1e-12
