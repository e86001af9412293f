"""Host mirror of crates/powers-of-tau (zkgpu/powers_of_tau.py) on the reference's own fixtures
(resources/ppot_0080_11.ptau and ppot_0080_11_raw, copied to tests/golden/): the parsing half runs on the CPU,
`from_parts` (G1 iFFT = K6) and the on-curve check need the GPU."""
import os

import numpy as np
import pytest

import oracle_lib as O
from zkgpu import powers_of_tau as pot

PTAU = os.path.join(O.GOLDEN, "ppot_0080_11.ptau")


def test_raw_reader_and_roundtrip():
    srs = pot.read(O.RAW11, pot.Format.Raw)
    ref = O.srs_read(O.RAW11, 0)
    assert srs.k == 11 and srs.n == 2048
    for a, b in ((srs.g, ref["g"]), (srs.g_lagrange, ref["g_lagrange"]), (srs.g2, ref["g2"]), (srs.s_g2, ref["s_g2"])):
        assert np.array_equal(a, b)
    assert srs.write_custom() == open(O.RAW11, "rb").read()          # params.bin layout (4 + 2*2048*64 + 256 bytes)
    with pytest.raises(ValueError):
        pot._read_raw(open(O.RAW11, "rb").read()[:-1])


def test_raw_equals_perpetual_parsing():
    """crates/powers-of-tau/lib.rs:267-281 (`raw_equals_perpetual`): k, g, g2, s_g2 of the two formats agree"""
    k, g, g2, s_g2 = pot._read_ptau_parts(open(PTAU, "rb").read())
    raw = pot.read(O.RAW11, pot.Format.Raw)
    assert k == raw.k
    assert np.array_equal(g, raw.g) and np.array_equal(g2, raw.g2) and np.array_equal(s_g2, raw.s_g2)
    assert O.g1_on_curve(g)


def test_file_paths(monkeypatch):
    monkeypatch.delenv("PTAU_RESOURCES_DIR", raising=False)
    assert pot.get_ptau_file_path(13, pot.Format.PerpetualPowersOfTau, "res").endswith(os.path.join("res", "ppot_0080_13.ptau"))
    monkeypatch.setenv("PTAU_RESOURCES_DIR", "/data")
    assert pot.get_ptau_file_path(11, pot.Format.Raw) == "/data/ppot_0080_11_raw"


@pytest.mark.gpu
def test_read_ptau_from_parts_on_gpu():
    """read(.ptau) = from_parts(k, g, None, g2, s_g2): g_lagrange from the GPU G1 iFFT equals the Lagrange block that
    halo2 itself wrote into the raw file; downsize recomputes it; the on-curve check rejects a damaged point."""
    import zkgpu
    zkgpu.init(0)
    srs = pot.read(PTAU, pot.Format.PerpetualPowersOfTau)
    raw = pot.read(O.RAW11, pot.Format.Raw)
    assert np.array_equal(srs.g_lagrange, raw.g_lagrange)
    small = srs.downsize(8)
    assert np.array_equal(small.g_lagrange, O.g_to_lagrange(raw.g[:256], 8))
    # test_commit_lagrange (lib.rs:248-264) through the mirrored API
    a = O.random_fr(3, 2048)
    params = srs.params()
    dom = zkgpu.EvaluationDomain(2, 11)
    assert np.array_equal(params.commit(dom.lagrange_to_coeff(a)), params.commit_lagrange(a))
    params.release()
    bad = np.zeros(1, dtype=np.uint64)
    g = raw.g.copy(); g[5, 0] ^= 1
    zkgpu._chk(zkgpu.lib().zkgpu_g1_on_curve(zkgpu._p(g), 2048, zkgpu._p(bad)))
    assert int(bad[0]) == 1
