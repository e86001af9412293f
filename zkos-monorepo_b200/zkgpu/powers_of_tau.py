"""Readers for the SRS files Shielder ships — the host mirror of `crates/powers-of-tau`
(/root/reference/crates/powers-of-tau/lib.rs:27-231) plus the params.bin (de)serialisation the hosts use
(`marshall_params` = `ParamsKZG::write_custom(RawBytes)`, /root/reference/crates/shielder_bindings/build.rs:24-31).

  Format.Raw                   halo2 `SerdeFormat::RawBytes`: k:u32 LE | g[n] | g_lagrange[n] | g2 | s_g2
  Format.PerpetualPowersOfTau  snarkjs .ptau: header size (u64) at byte 16, k (u32) at 24 + hs - 8, tau^i G1 at
                               24 + hs + 12, tau^i G2 at that + 64 (2n - 1) + 12; g_lagrange is NOT in the file:
                               `from_parts(k, g, None, g2, s_g2)` derives it with g_to_lagrange — here on the GPU (K6).
All coordinates are 32-byte little-endian Montgomery residues, i.e. already the in-memory layout of halo2curves
`Fq` (lib.rs:92-109 multiplies the `from_repr` value by R^-1 for exactly that reason), so arrays are views of the file.
"""
import enum
import os
import struct

import numpy as np

HEADER_SIZE_OFFSET = 16
HEADER_OFFSET = HEADER_SIZE_OFFSET + 8


class Format(enum.Enum):
    Raw = 0
    PerpetualPowersOfTau = 1


def get_ptau_file_path(k, fmt, default_dir=None):
    """lib.rs:40-55: $PTAU_RESOURCES_DIR first, else the crate-relative resources directory"""
    name = "ppot_0080_%d_raw" % k if fmt is Format.Raw else "ppot_0080_%d.ptau" % k
    return os.path.join(os.environ.get("PTAU_RESOURCES_DIR") or default_dir or "resources", name)


class Srs:
    """ParamsKZG<Bn256> as plain arrays: g, g_lagrange (n, 8) uint64; g2, s_g2 (16,) uint64 (x.c0, x.c1, y.c0, y.c1)."""

    def __init__(self, k, g, g_lagrange, g2, s_g2):
        self.k, self.n = int(k), 1 << int(k)
        self.g, self.g_lagrange, self.g2, self.s_g2 = g, g_lagrange, g2, s_g2

    @staticmethod
    def from_parts(k, g, g_lagrange, g2, s_g2):
        """`ParamsKZG::from_parts`: g_lagrange = None triggers g_to_lagrange (G1 iFFT on the GPU)"""
        import zkgpu
        g = np.ascontiguousarray(g, dtype=np.uint64).reshape(-1, 8)
        if g.shape[0] != 1 << k:
            raise ValueError("from_parts: g must hold 2^k points")
        if g_lagrange is None:
            g_lagrange = zkgpu.g_to_lagrange(g, k)
        return Srs(k, g, np.ascontiguousarray(g_lagrange, dtype=np.uint64).reshape(-1, 8), g2, s_g2)

    def downsize(self, k):
        """`ParamsKZG::downsize`: truncate g, recompute g_lagrange"""
        if k > self.k:
            raise ValueError("downsize: k larger than the parameters")
        if k == self.k:
            return self
        return Srs.from_parts(k, self.g[: 1 << k].copy(), None, self.g2, self.s_g2)

    def params(self):
        """device-resident ParamsKZG (fixed-base tables for commit / commit_lagrange)"""
        import zkgpu
        return zkgpu.ParamsKZG(self.k, self.g, self.g_lagrange)

    def write_custom(self):
        """RawBytes serialisation (params.bin)"""
        return b"".join([struct.pack("<I", self.k), self.g.tobytes(), self.g_lagrange.tobytes(), self.g2.tobytes(), self.s_g2.tobytes()])


def _read_raw(buf):
    if len(buf) < 4:
        raise ValueError("raw srs: short file")
    (k,) = struct.unpack_from("<I", buf, 0)
    if k > 28:
        raise ValueError("raw srs: k out of range")
    n = 1 << k
    if len(buf) != 4 + 2 * n * 64 + 256:
        raise ValueError("raw srs: size mismatch")
    a = np.frombuffer(buf, dtype=np.uint64, offset=4)
    g = a[: 8 * n].reshape(n, 8).copy()
    gl = a[8 * n: 16 * n].reshape(n, 8).copy()
    return Srs(k, g, gl, a[16 * n: 16 * n + 16].copy(), a[16 * n + 16: 16 * n + 32].copy())


def _read_ptau_parts(buf):
    (hs,) = struct.unpack_from("<Q", buf, HEADER_SIZE_OFFSET)
    (k,) = struct.unpack_from("<I", buf, HEADER_OFFSET + hs - 8)
    if k > 28:
        raise ValueError("ptau: k out of range")
    n = 1 << k
    g1_off = HEADER_OFFSET + hs + 12
    g2_off = g1_off + 2 * 32 * (2 * n - 1) + 12
    if len(buf) < g2_off + 256:
        raise ValueError("ptau: short file")
    g = np.frombuffer(buf, dtype=np.uint64, count=8 * n, offset=g1_off).reshape(n, 8).copy()
    g2s = np.frombuffer(buf, dtype=np.uint64, count=32, offset=g2_off).copy()
    return k, g, g2s[:16], g2s[16:]


def read(path, fmt):
    """`powers_of_tau::read(ptau_file, format)` (lib.rs:61-74)"""
    with open(path, "rb") as f:
        buf = f.read()
    if fmt is Format.Raw:
        return _read_raw(buf)
    import zkgpu
    k, g, g2, s_g2 = _read_ptau_parts(buf)
    bad = np.zeros(1, dtype=np.uint64)
    zkgpu._chk(zkgpu.lib().zkgpu_g1_on_curve(zkgpu._p(g), g.shape[0], zkgpu._p(bad)))
    if int(bad[0]):
        raise ValueError("ptau: %d points are not on the curve" % int(bad[0]))   # from_xy(..).unwrap()
    return Srs.from_parts(k, g, None, g2, s_g2)
