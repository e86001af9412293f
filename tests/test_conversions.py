"""Restatement of the reference's own unit tests for the byte-layout contract
(/root/reference/crates/type-conversions/lib.rs:121-214, endianess.rs:36-64) on the host mirror, cross-checked with
the oracle's Montgomery arithmetic."""
import numpy as np
import pytest

import oracle_lib as O
import pyref as P
from zkgpu import conversions as cv

BYTES_41 = bytes([41] + [0] * 31)
ADDR_41 = bytes(19) + bytes([0x29])
HEX_41 = "0x0000000000000000000000000000000000000000000000000000000000000029"


def test_between_field_and_u256():
    field = cv.fr(41)
    assert np.array_equal(cv.u256_to_field(41), field) and cv.field_to_u256(field) == 41
    assert cv.field_to_u256(cv.u256_to_field(41)) == 41


def test_between_field_and_bytes():
    field = cv.fr(41)
    assert np.array_equal(cv.bytes_to_field(BYTES_41), field)
    assert cv.field_to_bytes(field) == BYTES_41
    assert np.array_equal(cv.bytes_to_field(cv.field_to_bytes(field)), field)


def test_hex_and_u256_and_bytes():
    assert cv.hex_to_u256(HEX_41) == 41
    assert np.array_equal(cv.hex_32_to_f(HEX_41), cv.fr(41))
    assert cv.bytes_to_u256(BYTES_41) == 41 and cv.u256_to_bytes(41) == BYTES_41
    with pytest.raises(cv.HexU256ParseError):
        cv.hex_to_u256("29")
    with pytest.raises(cv.HexU256ParseError):
        cv.hex_to_u256("0xzz")


def test_between_address_and_field():
    field = cv.fr(41)
    assert np.array_equal(cv.address_to_field(ADDR_41), field)
    assert cv.field_to_address(field) == ADDR_41
    assert cv.address_to_u256(ADDR_41) == 41


def test_endianess():
    element = cv.fr(7)
    assert np.array_equal(cv.from_bytes_be(cv.to_bytes_be(element)), element)
    assert np.array_equal(cv.from_bytes_le(cv.to_bytes_le(element)), element)
    assert cv.to_bytes_be(element) == cv.to_bytes_le(element)[::-1]


def test_errors_and_montgomery_layout():
    with pytest.raises(cv.IncorrectVecLength):
        cv.bytes_to_field(b"\x01" * 31)
    with pytest.raises(cv.Halo2FieldElementCreationFailed):          # from_repr rejects non-canonical values
        cv.bytes_to_field(cv.R_MOD.to_bytes(32, "little"))
    assert cv.field_to_u256(cv.u256_to_field(cv.R_MOD + 5)) == 5        # From<[u64; 4]> reduces
    # the limbs are the oracle's (and Rust's) Montgomery memory layout
    vals = [0, 1, 41, cv.R_MOD - 1, 1 << 200]
    assert np.array_equal(np.stack([cv.fr(v) for v in vals]), O.to_mont(0, P.int_to_limbs(vals)))


def test_vec_to_path():
    rng = np.random.default_rng(1)
    vals = [int.from_bytes(rng.bytes(31), "little") for _ in range(cv.ARITY * cv.NOTE_TREE_HEIGHT)]
    raw = b"".join(v.to_bytes(32, "little") for v in vals)
    assert len(raw) == 2912
    path = cv.vec_to_path(raw)
    assert path.shape == (13, 7, 4)
    assert [cv.fr_value(path[i, j]) for i in range(13) for j in range(7)] == vals
    assert np.array_equal(cv.vec_to_f(raw[:32]), path[0, 0])
    with pytest.raises(cv.IncorrectVecLength):
        cv.vec_to_path(raw[:-32])


def test_encode_calldata():
    """ABI layout the generated verifier checks (templates/Halo2Verifier.sol:75-83, 240-263): selector, offsets 0x40 and
    0x40 + 0x20 + len(proof), proof length, instance count, instance words."""
    assert cv.VERIFY_PROOF_SELECTOR == O.keccak256(b"verifyProof(bytes,uint256[])")[:4]
    proof = bytes(range(256)) * 19 + bytes(224)          # 5088 bytes: the withdraw shape's proof length (a word multiple)
    inst = [cv.fr(v) for v in (1, 2, cv.R_MOD - 1)]
    cd = cv.encode_calldata(proof, inst)
    assert cd[:4] == cv.VERIFY_PROOF_SELECTOR
    w = lambda i: int.from_bytes(cd[4 + 32 * i: 36 + 32 * i], "big")
    assert w(0) == 0x40 and w(1) == 0x40 + 0x20 + len(proof) and w(2) == len(proof)
    assert cd[4 + 96: 4 + 96 + len(proof)] == proof
    base = 3 + len(proof) // 32
    assert [w(base), w(base + 1), w(base + 2), w(base + 3)] == [3, 1, 2, cv.R_MOD - 1]
    assert len(cd) == 4 + 32 * (base + 4)
    # unaligned proofs are zero-padded to a word boundary
    cd2 = cv.encode_calldata(b"\x01\x02\x03", [])
    assert len(cd2) == 4 + 32 * 5 and cd2[4 + 96: 4 + 128] == b"\x01\x02\x03" + bytes(29)
