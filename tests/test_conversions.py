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
    # `U256::from_str` (ruint) also reads un-prefixed decimal and 0o / 0b prefixes
    assert cv.hex_to_u256("29") == 29 and cv.hex_to_u256("0o51") == 41 and cv.hex_to_u256("0b101001") == 41 and cv.hex_to_u256("0X29") == 41
    for bad in ("0xzz", "", "0x", "-1", "4 1", "0x" + "1" + "0" * 64):
        with pytest.raises(cv.HexU256ParseError):
            cv.hex_to_u256(bad)


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


def test_serialize_public_input_orders_follow_the_contract():
    """instance-column order = the order Shielder.sol feeds the verifier (contracts/Shielder.sol:347-370, 505-519, 680-701)"""
    w = {k: cv.field_to_bytes(cv.fr(10 + i)) for i, k in enumerate(
        ["merkle_root", "h_nullifier_old", "h_note_new", "withdrawal_value", "token_address", "commitment", "mac_salt", "mac_commitment"])}
    inst = cv.serialize_public_input("withdraw", w)
    assert inst.shape == (8, 4) and [cv.fr_value(x) for x in inst] == list(range(10, 18))
    d = dict(w); d["value"] = d.pop("withdrawal_value")
    got = [cv.fr_value(x) for x in cv.serialize_public_input("deposit", d)]
    assert got == [10, 11, 12, 13, 15, 14, 16, 17]          # deposit: commitment BEFORE token_address
    assert len(cv.INSTANCE_ORDER["new_account"]) == 13
    with pytest.raises(KeyError):
        cv.serialize_public_input("withdraw", d)
    bad = dict(w); bad["mac_salt"] = b"\xff" * 32            # not a canonical field element
    with pytest.raises(Exception):
        cv.serialize_public_input("withdraw", bad)


def test_commitment_word_and_keccak_known_answer():
    """zkgpu_keccak256 needs no GPU: the reference's own known answer (crates/shielder-account/src/secrets.rs:75-92), and the
    `>> 4` the contract applies so that the commitment is below r"""
    import zkgpu
    m1 = (15).to_bytes(32, "big") + b"nullifier" + (0xFF).to_bytes(4, "big")
    assert zkgpu._keccak256(m1).hex() == "375a07a9503d15a291307e33ad0c297c9768fea4712947172ad09f2df34d8015"
    word = cv.commitment_word(m1)
    assert int.from_bytes(word, "little") == int("375a07a9503d15a291307e33ad0c297c9768fea4712947172ad09f2df34d8015", 16) >> 4
    cv.vec_to_f(word)                                         # below r: a valid public input
