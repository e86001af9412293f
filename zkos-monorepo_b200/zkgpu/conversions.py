"""Byte-layout contract at the prover boundary — host mirror of `crates/type-conversions`
(/root/reference/crates/type-conversions/lib.rs:34-118, endianess.rs:3-34) and of the witness decoders of the
bindings (`vec_to_f`, `vec_to_path`: /root/reference/crates/shielder_bindings/src/utils.rs:32-60).

A field element on the Python side is what crosses the C ABI: four little-endian u64 limbs in Montgomery form
(numpy uint64[4], the memory of Rust `bn256::Fr`).  U256 values are Python ints, addresses 20-byte `bytes`.
These are scalar host-side codecs (a handful per proof), not part of the GPU path.
"""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_MONT_R = 1 << 256
_MONT_R_INV = pow(_MONT_R, -1, R_MOD)
ARITY, NOTE_TREE_HEIGHT = 7, 13          # crates/shielder-setup/lib.rs:4-5
FR_SIZE = 32


class ConversionError(ValueError):
    pass


class IncorrectVecLength(ConversionError):
    def __init__(self, expected, actual):
        super().__init__("incorrect vec length: expected %d, got %d" % (expected, actual))
        self.expected, self.actual = expected, actual


class Halo2FieldElementCreationFailed(ConversionError):
    def __init__(self):
        super().__init__("halo2 failed to create field element")


class HexU256ParseError(ConversionError):
    def __init__(self):
        super().__init__("failed to parse hex string to U256")


def _to_limbs(v):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def _from_limbs(f):
    f = np.asarray(f, dtype=np.uint64).reshape(4)
    return sum(int(f[i]) << (64 * i) for i in range(4))


def fr(value):
    """`Fr::from(u64)` / from any integer: Montgomery limbs of value mod r"""
    return _to_limbs((int(value) % R_MOD) * _MONT_R % R_MOD)


def fr_value(f):
    """canonical integer of a field element"""
    return _from_limbs(f) * _MONT_R_INV % R_MOD


def u256_to_field(value):
    """lib.rs:35-37 — `F::from(limbs)`: the 256-bit integer reduced into the field"""
    if not 0 <= int(value) < 1 << 256:
        raise ConversionError("not a U256")
    return fr(value)


def field_to_u256(f):
    """lib.rs:40-44 — `U256::from_le_bytes(to_repr())`"""
    return fr_value(f)


def _array(data, length):
    data = bytes(data)
    if len(data) != length:
        raise IncorrectVecLength(length, len(data))
    return data


def bytes_to_field(data):
    """lib.rs:59-66 — `F::from_repr`: 32 little-endian bytes of a CANONICAL value (< r), else an error"""
    v = int.from_bytes(_array(data, FR_SIZE), "little")
    if v >= R_MOD:
        raise Halo2FieldElementCreationFailed()
    return fr(v)


def field_to_bytes(f):
    """lib.rs:69-73 — `to_repr()`: canonical little-endian"""
    return fr_value(f).to_bytes(FR_SIZE, "little")


def hex_to_u256(text):
    """lib.rs:82-84 — `U256::from_str` (ruint): 0x / 0X hex, 0o octal, 0b binary prefixes, otherwise decimal digits"""
    try:
        low = text[:2].lower()
        radix = {"0x": 16, "0o": 8, "0b": 2}.get(low, 10)
        digits = text[2:] if radix != 10 else text
        if not digits or digits[0] in "+-_" or "_" in digits or not digits.isascii():
            raise ValueError
        v = int(digits, radix)
    except ValueError:
        raise HexU256ParseError()
    if v >= 1 << 256:
        raise HexU256ParseError()
    return v


def hex_32_to_f(text):
    return u256_to_field(hex_to_u256(text))


def bytes_to_u256(data):
    return int.from_bytes(_array(data, 32), "little")


def u256_to_bytes(value):
    return int(value).to_bytes(32, "little")


def address_to_u256(address):
    return int.from_bytes(_array(address, 20), "big")


def address_to_field(address):
    """lib.rs:97-102 — `uint256(uint160(address))`"""
    return u256_to_field(address_to_u256(address))


def field_to_address(f):
    """lib.rs:105-113 — low 20 bytes of the big-endian representation"""
    return to_bytes_be(f)[12:]


# ---- Endianess trait (endianess.rs) ----------------------------------------------------------------------
def to_bytes_le(f):
    return field_to_bytes(f)


def to_bytes_be(f):
    return field_to_bytes(f)[::-1]


def from_bytes_le(data):
    return bytes_to_field(data)


def from_bytes_be(data):
    return bytes_to_field(_array(data, FR_SIZE)[::-1])


# ---- witness decoders of the bindings --------------------------------------------------------------------
def vec_to_f(v):
    """utils.rs:32-34"""
    return bytes_to_field(v)


def vec_to_path(v):
    """utils.rs:36-60 — NOTE_TREE_HEIGHT x ARITY field elements from 7 * 13 * 32 = 2912 bytes"""
    v = bytes(v)
    if len(v) != NOTE_TREE_HEIGHT * ARITY * FR_SIZE:
        raise IncorrectVecLength(NOTE_TREE_HEIGHT * ARITY * FR_SIZE, len(v))
    out = np.empty((NOTE_TREE_HEIGHT, ARITY, 4), dtype=np.uint64)
    for i in range(NOTE_TREE_HEIGHT):
        for j in range(ARITY):
            o = (i * ARITY + j) * FR_SIZE
            out[i, j] = bytes_to_field(v[o:o + FR_SIZE])
    return out


# ---- serialize_public_input: the instance column of a proof from the bindings' public-input byte structs -----------
# Field names: `NewAccountPubInputsBytes` / `DepositPubInputsBytes` / `WithdrawPubInputsBytes`
# (/root/reference/crates/shielder_bindings/src/circuits/{new_account.rs:19-33, deposit.rs:18-27, withdraw.rs:18-27}); the ORDER of the
# instance column ("needs to match the order in the circuit") is the one the contract feeds the verifier:
# /root/reference/contracts/Shielder.sol:347-370 (13 inputs), :505-519 (8), :680-701 (8).
INSTANCE_ORDER = {
    "new_account": ("hashed_note", "prenullifier", "initial_deposit", "commitment", "token_address", "anonymity_revoker_public_key_x",
                    "anonymity_revoker_public_key_y", "sym_key_encryption_1_x", "sym_key_encryption_1_y", "sym_key_encryption_2_x",
                    "sym_key_encryption_2_y", "mac_salt", "mac_commitment"),
    "deposit": ("merkle_root", "h_nullifier_old", "h_note_new", "value", "commitment", "token_address", "mac_salt", "mac_commitment"),
    "withdraw": ("merkle_root", "h_nullifier_old", "h_note_new", "withdrawal_value", "token_address", "commitment", "mac_salt", "mac_commitment"),
}


def serialize_public_input(circuit, pub_inputs):
    """`PublicInputProvider::serialize_public_input` over the byte struct of `circuit` ("new_account" / "deposit" / "withdraw"):
    `pub_inputs` maps the struct's field names to canonical little-endian 32-byte words; returns the (num_pi, 4) Montgomery array
    `zkgpu_prove_batch` takes as the proof's instance column.  Missing or unknown fields and non-canonical words are errors."""
    order = INSTANCE_ORDER[circuit]
    extra, missing = set(pub_inputs) - set(order), [k for k in order if k not in pub_inputs]
    if extra or missing:
        raise KeyError("public inputs of %s: missing %s, unknown %s" % (circuit, missing, sorted(extra)))
    return np.stack([vec_to_f(pub_inputs[k]) for k in order])


def commitment_word(packed):
    """the `commitment` public input: keccak256 of the abi.encodePacked call context shifted right by 4 bits so that it is below r
    (/root/reference/contracts/Shielder.sol:351-356, 510-515, 688-699); `packed` = the abi.encodePacked bytes; returns the
    canonical little-endian word the bindings take"""
    from . import _keccak256
    return (int.from_bytes(_keccak256(bytes(packed)), "big") >> 4).to_bytes(32, "little")


# ---- calldata of the on-chain verifier -------------------------------------------------------------------
VERIFY_PROOF_SELECTOR = bytes.fromhex("1e8e1e13")   # keccak256("verifyProof(bytes,uint256[])")[:4]


def encode_calldata(proof, instances):
    """`verifier_contract::encode_calldata(proof, instances)`
    (/root/reference/crates/halo2-verifier/src/lib/verifier_contract.rs:14-20): Solidity ABI encoding of
    `Halo2Verifier.verifyProof(bytes proof, uint256[] instances)` — selector, two head offsets, the proof as
    length-prefixed bytes padded to a word, the instances as length-prefixed big-endian words."""
    proof = bytes(proof)
    word = lambda v: int(v).to_bytes(32, "big")
    padded = proof + b"\x00" * (-len(proof) % 32)
    inst = [field_to_u256(f) for f in instances]
    head = word(0x40) + word(0x40 + 32 + len(padded))
    return VERIFY_PROOF_SELECTOR + head + word(len(proof)) + padded + word(len(inst)) + b"".join(word(v) for v in inst)
