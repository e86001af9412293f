// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
//
// Poseidon2 over BN254 Fr with t = 8 (rate 7, capacity 1), S-box x^7, 8 full + 48 partial rounds: the hash of the
// Shielder note tree (/root/reference/contracts/MerkleTree.sol:7,134-147) and of `shielder_bindings::hash::poseidon_hash`
// (/root/reference/crates/shielder_bindings/src/hash.rs:16-27).  Restated from the reference's own generator of the
// on-chain implementation, /root/reference/poseidon2-solidity/generate_t8.py:
//   init           :546-561  state = (in_0..in_6, tag), tag = 7 * 2^64, then the external linear layer
//   full round r   :563-570  state_i <- (state_i + C[8r+i])^7 for all i, external linear layer      (fr_intro :518-543)
//   partial round r:572-588  state_0 <- (state_0 + C[8r])^7; s = sum(state); state_i <- D_i * state_i + s
//   external layer :486-516  M4 on each half (mm4 :480-498), then state_i += state_i + state_{i+4 mod 8} over halves
//   output         : state_0 (utils.py:141)
// The equality of this function with shielder-circuits' `poseidon::off_circuit::hash::<7>` is what the reference tests at
// /root/reference/crates/integration-tests/src/poseidon2.rs:35-53.
// Pinned by tests/golden/poseidon2_t8.json = outputs of the reference's generated code executed in the build container
// (tests/golden/make_poseidon2_vectors.py).  Inputs shorter than 7 (`hash_variable_length`,
// /root/reference/crates/shielder_bindings/src/utils.rs:14-30) are zero-padded with tag = len * 2^64, the halo2
// `ConstantLength<L>` domain [UPSTREAM-MEMORY: shielder-circuits is not vendored] — parity unpinned for len < 7.
#pragma once
#include "bn254.hpp"
#include "poseidon2_consts.inc"

namespace oracle {

struct Poseidon2 {
    static const int T = 8, RATE = 7;
    Fr rc_full[ORC_P2_RF * 8], rc_part[ORC_P2_RP], diag[8];
    Poseidon2() {
        auto ld = [](const uint64_t* l) { U256 t; memcpy(t.l, l, 32); return Fr::from_u256(t); };
        for (int i = 0; i < ORC_P2_RF * 8; ++i) rc_full[i] = ld(ORC_P2_RC_FULL[i]);
        for (int i = 0; i < ORC_P2_RP; ++i) rc_part[i] = ld(ORC_P2_RC_PART[i]);
        for (int i = 0; i < 8; ++i) diag[i] = ld(ORC_P2_DIAG[i]);
    }
    static const Poseidon2& get() { static const Poseidon2 p; return p; }

    static Fr pow7(const Fr& x) { Fr x2 = x.square(), x4 = x2.square(); return x4 * x2 * x; }   // utils.py:86-93
    // generate_t8.py:480-498
    static void mm4(Fr& a, Fr& b, Fr& c, Fr& d) {
        Fr t0 = a + b, t1 = c + d, t2 = b + b + t1, t3 = d + d + t0;
        Fr t4 = t1 + t1; t4 = t4 + t4 + t3;
        Fr t5 = t0 + t0; t5 = t5 + t5 + t2;
        a = t3 + t5; b = t5; c = t2 + t4; d = t4;
    }
    // generate_t8.py:500-516
    static void external(Fr s[8]) {
        mm4(s[0], s[1], s[2], s[3]);
        mm4(s[4], s[5], s[6], s[7]);
        for (int i = 0; i < 4; ++i) { Fr u = s[i] + s[i + 4]; s[i] = s[i] + u; s[i + 4] = s[i + 4] + u; }
    }
    void permute(Fr s[8]) const {
        external(s);
        const int half = ORC_P2_RF / 2;
        for (int r = 0; r < half; ++r) {
            for (int i = 0; i < 8; ++i) s[i] = pow7(s[i] + rc_full[8 * r + i]);
            external(s);
        }
        for (int r = 0; r < ORC_P2_RP; ++r) {
            s[0] = pow7(s[0] + rc_part[r]);
            Fr sum = s[0];
            for (int i = 1; i < 8; ++i) sum = sum + s[i];
            for (int i = 0; i < 8; ++i) s[i] = diag[i] * s[i] + sum;
        }
        for (int r = half; r < ORC_P2_RF; ++r) {
            for (int i = 0; i < 8; ++i) s[i] = pow7(s[i] + rc_full[8 * r + i]);
            external(s);
        }
    }
    // hash of 1..7 field elements (hash_variable_length)
    Fr hash(const Fr* in, size_t len) const {
        if (len < 1 || len > 7) throw std::runtime_error("poseidon2: input length must be between 1 and 7");
        Fr s[8];
        for (size_t i = 0; i < 7; ++i) s[i] = i < len ? in[i] : Fr::zero();
        U256 tag{{0, (u64)len, 0, 0}};   // len * 2^64
        s[7] = Fr::from_u256(tag);
        permute(s);
        return s[0];
    }
    // Root of a Merkle path laid out as MerkleTree.sol:getMerklePath returns it (:88-113): `height` levels of ARITY = 7
    // siblings, leaf level first.  Level i+1 must contain hash(level i) (what _addNote stores in the parent, :134-147);
    // *consistent reports whether it does at every level.  Returns hash(level height-1).
    Fr merkle_root(const Fr* path, size_t height, bool* consistent) const {
        bool ok = true;
        Fr h = Fr::zero();
        for (size_t l = 0; l < height; ++l) {
            if (l) {
                bool found = false;
                for (int j = 0; j < 7; ++j) found = found || path[7 * l + j] == h;
                ok = ok && found;
            }
            h = hash(path + 7 * l, 7);
        }
        if (consistent) *consistent = ok;
        return h;
    }
};

}  // namespace oracle
