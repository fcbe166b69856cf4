#!/usr/bin/env python3
"""Writes tests/golden/harness_vectors.json: the expected-bytes file of tools/rust_harness (ladder step 5 of SURVEY.md section 8c).
Every vector carries what `RangeProof::prove_multiple_with_rng` needs -- the SHIFTED u64 values and the (padded) blindings of each chunk, the nonce-key
inputs -- and the proof / commitment bytes the oracle produces for them.  tests/test_golden.py checks that the oracle and the GPU library reproduce the
file; the Rust harness checks that the REAL crates do."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

L = 2**252 + 27742317777372353535851937790883648493


def signed(sc_bytes):
    v = int.from_bytes(bytes(sc_bytes), "little")
    return v - L if v > L // 2 else v


def range_vector(name, values, blind_seed, seed, rb, P, nb, fr):
    v = np.asarray(values, np.float32); D = v.size
    bl = oracle.rnd_scalar_vec(blind_seed, D)
    rc, proofs, commits = oracle.range_prove(v, bl, rb, P, nb, fr, seed); assert rc == 0 and oracle.range_verify(proofs, commits, rb, seed) == 1
    Dp = oracle.next_pow2(D); C = min(Dp, P); m = Dp // C
    raw = [signed(s) for s in oracle.f32_to_scalar_vec(v, nb, fr)]
    shifted = [((r + (1 << (rb - 1))) % L) & ((1 << nb) - 1) for r in raw] + [0] * (Dp - D)            # range_proof_vec/mod.rs:36-51
    blind = [bytes(b).hex() for b in bl] + ["00" * 32] * (Dp - D)
    # commitments to the SHIFTED values (what prove_multiple returns): V = C + 2^(rb-1) B for real elements, identity for the padding
    off = oracle.scalarmult_base((1 << (rb - 1)).to_bytes(32, "little"))
    V = [bytes(oracle.point_add(bytes(c), off)).hex() for c in commits] + ["00" * 32] * (Dp - D)
    chunks = [dict(chunk=c, values=shifted[c * m:(c + 1) * m], blindings=blind[c * m:(c + 1) * m], proof=bytes(proofs[c]).hex(), commitments=V[c * m:(c + 1) * m]) for c in range(C)]
    return dict(name=name, label="RangeProof", domain=1, seed=seed.hex(), n=rb, n_bits=nb, frac=fr, n_partition=P, f32_values=[float(x) for x in v], blind_seed=blind_seed.hex(), chunks=chunks)


def l2_vector(name, values, blind_seed, seed, rb, nb, fr):
    v = np.asarray(values, np.float32); D = v.size
    bl = oracle.rnd_scalar_vec(blind_seed, D)
    rc, proof, commit = oracle.l2_prove(v, bl, rb, nb, fr, seed); assert rc == 0 and oracle.l2_verify(proof, commit, rb, seed) == 1
    raw = [signed(s) for s in oracle.f32_to_scalar_vec(v, nb, fr)]
    total = sum(r * r for r in raw) % L & ((1 << nb) - 1)                                                # l2_range_proof_vec/mod.rs:37-42,69-73
    bsum = sum(int.from_bytes(bytes(b), "little") for b in bl) % L
    return dict(name=name, label="L2RangeProof", domain=4, seed=seed.hex(), n=rb, n_bits=nb, frac=fr, f32_values=[float(x) for x in v], blind_seed=blind_seed.hex(),
                chunks=[dict(chunk=0, values=[total], blindings=[bsum.to_bytes(32, "little").hex()], proof=bytes(proof).hex(), commitments=[bytes(commit).hex()])])


def main():
    oracle.build()
    rng = np.random.default_rng(2024)
    vecs = [
        range_vector("range 8-bit, D=5, P=2 (padded to 8)", rng.uniform(-0.99, 0.99, 5), b"\x21" * 32, b"\x22" * 32, 8, 2, 16, 7),
        range_vector("range 16-bit, D=3, P=1 (padded to 4)", rng.uniform(-255.9, 255.9, 3), b"\x23" * 32, b"\x24" * 32, 16, 1, 16, 7),
        range_vector("range 8-bit with the extremes and zero blindings as rofl_service uses them", [0.9921875, -0.9921875, 0.0, 0.5], b"\x00" * 32, b"\x25" * 32, 8, 4, 16, 7),
        l2_vector("L2 sum proof 32-bit, D=6", rng.integers(-24, 25, 6) / 128, b"\x26" * 32, b"\x27" * 32, 32, 32, 7),
    ]
    # zero blindings for the third vector (client.rs:70-72 derive_dummy_blindings): regenerate it with explicit zeros
    v3 = np.array([0.9921875, -0.9921875, 0.0, 0.5], np.float32); z = np.zeros((4, 32), np.uint8)
    rc, proofs, commits = oracle.range_prove(v3, z, 8, 4, 16, 7, b"\x25" * 32); assert rc == 0
    off = oracle.scalarmult_base((1 << 7).to_bytes(32, "little"))
    raw = [signed(s) for s in oracle.f32_to_scalar_vec(v3, 16, 7)]
    vecs[2]["chunks"] = [dict(chunk=c, values=[((raw[c] + 128) % L) & 0xffff], blindings=["00" * 32], proof=bytes(proofs[c]).hex(), commitments=[bytes(oracle.point_add(bytes(commits[c]), off)).hex()]) for c in range(4)]
    vecs[2]["blind_seed"] = None
    gens = [[w, p, i, bytes(oracle.bp_gens(w, p, i + 1)[i]).hex()] for w, p, i in [("G", 0, 0), ("G", 0, 1), ("H", 0, 0), ("G", 1, 0), ("H", 3, 7), ("G", 1023, 15)]]
    out = dict(format="rofl_b200 harness vectors v1", pedersen=[bytes(oracle.basepoint()).hex(), bytes(oracle.blinding_basepoint()).hex()], generators=gens, vectors=vecs)
    path = os.path.join(ROOT, "tests", "golden", "harness_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1); f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
