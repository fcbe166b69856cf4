"""The product's __host__ __device__ arithmetic (radix-2^25.5 field, Montgomery scalars, ristretto, Keccak/Merlin,
ChaCha, fixed-base tables, ladders, square proof) compiled for the CPU with limb-bound assertions enabled
(tests/hostsim) and compared with the oracle on the same inputs.  CPU only; the GPU tests run the same code on
the device through the C ABI."""
import ctypes as C
import hashlib
import os
import subprocess
import numpy as np
import pytest
from conftest import locked_make

HERE = os.path.dirname(os.path.abspath(__file__))
L = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19


@pytest.fixture(scope="module")
def hs():
    d = os.path.join(HERE, "hostsim")
    locked_make(d, env={"CXX": "g++"})
    return C.CDLL(os.path.join(d, "libhostsim.so"))


def call(lib, name, outlen, *args):
    o = C.create_string_buffer(outlen)
    rc = getattr(lib, name)(o, *[C.c_char_p(bytes(a)) if isinstance(a, (bytes, bytearray)) else a for a in args])
    return rc, o.raw


def rs(rng): return (int.from_bytes(rng.bytes(32), "little") % L).to_bytes(32, "little")
def rf(rng): return (int.from_bytes(rng.bytes(32), "little") % P).to_bytes(32, "little")


def test_field(hs, oracle):
    rng = np.random.default_rng(0)
    edge = [0, 1, 2, 19, P - 1, P - 2, P - 19, 2**255 - 20, (1 << 254), (1 << 26) - 1, ((1 << 255) - 1) % P]
    vals = [x.to_bytes(32, "little") for x in edge] + [rf(rng) for _ in range(300)]
    for i in range(len(vals) - 3):
        a, b, c, d = vals[i:i + 4]
        ai, bi, ci, di = [int.from_bytes(x, "little") for x in (a, b, c, d)]
        assert int.from_bytes(call(hs, "hs_fe_mul", 32, a, b)[1], "little") == ai * bi % P
        assert int.from_bytes(call(hs, "hs_fe_sq", 32, a)[1], "little") == ai * ai % P
        assert int.from_bytes(call(hs, "hs_fe_mix", 32, a, b, c, d)[1], "little") == (ai + bi) * (ci - di) % P
        if ai:
            assert int.from_bytes(call(hs, "hs_fe_invert", 32, a)[1], "little") == pow(ai, P - 2, P)
    # non-canonical input 2^255-1 (bit 255 ignored, value >= p)
    top = bytes([0xff] * 32)
    assert int.from_bytes(call(hs, "hs_fe_mul", 32, top, (1).to_bytes(32, "little"))[1], "little") == (2**255 - 1) % P


def test_scalars(hs, oracle):
    rng = np.random.default_rng(1)
    edge = [0, 1, L - 1, L - 2, 2**252, 2**252 - 1]
    vals = [x.to_bytes(32, "little") for x in edge] + [rs(rng) for _ in range(300)]
    for a, b in zip(vals, vals[1:]):
        ai, bi = int.from_bytes(a, "little"), int.from_bytes(b, "little")
        assert int.from_bytes(call(hs, "hs_sc_mul", 32, a, b)[1], "little") == ai * bi % L
        assert int.from_bytes(call(hs, "hs_sc_add", 32, a, b)[1], "little") == (ai + bi) % L
        assert int.from_bytes(call(hs, "hs_sc_sub", 32, a, b)[1], "little") == (ai - bi) % L
    for w in [bytes(64), bytes([0xff] * 64)] + [rng.bytes(64) for _ in range(200)]:
        assert call(hs, "hs_sc_wide", 32, w)[1] == oracle.sc_reduce_wide(w)
    for a in vals[1:20]:
        assert call(hs, "hs_sc_invert", 32, a)[1] == oracle.sc_invert(a)
    for a in vals[1:6] + vals[6:] + [(2**k).to_bytes(32, "little") for k in (1, 31, 32, 64, 251)]:      # binary-Euclid inverse (Fiat-Shamir challenges)
        ai = int.from_bytes(a, "little")
        assert int.from_bytes(call(hs, "hs_sc_invert_vartime", 32, a)[1], "little") == pow(ai, L - 2, L)
    assert call(hs, "hs_sc_invert_vartime", 32, bytes(32))[1] == bytes(32)
    # division-step inversion (62-bit signed limbs): limb boundaries, both ends of the range, many random values
    more = [1, 2, L - 1, L - 2, (L + 1) // 2, 2**62 - 1, 2**62, 2**124, 2**186, 2**248, 2**252] + [int(rng.integers(1, 2**62)) << (62 * k) for k in range(4) for _ in range(8)]
    more += [int.from_bytes(rng.bytes(32), "little") % L or 1 for _ in range(3000)] + [(int.from_bytes(rng.bytes(32), "little") % 2**k) or 1 for k in range(1, 253, 3) for _ in range(3)]
    for ai in more:
        assert int.from_bytes(call(hs, "hs_sc_invert_vartime", 32, (ai % L).to_bytes(32, "little"))[1], "little") == pow(ai % L, L - 2, L)


def test_ristretto(hs, oracle):
    rng = np.random.default_rng(2)
    pts = [oracle.from_uniform_bytes(rng.bytes(64)) for _ in range(30)] + [oracle.basepoint(), bytes(32)]
    for u in [rng.bytes(64) for _ in range(60)] + [bytes(64), bytes([0xff] * 64)]:
        assert call(hs, "hs_from_uniform", 32, u)[1] == oracle.from_uniform_bytes(u)
    for p in pts:
        rc, o = call(hs, "hs_decompress_compress", 32, p)
        assert rc == 1 and o == p
    for e in [rng.bytes(32) for _ in range(200)] + [P.to_bytes(32, "little"), bytes([1] + [0] * 31), bytes([0xff] * 32)]:
        assert bool(call(hs, "hs_decompress_compress", 32, e)[0]) == oracle.point_valid(e)
    for p, q in zip(pts, pts[1:]):
        assert call(hs, "hs_point_add", 32, p, q, 0)[1] == oracle.point_add(p, q)
        assert call(hs, "hs_point_add", 32, p, q, 1)[1] == oracle.point_sub(p, q)
        assert call(hs, "hs_point_dbl", 32, p, 5)[1] == oracle.scalarmult((32).to_bytes(32, "little"), p)
    assert hs.hs_point_eq(pts[0], pts[0]) == 1 and hs.hs_point_eq(pts[0], pts[1]) == 0


def test_ladders_and_fold(hs, oracle):
    rng = np.random.default_rng(3)
    for w in (4, 5):
        for _ in range(6):
            p, q, s = oracle.from_uniform_bytes(rng.bytes(64)), oracle.from_uniform_bytes(rng.bytes(64)), rs(rng)
            assert call(hs, "hs_scalarmult_naf", 32, s, p, w)[1] == oracle.scalarmult(s, p)
            assert call(hs, "hs_fold", 32, s, q, p, w)[1] == oracle.point_add(q, oracle.scalarmult(s, p))
    p = oracle.basepoint()
    for s in [0, 1, 2, 3, 15, 16, 17, L - 1, 2**252]:
        sb = s.to_bytes(32, "little")
        assert call(hs, "hs_scalarmult_naf", 32, sb, p, 5)[1] == oracle.scalarmult(sb, p)
        assert call(hs, "hs_scalarmult_naf", 32, sb, bytes(32), 4)[1] == bytes(32)


def test_fixed_base(hs, oracle):
    rng = np.random.default_rng(4)
    H = oracle.blinding_basepoint()
    assert hs.hs_fixed_table(H) == 1
    for s in [0, 1, 127, 128, 129, 255, 256, 2**16 - 1, 2**64 - 1, L - 1, 2**252, int.from_bytes(bytes([0x80] * 31 + [0x0f]), "little")]:
        sb = s.to_bytes(32, "little")
        assert call(hs, "hs_fixed_mul", 32, sb)[1] == oracle.scalarmult(sb, H)
    for _ in range(20):
        sb = rs(rng)
        assert call(hs, "hs_fixed_mul", 32, sb)[1] == oracle.scalarmult(sb, H)


def test_hashes(hs, oracle):
    rng = np.random.default_rng(5)
    for n in [0, 1, 71, 72, 135, 136, 137, 500]:
        m = rng.bytes(n)
        assert call(hs, "hs_sha3_512", 64, m, C.c_size_t(n))[1] == hashlib.sha3_512(m).digest()
        o = C.create_string_buffer(400); hs.hs_shake256(o, C.c_size_t(400), C.c_char_p(m), C.c_size_t(n))
        assert o.raw == hashlib.shake_256(m).digest(400)
    o = C.create_string_buffer(32)
    hs.hs_merlin_simple(o, C.c_size_t(32), b"test protocol", b"some label", b"some data", C.c_size_t(9), b"challenge")
    assert o.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    key = rng.bytes(32)
    for ctr in [0, 1, 2**32 - 1, 2**32, 2**40 + 5]:
        assert call(hs, "hs_chacha", 64, key, C.c_uint64(ctr))[1] == oracle.chacha20_block(key, ctr)
    for which, party, n in [("G", 0, 8), ("H", 0, 8), ("G", 77, 16), ("H", 300000, 8)]:
        o = C.create_string_buffer(32 * n); hs.hs_gen_chain(o, ord(which), C.c_uint32(party), n)
        assert o.raw == oracle.bp_gens(which, party, n).tobytes()


def test_f32_conversion(hs, oracle):
    rng = np.random.default_rng(6)
    xs = np.concatenate([rng.uniform(-600, 600, 500), rng.uniform(-2, 2, 500), [0.0, -0.0, 0.5 / 128, 1.5 / 128, 2.5 / 128, 511.99, 512.0, 1e9, -1e9, 3.4e38, 1e-40, 2.0**25 + 2]]).astype(np.float32)
    for nb, fr in [(8, 7), (16, 7), (32, 7), (64, 7), (16, 0), (32, 12)]:
        ref = oracle.f32_to_scalar_vec(np.abs(xs), nb, fr)
        for x, r in zip(xs, ref):
            raw = C.c_uint64()
            assert hs.hs_f32_to_raw(C.byref(raw), C.c_float(float(x)), nb, fr) == 0
            assert raw.value == int.from_bytes(r.tobytes(), "little"), (x, nb, fr)


def test_square_proof_one(hs, oracle):
    rng = np.random.default_rng(7)
    B, H = oracle.basepoint(), oracle.blinding_basepoint()
    D = 4
    v = (rng.integers(-24, 25, D) / 128).astype(np.float32)
    r1 = oracle.rnd_scalar_vec(b"\x21" * 32, D); r2 = oracle.rnd_scalar_vec(b"\x22" * 32, D)
    cl = oracle.commit_f32(v, r1, 32, 7)
    seed = b"\x23" * 32
    rc, proofs, commits = oracle.square_prove(v, cl, r1, r2, 32, 7, seed)
    key = oracle.derive_key(seed, 3, 0)
    for i in range(D):
        p = C.create_string_buffer(160); c = C.create_string_buffer(64)
        hs.hs_square_prove_one(p, c, C.c_float(float(v[i])), cl[i].tobytes(), r1[i].tobytes(), r2[i].tobytes(), key, C.c_uint64(i), 32, 7, B, H)
        assert p.raw == proofs[i].tobytes() and c.raw == commits[i].tobytes()
        assert hs.hs_square_verify_one(p.raw, c.raw, B, H) == 1
        bad = bytearray(p.raw); bad[70] ^= 1
        assert hs.hs_square_verify_one(bytes(bad), c.raw, B, H) == 0
        bad = bytearray(p.raw); bad[64:96] = b"\xff" * 32
        assert hs.hs_square_verify_one(bytes(bad), c.raw, B, H) == -1
