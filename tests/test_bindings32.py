"""The reference's own `extern "C"` ABI (rofl_crypto/src/bindings32.rs:29-764) called BY ITS NAMES through ctypes, the way the reference's Python
consumer does: PyVec / PyRes structs, bincode payloads, leaked outputs.  CPU run: librofl_b200_bindings32 linked against the CUDA-on-CPU emulation
of the product library (tests/hostsim, a test tool); `-m gpu` run: the shipped librofl_b200_bindings32.so on the B200.  Results are compared with
the oracle (commitments, decrypted values, verdicts) and with the flat C ABI."""
import ctypes as C
import os
import struct
import subprocess
import numpy as np
import pytest
from conftest import locked_make

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class PyVec(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


class PyRes(C.Structure):
    _fields_ = [("ret", C.c_size_t), ("msg", C.c_char_p), ("res", C.c_void_p)]


SIGS = {
    "add_commitments": (PyVec, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t]),
    "add_commitments_transposed": (PyVec, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t]),
    "commit_no_blinding": (PyVec, [C.c_void_p, C.c_size_t]),
    "commit": (PyVec, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "generate_cancelling_blindings": (PyVec, [C.c_size_t, C.c_size_t]),
    "select_blindings": (PyVec, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "select_commitments": (PyVec, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "extract_values": (PyVec, [C.c_void_p, C.c_size_t]),
    "create_rangeproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "verify_rangeproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]),
    "create_randproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "verify_randproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "create_squarerandproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "verify_squarerandproof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "create_l2proof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "verify_l2proof": (PyRes, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "split_elgamal_pair_vector": (PyVec, [C.c_void_p, C.c_size_t]),
    "join_to_elgamal_pair_vector": (PyVec, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "split_squaretriple_pair_vector": (PyVec, [C.c_void_p, C.c_size_t]),
    "join_to_squaretriple_pair_vector": (PyVec, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "clip_to_range": (PyVec, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "quantize_probabilistic": (PyVec, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "commits_equal": (PyRes, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "equals_neutral_group_element_vec": (PyRes, [C.c_void_p, C.c_size_t]),
    "create_zero_scalar_vector": (PyVec, [C.c_size_t]),
    "create_zero_group_element_vector": (PyVec, [C.c_size_t]),
    "create_random_blinding_vector": (PyVec, [C.c_size_t]),
    "add_scalars": (PyVec, [C.c_void_p, C.c_size_t]),
    "filter_unequal_commits": (PyVec, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "say_hello": (PyVec, []),
    "rofl_bindings32_configure": (None, [C.c_int, C.c_int, C.c_int]),
    "rofl_bindings32_free": (None, [C.c_void_p]),
}


def bind(lib):
    for name, (res, args) in SIGS.items():
        f = getattr(lib, name)                 # AttributeError if one of the reference's exports is missing
        f.restype, f.argtypes = res, args
    return lib


def vec_bytes(v, itemsize=1):
    return C.string_at(v.data, v.len * itemsize) if v.len else b""


def vec_of_vecs(v):
    arr = C.cast(v.data, C.POINTER(PyVec))
    return [arr[i] for i in range(v.len)]


def res_vecs(r):
    assert r.ret == 0, r.msg
    return vec_of_vecs(C.cast(r.res, C.POINTER(PyVec))[0])


def res_bool(r):
    assert r.ret == 0, r.msg
    return bool(C.cast(r.res, C.POINTER(C.c_uint8))[0])


def bc32(items):                               # bincode Vec<Scalar> / Vec<RistrettoPoint>
    a = np.ascontiguousarray(items, np.uint8).reshape(-1, 32)
    return struct.pack("<Q", a.shape[0]) + a.tobytes()


def un32(b):
    n = struct.unpack_from("<Q", b)[0]
    assert len(b) == 8 + 32 * n
    return np.frombuffer(b, np.uint8, 32 * n, 8).reshape(n, 32).copy()


def unbytes(b, w):                             # bincode Vec<serialize_bytes(w)>
    n = struct.unpack_from("<Q", b)[0]
    assert len(b) == 8 + n * (8 + w)
    out = np.zeros((n, w), np.uint8)
    for i in range(n):
        assert struct.unpack_from("<Q", b, 8 + i * (8 + w))[0] == w
        out[i] = np.frombuffer(b, np.uint8, w, 16 + i * (8 + w))
    return out


def buf(b):
    return C.create_string_buffer(bytes(b), len(b))


def f32buf(v):
    return np.ascontiguousarray(v, np.float32)


def run_suite(lib, oracle):
    rng = np.random.default_rng(5)
    D = 6
    v = f32buf((rng.integers(-100, 101, D) / 128))
    bl = oracle.rnd_scalar_vec(b"\x61" * 32, D)
    # ---- commit / commit_no_blinding (bindings32.rs:118-149) == oracle commitments
    c0 = un32(vec_bytes(lib.commit_no_blinding(v.ctypes.data, D)))
    assert (c0 == oracle.commit_f32(v, None, 16, 7)).all()
    blb = buf(bc32(bl))
    c1 = un32(vec_bytes(lib.commit(v.ctypes.data, D, blb, len(blb))))
    assert (c1 == oracle.commit_f32(v, bl, 16, 7)).all()
    # ---- add_commitments (:64) + extract_values (:213): sum of three unblinded vectors decrypts to the exact sum
    xs = [f32buf([0.25, 1.25, -1.5]), f32buf([-0.75, 1.25, -2.0]), f32buf([0.5, 1.25, -3.0])]
    cs = [vec_bytes(lib.commit_no_blinding(x.ctypes.data, 3)) for x in xs]
    bufs = [buf(c) for c in cs]
    ptrs = (C.c_void_p * 3)(*[C.cast(b, C.c_void_p) for b in bufs]); lens = (C.c_size_t * 3)(*[len(c) for c in cs])
    s = vec_bytes(lib.add_commitments(ptrs, lens, 3))
    sb = buf(s)
    f = np.frombuffer(vec_bytes(lib.extract_values(sb, len(s)), 4), np.float32)
    assert f.tolist() == [0.0, 3.75, -6.5]                       # pedersen_ops.rs:138-162
    tr = vec_of_vecs(lib.add_commitments_transposed(ptrs, lens, 3))
    assert len(tr) == 3 and all(t.len == 32 for t in tr)
    assert vec_bytes(tr[0]) == bytes(oracle.commit_f32(np.array([0.0], np.float32), None, 16, 7)[0])        # 0.25 + 1.25 - 1.5 = 0
    # ---- create_rangeproof / verify_rangeproof (:228-287)
    r = lib.create_rangeproof(v.ctypes.data, D, blb, len(blb), 8, 2)
    pv, cv = res_vecs(r)
    proofs_b, commits_b = vec_bytes(pv), vec_bytes(cv)
    assert (un32(commits_b) == c1).all()
    n_proofs = struct.unpack_from("<Q", proofs_b)[0]; plen = struct.unpack_from("<Q", proofs_b, 8)[0]
    assert n_proofs == 2 and plen == 32 * (9 + 2 * 5) and len(proofs_b) == 8 + 2 * (8 + plen)
    proofs = unbytes(proofs_b, plen)
    assert oracle.range_verify(proofs, un32(commits_b), 8) == 1                # accepted by the reference-equivalent verifier
    pb, cb = buf(proofs_b), buf(commits_b)
    assert res_bool(lib.verify_rangeproof(cb, len(commits_b), pb, len(proofs_b), 8)) is True
    bad = bytearray(commits_b); bad[8:40] = commits_b[40:72]
    assert res_bool(lib.verify_rangeproof(buf(bad), len(bad), pb, len(proofs_b), 8)) is False
    rerr = lib.create_rangeproof(f32buf([5.0] * D).ctypes.data, D, blb, len(blb), 8, 2)
    assert rerr.ret == 1 and rerr.msg                              # ValueOutOfRangeError -> PyRes{ret: 1, msg}
    # ---- clip_to_range (:652) then provable
    cl = np.frombuffer(vec_bytes(lib.clip_to_range(f32buf([5.0, -5.0, 0.5]).ctypes.data, 3, 8), 4), np.float32)
    assert cl.tolist() == [0.9921875, -0.9921875, 0.5]
    # ---- create_randproof / verify_randproof (:295-370) and the pair split / join helpers (:555-596)
    rp, eg = res_vecs(lib.create_randproof(v.ctypes.data, D, blb, len(blb)))
    rp_b, eg_b = vec_bytes(rp), vec_bytes(eg)
    pairs = unbytes(eg_b, 64)
    assert (pairs[:, :32] == c1).all() and (pairs[:, 32:] == oracle.elgamal_R(bl)).all()
    assert oracle.rand_verify(unbytes(rp_b, 128), pairs) == 1
    egb = buf(eg_b)
    Lv, Rv = vec_of_vecs(lib.split_elgamal_pair_vector(egb, len(eg_b)))
    L_b, R_b = vec_bytes(Lv), vec_bytes(Rv)
    assert (un32(L_b) == pairs[:, :32]).all() and (un32(R_b) == pairs[:, 32:]).all()
    Lb, Rb, rpb = buf(L_b), buf(R_b), buf(rp_b)
    assert vec_bytes(lib.join_to_elgamal_pair_vector(Lb, len(L_b), Rb, len(R_b))) == eg_b
    assert res_bool(lib.verify_randproof(Lb, len(L_b), Rb, len(R_b), rpb, len(rp_b))) is True
    assert res_bool(lib.verify_randproof(Rb, len(R_b), Lb, len(L_b), rpb, len(rp_b))) is False
    # ---- create_l2proof / verify_l2proof (:441-552) and the triple split / join (:598-649); fp 32 / frac 7 like the L2 experiments
    lib.rofl_bindings32_configure(32, 7, 0)
    v2 = f32buf(rng.integers(-24, 25, D) / 128); b2 = oracle.rnd_scalar_vec(b"\x62" * 32, D); b2b = buf(bc32(b2))
    sp, sc, rpf, sq = res_vecs(lib.create_l2proof(v2.ctypes.data, D, blb, len(blb), b2b, len(b2b), 32, 1))
    sp_b, sc_b, rpf_b, sq_b = vec_bytes(sp), vec_bytes(sc), vec_bytes(rpf), vec_bytes(sq)
    assert len(rpf_b) == 616 and len(sq_b) == 32                   # the 616 bytes verify_l2proof hard-codes in the reference (bindings32.rs:524)
    com = unbytes(sc_b, 96)
    assert (com[:, :32] == oracle.commit_f32(v2, bl, 32, 7)).all()
    assert oracle.square_rand_verify(unbytes(sp_b, 192), com) == 1 and oracle.l2_verify(np.frombuffer(rpf_b[8:], np.uint8), np.frombuffer(sq_b, np.uint8), 32) == 1
    scb, spb, rpfb, sqb = buf(sc_b), buf(sp_b), buf(rpf_b), buf(sq_b)
    assert res_bool(lib.verify_l2proof(scb, len(sc_b), spb, len(sp_b), rpfb, sqb, 32)) is True
    assert res_bool(lib.verify_squarerandproof(scb, len(sc_b), spb, len(sp_b))) is True
    a_, b_, c_ = vec_of_vecs(lib.split_squaretriple_pair_vector(scb, len(sc_b)))
    ab, bb, cb2 = vec_bytes(a_), vec_bytes(b_), vec_bytes(c_)
    assert vec_bytes(lib.join_to_squaretriple_pair_vector(buf(ab), len(ab), buf(bb), len(bb), buf(cb2), len(cb2))) == sc_b
    wrong = bytearray(sq_b); wrong[:] = bytes(oracle.basepoint())
    rbad = lib.verify_l2proof(scb, len(sc_b), spb, len(sp_b), rpfb, buf(wrong), 32)
    assert rbad.ret == 1                                           # L2RangeProofError::SumError
    lib.rofl_bindings32_configure(16, 7, 0)
    # ---- blindings: cancelling vectors sum to zero (:154-167, pedersen_ops.rs:110-122), selectors, add_scalars, zero / random vectors
    gv = vec_of_vecs(lib.generate_cancelling_blindings(3, 4))
    vs = [un32(vec_bytes(g)) for g in gv]
    L = 2**252 + 27742317777372353535851937790883648493
    for j in range(4):
        assert sum(int.from_bytes(x[j].tobytes(), "little") for x in vs) % L == 0
    g0 = vec_bytes(gv[0]); idx = (C.c_size_t * 2)(3, 1)
    sel = un32(vec_bytes(lib.select_blindings(buf(g0), len(g0), idx, 2)))
    assert (sel[0] == vs[0][3]).all() and (sel[1] == vs[0][1]).all()
    selc = un32(vec_bytes(lib.select_commitments(cb, len(commits_b), idx, 2)))
    assert (selc[0] == c1[3]).all() and (selc[1] == c1[1]).all()
    tot = vec_bytes(lib.add_scalars(buf(g0), len(g0)))
    assert int.from_bytes(tot, "little") == sum(int.from_bytes(x.tobytes(), "little") for x in vs[0]) % L
    assert un32(vec_bytes(lib.create_zero_scalar_vector(3))).sum() == 0 and un32(vec_bytes(lib.create_zero_group_element_vector(2))).sum() == 0
    r1, r2 = un32(vec_bytes(lib.create_random_blinding_vector(5))), un32(vec_bytes(lib.create_random_blinding_vector(5)))
    assert r1.shape == (5, 32) and (r1 != r2).any() and all(int.from_bytes(x.tobytes(), "little") < L for x in r1)
    # ---- comparisons (:675-764)
    assert res_bool(lib.commits_equal(cb, cb, len(commits_b))) is True and res_bool(lib.commits_equal(cb, buf(bad), len(commits_b))) is False
    z = vec_bytes(lib.create_zero_group_element_vector(2)); assert res_bool(lib.equals_neutral_group_element_vec(buf(z), len(z))) is True
    assert res_bool(lib.equals_neutral_group_element_vec(cb, len(commits_b))) is False
    lv, rv = vec_of_vecs(lib.filter_unequal_commits(cb, buf(bad), len(commits_b)))
    assert un32(vec_bytes(lv)).shape[0] == 1 and (un32(vec_bytes(lv))[0] == c1[0]).all() and (un32(vec_bytes(rv))[0] == c1[1]).all()
    assert un32(vec_bytes(lib.say_hello())).shape == (1, 32)
    lib.rofl_bindings32_free(lv.data)


def test_bindings32_exports_by_reference_names_emulated(oracle):
    d = os.path.join(HERE, "hostsim")
    locked_make(d, "libemul.so", "libbindings32_emul.so", env={"CXX": "g++"})
    run_suite(bind(C.CDLL(os.path.join(d, "libbindings32_emul.so"))), oracle)


def test_shipped_bindings32_library_exports_every_reference_symbol():
    """No compute (no GPU here): the shipped library exists after build() and exports every `#[no_mangle] pub extern "C"` name of bindings32.rs."""
    path = os.path.join(ROOT, "rofl-project-code_b200", "librofl_b200_bindings32.so")
    locked_make(os.path.join(ROOT, "rofl-project-code_b200", "csrc"), jobs=8)
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    names = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = [n for n in SIGS if n not in names]
    assert not missing, missing


@pytest.mark.gpu
def test_bindings32_exports_by_reference_names_on_gpu(oracle):
    run_suite(bind(C.CDLL(os.path.join(ROOT, "rofl-project-code_b200", "librofl_b200_bindings32.so"))), oracle)
