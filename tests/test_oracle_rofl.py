"""The reference's own literal-value / verdict unit tests, re-expressed against the CPU oracle.
Each test names the reference test it mirrors (paths relative to rofl_crypto/src).  CPU only."""
import numpy as np
import pytest

L = 2**252 + 27742317777372353535851937790883648493
NB, FR = 16, 7          # the reference's default cargo features (fp.rs:118-137)


def sc_int(b):
    return int.from_bytes(bytes(b), "little")


# ---- conversion32.rs tests ---------------------------------------------------------------------
def test_conversion_lossless(oracle):                       # conversion32.rs:182-194
    vals = np.array([0.5, -1.25, 65535 / 128], np.float32)
    s = oracle.f32_to_scalar_vec(vals, NB, FR)
    assert sc_int(s[0]) == 64 and sc_int(s[1]) == L - 160 and sc_int(s[2]) == 65535
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == vals).all()


def test_conversion_lossy_rounded(oracle):                  # conversion32.rs:197-214
    vals = np.array([65535 / 128 - 0.1, 0 + 1.0 / 3.0], np.float32)
    back = oracle.scalar_to_f32_vec(oracle.f32_to_scalar_vec(vals, NB, FR), NB, FR)
    assert (np.abs(back - vals) <= 2.0 ** -(FR + 1)).all()


def test_round_ties_to_even(oracle):                        # fixed 0.3.3 saturating_from_float
    # raw*128 = k + 0.5 exactly -> even neighbour
    vals = np.array([0.5 / 128, 1.5 / 128, 2.5 / 128, 3.5 / 128], np.float32)
    s = oracle.f32_to_scalar_vec(vals, NB, FR)
    assert [sc_int(x) for x in s] == [0, 2, 2, 4]


def test_conversion_saturated(oracle):                      # conversion32.rs:217-230
    mx = np.float32(65535 / 128)
    s = oracle.f32_to_scalar_vec(np.array([mx + 5, -mx - 100], np.float32), NB, FR)
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == np.array([mx, -mx], np.float32)).all()


def test_square_fn(oracle):                                 # conversion32.rs:164-178
    for v in [2.0, 4.0, 2.25, 2.5, 12.5, 112.5, -2.0]:     # 112.5^2 needs the fp32 feature set
        s = oracle.f32_to_scalar_vec(np.array([v], np.float32), 32, FR)[0]
        sq = oracle.square(s.tobytes(), 32, FR)
        assert oracle.scalar_to_f32_vec(np.frombuffer(sq, np.uint8), 32, FR)[0] == np.float32(v * v)
    with pytest.raises(OverflowError):
        oracle.square(oracle.f32_to_scalar_vec(np.array([112.5], np.float32), NB, FR)[0].tobytes(), NB, FR)


def test_clip_bounds(oracle):                               # conversion32.rs:56-64
    assert oracle.clip_bounds(8, 16, 7) == (-127 / 128, 127 / 128)
    assert oracle.clip_bounds(16, 16, 7) == (-(2**15 - 1) / 128, (2**15 - 1) / 128)
    assert oracle.l2_clip_bound(32, 32, 7) == np.float32((2**32 - 1) / 128)
    assert oracle.next_pow2(1) == 1 and oracle.next_pow2(127) == 128 and oracle.next_pow2(1 << 31) == 1 << 31   # range_proof_vec/mod.rs:260-264
    assert oracle.next_pow2(128) == 128 and oracle.next_pow2(129) == 256


# ---- pedersen_ops.rs tests ----------------------------------------------------------------------
def test_commit_no_blinding_dlog_roundtrip(oracle):         # pedersen_ops.rs:191-200, conversion32.rs:233-246
    x = np.array([0.25, 1.25, -1.5, 0.0, 65535 / 128 + 5, -65535 / 128 - 100], np.float32)
    com = oracle.commit_f32(x, None, NB, FR)
    rc, s = oracle.dlog(com, 1 << 16, 16)
    assert rc == 0
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == np.array([0.25, 1.25, -1.5, 0.0, 65535 / 128, -65535 / 128], np.float32)).all()


def test_cancelling_blindings_sum(oracle):                  # pedersen_ops.rs:138-162,250-277
    x, y, z = ([0.25, 1.25, -1.5], [-0.75, 1.25, -2.0], [0.5, 1.25, -3.0])
    b1 = oracle.rnd_scalar_vec(b"\x01" * 32, 3); b2 = oracle.rnd_scalar_vec(b"\x02" * 32, 3)
    b3 = np.stack([np.frombuffer(((-(sc_int(b1[i]) + sc_int(b2[i]))) % L).to_bytes(32, "little"), np.uint8) for i in range(3)])
    cs = np.stack([oracle.commit_f32(np.array(v, np.float32), b, NB, FR) for v, b in [(x, b1), (y, b2), (z, b3)]])
    rc, s = oracle.dlog(oracle.aggregate(cs, 0))
    assert rc == 0 and (oracle.scalar_to_f32_vec(s, NB, FR) == np.array([0.0, 3.75, -6.5], np.float32)).all()


def test_dlog_small_table_giant_steps(oracle):              # bsgs32.rs tests / pedersen_ops.rs:280-296 (fp8: table 2^(4+3))
    x = np.array([0.5, -0.75, 1.9921875, -1.9921875], np.float32)    # raw up to 255 at frac 7
    com = oracle.commit_f32(x, None, 8, 7)
    rc, s = oracle.dlog(com, 1 << 7, 8)
    assert rc == 0 and (oracle.scalar_to_f32_vec(s, 8, 7) == x).all()


def test_accumulator_unity_quirk(oracle):                   # (B,B) start => +1 LSB
    x = np.array([0.25, -0.5], np.float32)
    agg = oracle.aggregate(oracle.commit_f32(x, None, NB, FR).reshape(1, 2, 32), 1)
    rc, s = oracle.dlog(agg)
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == x + np.float32(1 / 128)).all()


# ---- range_proof_vec tests -----------------------------------------------------------------------
def test_rangeproof_roundtrip(oracle):                      # range_proof_vec/mod.rs:267-277
    v = oracle.clip_f32_to_range_vec(np.array([-1.25, 0.5, -65535 / 128], np.float32), 16, NB, FR)
    rc, proofs, commits = oracle.range_prove(v, oracle.rnd_scalar_vec(b"\x05" * 32, 3), 16, 4, NB, FR)
    assert rc == 0 and proofs.shape == (4, 32 * (9 + 2 * 4))
    assert oracle.range_verify(proofs, commits, 16) == 1


def test_rangeproof_par_roundtrip_and_fake(oracle):         # :296-315, :335-366
    rng = np.random.default_rng(7)
    v = oracle.clip_f32_to_range_vec(rng.uniform(-255.9, 255.9, 100).astype(np.float32), 16, NB, FR)
    rc, proofs, commits = oracle.range_prove(v, oracle.rnd_scalar_vec(b"\x06" * 32, 100), 16, 4, NB, FR)
    assert rc == 0 and oracle.range_verify(proofs, commits, 16) == 1
    fake = commits.copy()
    fake[0] = np.frombuffer(oracle.commit_scalars(np.frombuffer((1 << 17).to_bytes(32, "little"), np.uint8), oracle.rnd_scalar_vec(b"\x07" * 32, 1))[0], np.uint8)
    assert oracle.range_verify(proofs, fake, 16) == 0


def test_create_rangeproof_correct_shift_and_clipped(oracle):   # :318-332, :402-417
    x = np.array([0.25, 1.25, -1.5], np.float32)
    rc, _, commits = oracle.range_prove(x, np.zeros((3, 32), np.uint8), 16, 4, NB, FR)
    rc2, s = oracle.dlog(commits)
    assert rc == 0 and rc2 == 0 and (oracle.scalar_to_f32_vec(s, NB, FR) == x).all()
    mn, mx = oracle.clip_bounds(16, NB, FR)
    xc = oracle.clip_f32_to_range_vec(np.array([mn - 2, mx + 3], np.float32), 16, NB, FR)
    rc, _, commits = oracle.range_prove(xc, np.zeros((2, 32), np.uint8), 16, 2, NB, FR)
    rc2, s = oracle.dlog(commits)
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == np.array([mn, mx], np.float32)).all()


def test_rangeproof_out_of_range_and_bad_partition(oracle):    # :26-29 ValueOutOfRangeError; :137-140 panic
    rc, _, _ = oracle.range_prove(np.array([1.5], np.float32), np.zeros((1, 32), np.uint8), 8, 4, NB, FR)
    assert rc == 2
    rc, _, _ = oracle.range_prove(np.zeros(8, np.float32), np.zeros((8, 32), np.uint8), 8, 3, NB, FR)
    assert rc == -99
    rc, _, _ = oracle.range_prove(np.zeros(8, np.float32), np.zeros((8, 32), np.uint8), 12, 4, NB, FR)
    assert rc == -7            # InvalidBitsize


def test_rangeproof_cancelling_blindings(oracle):           # :369-399
    vecs = [[0.25, 1.25, -1.5], [-0.75, 1.25, -2.0], [0.5, 1.25, -3.0]]
    b1 = oracle.rnd_scalar_vec(b"\x08" * 32, 3); b2 = oracle.rnd_scalar_vec(b"\x09" * 32, 3)
    b3 = np.stack([np.frombuffer(((-(sc_int(b1[i]) + sc_int(b2[i]))) % L).to_bytes(32, "little"), np.uint8) for i in range(3)])
    cs = []
    for v, b in zip(vecs, [b1, b2, b3]):
        rc, p, c = oracle.range_prove(np.array(v, np.float32), b, 16, 4, NB, FR)
        assert rc == 0 and oracle.range_verify(p, c, 16) == 1
        cs.append(c)
    rc, s = oracle.dlog(oracle.aggregate(np.stack(cs), 0))
    assert (oracle.scalar_to_f32_vec(s, NB, FR) == np.array([0.0, 3.75, -6.5], np.float32)).all()


def test_rangeproof_malformed_proof_is_error(oracle):       # RangeProof::from_bytes FormatError propagates (:210-215)
    v = np.array([0.5, 0.25], np.float32)
    rc, proofs, commits = oracle.range_prove(v, np.zeros((2, 32), np.uint8), 8, 1, NB, FR)
    bad = proofs.copy(); bad[0, 128:160] = 0xff             # non-canonical t_x
    assert oracle.range_verify(bad, commits, 8) == -1
    bad = proofs.copy(); bad[0, 0:32] = 0                   # identity A -> VerificationError -> Ok(false)
    assert oracle.range_verify(bad, commits, 8) == 0


# ---- l2_range_proof_vec tests --------------------------------------------------------------------
def test_l2_roundtrip_and_value(oracle):                    # l2_range_proof_vec/mod.rs:303-327, :414-431
    v = np.array([0.25, 1.25, -1.5], np.float32)
    rc, proof, commit = oracle.l2_prove(v, np.zeros((3, 32), np.uint8), 16, 16, 7)
    assert rc == 0 and oracle.l2_verify(proof, commit, 16) == 1
    rc, s = oracle.dlog(commit)
    assert oracle.scalar_to_f32_vec(s, 16, 7)[0] == 3.875 * 128


def test_l2_bounds(oracle):                                 # :329-373
    for v in ([7.9], [-7.9]):
        rc, proof, commit = oracle.l2_prove(np.array(v, np.float32), oracle.rnd_scalar_vec(b"\x0a" * 32, 1), 32, 32, 7)
        assert rc == 0 and oracle.l2_verify(proof, commit, 32) == 1
    assert oracle.l2_prove(np.array([8.0], np.float32), np.zeros((1, 32), np.uint8), 16, 32, 7)[0] == 4
    assert oracle.l2_prove(np.array([6.0, 6.0], np.float32), np.zeros((2, 32), np.uint8), 16, 32, 7)[0] == 4


def test_l2_fake_proof(oracle):                             # :376-389
    rc, proof, _ = oracle.l2_prove(np.array([0.5], np.float32), oracle.rnd_scalar_vec(b"\x0b" * 32, 1), 16, 16, 7)
    fake = oracle.commit_scalars(np.frombuffer((1 << 17).to_bytes(32, "little"), np.uint8), oracle.rnd_scalar_vec(b"\x0c" * 32, 1))[0]
    assert oracle.l2_verify(proof, fake, 16) == 0


def test_l2_sum_commit_equals_sum_of_square_commits(oracle):    # :539-561
    rng = np.random.default_rng(9)
    v = (rng.integers(-24, 25, 40) / 128).astype(np.float32)
    r1 = oracle.rnd_scalar_vec(b"\x0d" * 32, 40); r2 = oracle.rnd_scalar_vec(b"\x0e" * 32, 40)
    rc, _, l2c = oracle.l2_prove(v, r2, 32, 32, 7)
    c1 = oracle.commit_f32(v, r1, 32, 7)
    rc2, sp, sc = oracle.square_prove(v, c1, r1, r2, 32, 7)
    assert rc == 0 and rc2 == 0
    assert oracle.aggregate(sc[:, 32:].reshape(40, 1, 32), 0)[0].tobytes() == l2c.tobytes()


# ---- square_proof tests ---------------------------------------------------------------------------
def test_square_proof_roundtrip_and_tamper(oracle):         # square_proof/mod.rs:214-252, square_proof_vec/mod.rs:166-193
    rng = np.random.default_rng(10)
    v = rng.uniform(-1, 1, 25).astype(np.float32)
    r1 = oracle.rnd_scalar_vec(b"\x0f" * 32, 25); r2 = oracle.rnd_scalar_vec(b"\x10" * 32, 25)
    c1 = oracle.commit_f32(v, r1, 32, 7)
    rc, proofs, commits = oracle.square_prove(v, c1, r1, r2, 32, 7)
    assert rc == 0 and proofs.shape == (25, 160) and commits.shape == (25, 64)
    assert (commits[:, :32] == c1).all()
    assert oracle.square_verify(proofs, commits) == 1
    bad = commits.copy(); bad[3, 32:] = commits[4, 32:]     # wrong c_sq
    assert oracle.square_verify(proofs, bad) == 0
    bad = proofs.copy(); bad[0, 64:96] = 0xff               # non-canonical scalar -> FormatError
    assert oracle.square_verify(bad, commits) == -1
