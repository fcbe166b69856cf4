//! cargo run --release -- ../../tests/golden/harness_vectors.json
//!
//! For every vector of the file (written by tools/gen_harness_vectors.py from the oracle; the GPU library produces the same bytes, tests/test_golden.py)
//! this program runs the REAL third-party crates the reference links -- bulletproofs 4.0.0 `RangeProof::prove_multiple_with_rng`, merlin 3.0.0,
//! curve25519-dalek-ng 4.1.1 -- with the nonce stream injected as `ChaCha20Rng::from_seed(key)` and compares proof and commitment bytes.
//!   key   = SHA3-256(seed || u32le(domain) || u64le(chunk))          (engine.cuh derive_key; domain 1 = range proofs, 4 = L2 sum proof)
//!   draws = per party a_blinding, s_blinding, s_L[0..n], s_R[0..n]; then per party t1_blinding, t2_blinding; one 64-byte block each
//!           (`Scalar::random(rng)`), exactly the order of prove_multiple_with_rng's party / dealer state machine.
//! Exit code 0 = every vector matches = the restatement the whole test suite is anchored on IS the reference's arithmetic.
use bulletproofs::{BulletproofGens, PedersenGens, RangeProof};
use curve25519_dalek_ng::scalar::Scalar;
use merlin::Transcript;
use rand_chacha::ChaCha20Rng;
use rand_core::SeedableRng;
use serde::Deserialize;
use sha3::{Digest, Sha3_256};

#[derive(Deserialize)]
struct Chunk { chunk: u64, values: Vec<u64>, blindings: Vec<String>, proof: String, commitments: Vec<String> }
#[derive(Deserialize)]
struct Vector { name: String, label: String, domain: u32, seed: String, n: usize, chunks: Vec<Chunk> }
#[derive(Deserialize)]
struct File { format: String, vectors: Vec<Vector>, generators: Vec<(String, u32, usize, String)>, pedersen: (String, String) }

fn key(seed: &[u8], domain: u32, chunk: u64) -> [u8; 32] {
    let mut h = Sha3_256::new();
    h.update(seed); h.update(&domain.to_le_bytes()); h.update(&chunk.to_le_bytes());
    let mut k = [0u8; 32]; k.copy_from_slice(&h.finalize()); k
}
fn scalar(hexs: &str) -> Scalar { let mut b = [0u8; 32]; b.copy_from_slice(&hex::decode(hexs).unwrap()); Scalar::from_canonical_bytes(b).expect("canonical scalar") }

fn main() {
    let path = std::env::args().nth(1).expect("usage: rofl_b200_harness harness_vectors.json");
    let file: File = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    assert_eq!(file.format, "rofl_b200 harness vectors v1");
    let pc = PedersenGens::default();
    let mut bad = 0;
    // generators and Pedersen bases ("unverified against upstream" until this program has run once)
    if hex::encode(pc.B.compress().as_bytes()) != file.pedersen.0 || hex::encode(pc.B_blinding.compress().as_bytes()) != file.pedersen.1 { println!("MISMATCH pedersen bases"); bad += 1; }
    for (which, party, i, want) in &file.generators {
        let g = BulletproofGens::new(*i + 1, *party as usize + 1);
        let p = if which == "G" { g.share(*party as usize).G(*i + 1).last().unwrap().compress() } else { g.share(*party as usize).H(*i + 1).last().unwrap().compress() };
        if hex::encode(p.as_bytes()) != *want { println!("MISMATCH generator {}_{}[{}]", which, party, i); bad += 1; }
    }
    for v in &file.vectors {
        let seed = hex::decode(&v.seed).unwrap();
        for c in &v.chunks {
            let m = c.values.len();
            let blindings: Vec<Scalar> = c.blindings.iter().map(|b| scalar(b)).collect();
            let mut rng = ChaCha20Rng::from_seed(key(&seed, v.domain, c.chunk));
            let mut t = Transcript::new(v.label.as_bytes());
            let (proof, commits) = RangeProof::prove_multiple_with_rng(&BulletproofGens::new(v.n, m), &pc, &mut t, &c.values, &blindings, v.n, &mut rng).unwrap();
            let ok_p = hex::encode(proof.to_bytes()) == c.proof;
            let ok_c = commits.iter().zip(&c.commitments).all(|(a, b)| hex::encode(a.as_bytes()) == *b);
            // and the reference verifier accepts the expected bytes
            let mut tv = Transcript::new(v.label.as_bytes());
            let ok_v = RangeProof::from_bytes(&hex::decode(&c.proof).unwrap()).unwrap().verify_multiple(&BulletproofGens::new(v.n, m), &pc, &mut tv, &commits, v.n).is_ok();
            println!("{} chunk {}: proof {} commitments {} reference-verifier {}", v.name, c.chunk, if ok_p { "ok" } else { "MISMATCH" }, if ok_c { "ok" } else { "MISMATCH" }, if ok_v { "accepts" } else { "REJECTS" });
            if !(ok_p && ok_c && ok_v) { bad += 1; }
        }
    }
    if bad > 0 { println!("{} mismatches", bad); std::process::exit(1); }
    println!("all vectors match: oracle == rofl_crypto's third-party arithmetic, byte for byte");
}
