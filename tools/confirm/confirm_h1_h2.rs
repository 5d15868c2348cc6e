// confirm_h1_h2.rs -- one-shot confirmation of the two facts about un-vendored crates that this repository's bit-exactness
// against the real qope/SIPP rests on.  It cannot be run in the build image
// (no Rust toolchain, no network); a maintainer with the reference checked out runs it ONCE:
//
//     cp tools/confirm/confirm_h1_h2.rs <SIPP checkout>/src/bin/confirm_h1_h2.rs
//     cargo run --release --bin confirm_h1_h2 > got.txt        # add [[bin]] name = "confirm_h1_h2" to Cargo.toml if needed
//     diff <(grep -v '^#' tools/confirm/expected_h1_h2.txt | grep -E 'pairing_exact_H1|myfq12_coeffs_H2') got.txt
//
// Empty diff: H1 (exact final exponent) and H2 (w-power-basis MyFq12) hold and every digest under tests/golden/ is the real
// prover's.  If `pairing` equals the `pairing_ark_notH1` line instead, set SIPP_OPT_FE_NORMALISATION = 1; if the coefficients equal
// the `myfq12_coeffs_notH2` lines, set SIPP_OPT_FQ12_ORDER = 1 -- both switches exist in the library, the oracle and the tests.
use ark_bn254::{Fq12, G1Affine, G2Affine};
use ark_ec::AffineRepr;
use ark_ff::{BigInteger, PrimeField};
use ark_serialize::CanonicalSerialize;
use plonky2_bn254::fields::native::MyFq12;
use plonky2_bn254_pairing::pairing::pairing;

fn main() {
    let e: Fq12 = pairing(G1Affine::generator(), G2Affine::generator());
    let mut bytes = Vec::new();
    e.serialize_uncompressed(&mut bytes).unwrap();
    println!("pairing_exact_H1      = {}", hex::encode(&bytes));
    let my: MyFq12 = e.into();
    for (i, c) in my.coeffs.iter().enumerate() {
        println!("myfq12_coeffs_H2[{:2}]   = {}", i, hex::encode(c.into_bigint().to_bytes_le()));
    }
}
