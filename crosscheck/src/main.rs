//! Dumps, from the UNMODIFIED reference and the arkworks it is built against, every value the oracle of
//! polymath_b200 claims to reproduce ("Next round" item 2):
//!
//!   rng.json       first `next_u64`s and `Fr::rand`s of `StdRng::seed_from_u64(s)` (benches/bench.rs:65-68)
//!   transcript.json Merlin / Keccak256 / Blake3 challenges x1, x2 for a fixed statement (src/common.rs:21-37)
//!   kernels.json   `fft`, `ifft`, `msm_unchecked` on the inputs of tests/golden/kernels.json (src/prover.rs:241,319,383)
//!   dummy_seed0.json, mimc322_seed1.json, mimc8_seed7.json
//!                  vk and proof bytes of the flows of tests/dummy.rs:37-74 and tests/mimc.rs:146-216 with the seeds
//!                  and draw order of tests/golden/make_golden.py
//!
//! usage: polymath-crosscheck <output dir> <tests/golden/kernels.json>
use std::fmt::Write as _;
use std::fs;
use std::path::Path;

use ark_bls12_381::{Bls12_381, Fr, G1Affine, G1Projective};
use ark_crypto_primitives::snark::{CircuitSpecificSetupSNARK, SNARK};
use ark_ec::{CurveGroup, PrimeGroup, VariableBaseMSM};
use ark_ff::{Field, PrimeField, UniformRand};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_relations::{
    lc,
    r1cs::{ConstraintSynthesizer, ConstraintSystemRef, SynthesisError, Variable},
};
use ark_serialize::CanonicalSerialize;
use ark_std::rand::{rngs::StdRng, RngCore, SeedableRng};
use sigma0_polymath::{
    blake3::Blake3Transcript, keccak256::Keccak256Transcript, merlin::MerlinFieldTranscript, Polymath, Transcript,
};

type Pm = Polymath<Bls12_381, MerlinFieldTranscript<Fr>>;

// ---- circuits: the constraint systems of tests/dummy.rs:20-35 and tests/mimc.rs:66-143 (round count as a parameter) ----

struct Product {
    a: Option<Fr>,
    b: Option<Fr>,
}

impl ConstraintSynthesizer<Fr> for Product {
    fn generate_constraints(self, cs: ConstraintSystemRef<Fr>) -> Result<(), SynthesisError> {
        let a = cs.new_witness_variable(|| self.a.ok_or(SynthesisError::AssignmentMissing))?;
        let b = cs.new_witness_variable(|| self.b.ok_or(SynthesisError::AssignmentMissing))?;
        let prod = self.a.zip(self.b).map(|(a, b)| a * b);
        let c = cs.new_input_variable(|| prod.ok_or(SynthesisError::AssignmentMissing))?;
        cs.enforce_constraint(lc!() + a, lc!() + b, lc!() + c)
    }
}

fn mimc_hash(mut xl: Fr, mut xr: Fr, constants: &[Fr]) -> Fr {
    for c in constants {
        let t = xl + c;
        let next = t.square() * t + xr;
        xr = xl;
        xl = next;
    }
    xl
}

struct Mimc<'a> {
    xl: Option<Fr>,
    xr: Option<Fr>,
    constants: &'a [Fr],
}

impl<'a> ConstraintSynthesizer<Fr> for Mimc<'a> {
    fn generate_constraints(self, cs: ConstraintSystemRef<Fr>) -> Result<(), SynthesisError> {
        let (mut xl_v, mut xr_v) = (self.xl, self.xr);
        let mut xl = cs.new_witness_variable(|| xl_v.ok_or(SynthesisError::AssignmentMissing))?;
        let mut xr = cs.new_witness_variable(|| xr_v.ok_or(SynthesisError::AssignmentMissing))?;
        let rounds = self.constants.len();
        for (i, k) in self.constants.iter().enumerate() {
            // t = (xl + k)^2
            let t_v = xl_v.map(|x| (x + k).square());
            let t = cs.new_witness_variable(|| t_v.ok_or(SynthesisError::AssignmentMissing))?;
            cs.enforce_constraint(lc!() + xl + (*k, Variable::One), lc!() + xl + (*k, Variable::One), lc!() + t)?;
            // next - xr = t * (xl + k); the last `next` is the public image
            let next_v = xl_v.map(|x| (x + k) * t_v.unwrap() + xr_v.unwrap());
            let next = if i + 1 == rounds {
                cs.new_input_variable(|| next_v.ok_or(SynthesisError::AssignmentMissing))?
            } else {
                cs.new_witness_variable(|| next_v.ok_or(SynthesisError::AssignmentMissing))?
            };
            cs.enforce_constraint(lc!() + t, lc!() + xl + (*k, Variable::One), lc!() + next - xr)?;
            xr = xl;
            xr_v = xl_v;
            xl = next;
            xl_v = next_v;
        }
        Ok(())
    }
}

// ---- small JSON / formatting helpers (no serde: the dependency set stays the reference's own) ----

fn hex(bytes: &[u8]) -> String {
    let mut s = String::with_capacity(bytes.len() * 2);
    for b in bytes {
        write!(s, "{b:02x}").unwrap();
    }
    s
}

fn ser<T: CanonicalSerialize>(v: &T) -> String {
    let mut buf = Vec::new();
    v.serialize_compressed(&mut buf).unwrap();
    hex(&buf)
}

fn dec(v: &Fr) -> String {
    v.into_bigint().to_string()
}

fn dec_list(vs: &[Fr]) -> String {
    let items: Vec<String> = vs.iter().map(|v| format!("\"{}\"", dec(v))).collect();
    format!("[{}]", items.join(", "))
}

/// The array of decimal strings stored under `key` in a flat JSON object.
fn read_list(json: &str, key: &str) -> Vec<Fr> {
    let start = json.find(&format!("\"{key}\"")).unwrap_or_else(|| panic!("key {key} missing"));
    let open = start + json[start..].find('[').unwrap();
    let close = open + json[open..].find(']').unwrap();
    json[open + 1..close]
        .split(',')
        .map(|s| s.trim().trim_matches('"'))
        .filter(|s| !s.is_empty())
        .map(|s| s.parse::<Fr>().unwrap_or_else(|_| panic!("bad field element {s}")))
        .collect()
}

fn write(dir: &Path, name: &str, body: String) {
    fs::write(dir.join(name), body).unwrap();
    println!("wrote {name}");
}

// ---- the dumps ----

fn dump_rng(dir: &Path) {
    let mut entries = Vec::new();
    for seed in [0u64, 1, 7, 2024] {
        let mut a = StdRng::seed_from_u64(seed);
        let words: Vec<String> = (0..8).map(|_| format!("\"{}\"", a.next_u64())).collect();
        let mut b = StdRng::seed_from_u64(seed);
        let frs: Vec<Fr> = (0..6).map(|_| Fr::rand(&mut b)).collect();
        entries.push(format!(
            "  {{\"seed\": {seed}, \"next_u64\": [{}], \"fr_rand\": {}}}",
            words.join(", "),
            dec_list(&frs)
        ));
    }
    write(dir, "rng.json", format!("[\n{}\n]\n", entries.join(",\n")));
}

/// x1, x2 exactly as `compute_x1` / `compute_x2` (src/common.rs:21-37) build them, through the public `Transcript` trait.
fn challenges<T: Transcript<Challenge = Fr>>(public: &[Fr], commitments: &[G1Affine], values: &[Fr]) -> (Fr, Fr) {
    let bytes = |f: &dyn Fn(&mut Vec<u8>)| {
        let mut buf = Vec::new();
        f(&mut buf);
        buf
    };
    let mut t = T::new(b"polymath");
    t.append_message(b"public_inputs", bytes(&|b| public.serialize_compressed(b).unwrap()));
    t.append_message(b"commitments", bytes(&|b| commitments.serialize_compressed(b).unwrap()));
    let x1 = t.challenge(b"x1");
    t.append_message(b"x1", bytes(&|b| x1.serialize_compressed(b).unwrap()));
    t.append_message(b"values", bytes(&|b| values.serialize_compressed(b).unwrap()));
    let x2 = t.challenge(b"x2");
    (x1, x2)
}

fn dump_transcripts(dir: &Path) {
    let public = [Fr::from(1u64), Fr::from(42u64)];
    let g = G1Projective::generator();
    let commitments = [(g * Fr::from(3u64)).into_affine(), (g * Fr::from(5u64)).into_affine()];
    let values = [Fr::from(7u64), Fr::from(11u64)];
    let m = challenges::<MerlinFieldTranscript<Fr>>(&public, &commitments, &values);
    let k = challenges::<Keccak256Transcript<Fr>>(&public, &commitments, &values);
    let b = challenges::<Blake3Transcript<Fr>>(&public, &commitments, &values);
    let row = |name: &str, v: (Fr, Fr)| format!("  \"{name}\": {{\"x1\": \"{}\", \"x2\": \"{}\"}}", dec(&v.0), dec(&v.1));
    write(
        dir,
        "transcript.json",
        format!(
            "{{\n  \"public_inputs\": [\"1\", \"42\"], \"commitment_scalars\": [\"3\", \"5\"], \"values\": [\"7\", \"11\"],\n{},\n{},\n{}\n}}\n",
            row("merlin", m),
            row("keccak256", k),
            row("blake3", b)
        ),
    );
}

fn dump_kernels(dir: &Path, fixture: &Path) {
    let json = fs::read_to_string(fixture).expect("tests/golden/kernels.json");
    let vals = read_list(&json, "ntt_in");
    let domain = Radix2EvaluationDomain::<Fr>::new(vals.len()).unwrap();
    let fwd = domain.fft(&vals);
    let mut inv = vals.clone();
    domain.ifft_in_place(&mut inv);
    let base_scalars = read_list(&json, "base_scalars");
    let scalars = read_list(&json, "scalars");
    let g = G1Projective::generator();
    let bases: Vec<G1Affine> = base_scalars.iter().map(|s| (g * s).into_affine()).collect();
    let msm = G1Projective::msm_unchecked(&bases, &scalars).into_affine();
    let (x, y) = (msm.x, msm.y);
    write(
        dir,
        "kernels.json",
        format!(
            "{{\n \"ntt_fwd\": {},\n \"ntt_inv\": {},\n \"msm\": [\"{}\", \"{}\"],\n \"msm_compressed\": \"{}\"\n}}\n",
            dec_list(&fwd),
            dec_list(&inv),
            x.into_bigint(),
            y.into_bigint(),
            ser(&msm)
        ),
    );
}

/// tests/dummy.rs:37-74 with an explicit seed; draw order as tests/golden/make_golden.py::dummy.
fn dump_dummy(dir: &Path, seed: u64) {
    let mut rng = StdRng::seed_from_u64(seed);
    let (pk, vk) = Pm::setup(Product { a: None, b: None }, &mut rng).unwrap();
    let (a, b) = (Fr::rand(&mut rng), Fr::rand(&mut rng));
    let proof = Pm::prove(&pk, Product { a: Some(a), b: Some(b) }, &mut rng).unwrap();
    assert!(Pm::verify(&vk, &[a * b], &proof).unwrap());
    write(
        dir,
        &format!("dummy_seed{seed}.json"),
        format!(
            "{{\"seed\": {seed}, \"a\": \"{}\", \"b\": \"{}\", \"public_input\": \"{}\",\n \"vk_hex\": \"{}\",\n \"proof_hex\": \"{}\"}}\n",
            dec(&a),
            dec(&b),
            dec(&(a * b)),
            ser(&vk),
            ser(&proof)
        ),
    );
}

/// tests/mimc.rs:146-216 with an explicit seed and round count; draw order as tests/golden/make_golden.py::mimc
/// (`rng.gen()` at tests/mimc.rs:156,196 is `Fr::rand`: ark-ff's `Standard` distribution delegates to `UniformRand`).
fn dump_mimc(dir: &Path, seed: u64, rounds: usize) {
    let mut rng = StdRng::seed_from_u64(seed);
    let constants: Vec<Fr> = (0..rounds).map(|_| Fr::rand(&mut rng)).collect();
    let (pk, vk) = Pm::setup(Mimc { xl: None, xr: None, constants: &constants }, &mut rng).unwrap();
    let (xl, xr) = (Fr::rand(&mut rng), Fr::rand(&mut rng));
    let image = mimc_hash(xl, xr, &constants);
    let proof = Pm::prove(&pk, Mimc { xl: Some(xl), xr: Some(xr), constants: &constants }, &mut rng).unwrap();
    assert!(Pm::verify(&vk, &[image], &proof).unwrap());
    write(
        dir,
        &format!("mimc{rounds}_seed{seed}.json"),
        format!(
            "{{\"seed\": {seed}, \"rounds\": {rounds}, \"xl\": \"{}\", \"xr\": \"{}\", \"image\": \"{}\", \"n\": {},\n \"vk_hex\": \"{}\",\n \"proof_hex\": \"{}\"}}\n",
            dec(&xl),
            dec(&xr),
            dec(&image),
            vk.n,
            ser(&vk),
            ser(&proof)
        ),
    );
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    assert!(args.len() == 3, "usage: polymath-crosscheck <output dir> <tests/golden/kernels.json>");
    let dir = Path::new(&args[1]);
    fs::create_dir_all(dir).unwrap();
    dump_rng(dir);
    dump_transcripts(dir);
    dump_kernels(dir, Path::new(&args[2]));
    dump_dummy(dir, 0);
    dump_mimc(dir, 1, 322);
    dump_mimc(dir, 7, 8);
}
