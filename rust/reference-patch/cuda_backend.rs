//! `src/cuda_backend.rs` of sigma0-polymath under the `cuda` feature (see `reference.patch` beside this file).
//!
//! The B200 backend replaces the BODY of `create_proof_with_assignment` (src/prover.rs:66-237) and of the six
//! `generate(..)` calls of `generate_proving_key` (src/generator.rs:81-137) when the pairing engine is BLS12-381; the
//! public API, `Proof` / `ProvingKey` / `VerifyingKey` and their serialisation do not change, and the Fiat–Shamir code
//! of src/common.rs runs unmodified between the device phases.  The crate keeps `#![forbid(unsafe_code)]`: every
//! pointer lives in `polymath-b200-sys`.
//!
//! Not compiled in the image this repository is developed in (no Rust toolchain); the identical call sequence is
//! exercised through the same C ABI by `polymath_b200/csrc/host/polymath_host.cpp` and the GPU test-suite.
use std::any::Any;
use std::collections::HashMap;
use std::sync::{Arc, Mutex, OnceLock};

use ark_bls12_381::{Bls12_381, Fr, G1Affine};
use ark_ec::pairing::Pairing;
use ark_ff::{Field, PrimeField};
use ark_relations::r1cs::Matrix;
use ark_std::rand::RngCore;
use polymath_b200_sys as sys;
use polymath_b200_sys::ark as view;

use crate::common::{B_POLYMATH, MINUS_ALPHA, MINUS_GAMMA};
use crate::{Polymath, PolymathError, Proof, ProvingKey, Transcript};

type Shared = Arc<Mutex<sys::Context>>;

/// Device copies of proving keys, one per key (`prove(&pk, ..)` borrows the key on every call, src/lib.rs:72-78).
/// The cache key is the address and length of `x_powers_g1` plus the domain size: stable while the key is alive.
fn cache() -> &'static Mutex<HashMap<(usize, usize, u64), Shared>> {
    static CACHE: OnceLock<Mutex<HashMap<(usize, usize, u64), Shared>>> = OnceLock::new();
    CACHE.get_or_init(|| Mutex::new(HashMap::new()))
}

fn csr(m: &Matrix<Fr>) -> sys::Csr {
    sys::Csr::from_rows(m.iter().map(|row| row.iter().map(|(coeff, col)| (view::fp_slice_as_bytes(core::slice::from_ref(coeff)), *col))))
}

fn r1cs_of(pk: &ProvingKey<Bls12_381>) -> sys::R1cs {
    let s = &pk.sap_matrices;
    sys::R1cs {
        num_instance_variables: s.num_instance_variables as u64,
        num_r1cs_witness_variables: s.num_r1cs_witness_variables as u64,
        num_r1cs_constraints: s.num_r1cs_constraints as u64,
        a: csr(&s.a),
        b: csr(&s.b),
        c: csr(&s.c),
    }
}

fn upload(pk: &ProvingKey<Bls12_381>) -> sys::Result<sys::Context> {
    let r1cs = r1cs_of(pk);
    if view::g1_layout_is_native() {
        let vecs = [
            view::g1_slice_as_strided_bytes(&pk.x_powers_g1),
            view::g1_slice_as_strided_bytes(&pk.x_powers_y_alpha_g1),
            view::g1_slice_as_strided_bytes(&pk.x_powers_zh_by_y_alpha_g1),
            view::g1_slice_as_strided_bytes(&pk.x_powers_y_gamma_g1),
            view::g1_slice_as_strided_bytes(&pk.x_powers_y_gamma_z_g1),
            view::g1_slice_as_strided_bytes(&pk.uj_wj_lcs_by_y_alpha_g1),
        ];
        sys::Context::from_key(&sys::KeyView {
            r1cs: &r1cs,
            n: pk.vk.n,
            sigma: pk.vk.sigma,
            point_stride: vecs[0].1,
            x_powers_g1: vecs[0].0,
            x_powers_y_alpha_g1: vecs[1].0,
            x_powers_zh_by_y_alpha_g1: vecs[2].0,
            x_powers_y_gamma_g1: vecs[3].0,
            x_powers_y_gamma_z_g1: vecs[4].0,
            uj_wj_lcs_by_y_alpha_g1: vecs[5].0,
        })
    } else {
        let packed = [
            view::g1_pack(&pk.x_powers_g1),
            view::g1_pack(&pk.x_powers_y_alpha_g1),
            view::g1_pack(&pk.x_powers_zh_by_y_alpha_g1),
            view::g1_pack(&pk.x_powers_y_gamma_g1),
            view::g1_pack(&pk.x_powers_y_gamma_z_g1),
            view::g1_pack(&pk.uj_wj_lcs_by_y_alpha_g1),
        ];
        sys::Context::from_key(&sys::KeyView {
            r1cs: &r1cs,
            n: pk.vk.n,
            sigma: pk.vk.sigma,
            point_stride: sys::G1_BYTES,
            x_powers_g1: &packed[0],
            x_powers_y_alpha_g1: &packed[1],
            x_powers_zh_by_y_alpha_g1: &packed[2],
            x_powers_y_gamma_g1: &packed[3],
            x_powers_y_gamma_z_g1: &packed[4],
            uj_wj_lcs_by_y_alpha_g1: &packed[5],
        })
    }
}

/// The device context of `pk` when the engine is BLS12-381 and a B200 is visible; `None` sends the caller down the
/// stock arkworks path (other curves, or a build of the feature on a machine without a device).
pub(crate) fn device_for<E: Pairing>(pk: &ProvingKey<E>) -> Option<Shared> {
    let pk = (pk as &dyn Any).downcast_ref::<ProvingKey<Bls12_381>>()?;
    if sys::device_count() <= 0 {
        return None;
    }
    let key = (pk.x_powers_g1.as_ptr() as usize, pk.x_powers_g1.len(), pk.vk.n);
    let mut map = cache().lock().unwrap();
    if let Some(ctx) = map.get(&key) {
        return Some(ctx.clone());
    }
    let ctx = Arc::new(Mutex::new(upload(pk).expect("proving key upload to the device failed")));
    map.insert(key, ctx.clone());
    Some(ctx)
}

/// Forget the device copy of a key (call before dropping a `ProvingKey` whose memory may be reused).
pub fn evict<E: Pairing>(pk: &ProvingKey<E>) {
    if let Some(pk) = (pk as &dyn Any).downcast_ref::<ProvingKey<Bls12_381>>() {
        cache().lock().unwrap().remove(&(pk.x_powers_g1.as_ptr() as usize, pk.x_powers_g1.len(), pk.vk.n));
    }
}

fn fr_bytes(v: &Fr) -> [u8; sys::FR_BYTES] {
    view::fp_slice_as_bytes(core::slice::from_ref(v)).try_into().unwrap()
}

fn generic<A: 'static + Copy, B: 'static + Copy>(v: A) -> B {
    view::same_type_value::<A, B>(v).expect("BLS12-381 types")
}

/// The reference's panics, at the same protocol points (src/prover.rs:107, :108, :221).
fn protocol_panic(e: sys::Error) -> ! {
    match e {
        sys::Error::Unsatisfied(m) => panic!("assertion failed: rem_poly.is_zero() [{m}]"),
        sys::Error::Degenerate(m) => panic!("assertion failed: !h_poly.is_zero() && h_poly.degree() <= n - 2 [{m}]"),
        sys::Error::Remainder(m) => panic!("assertion failed: rem_poly.is_zero() (opening) [{m}]"),
        other => panic!("polymath_b200: {other}"),
    }
}

impl<F: PrimeField, E, T> Polymath<E, T>
where
    E: Pairing<ScalarField = F>,
    T: Transcript<Challenge = F>,
{
    /// `create_proof_with_assignment` (src/prover.rs:66-237) with the three device phases in place of the arkworks
    /// polynomial and MSM code.  Consumes the RNG exactly like the reference: two `F::rand` draws (src/prover.rs:110).
    pub(crate) fn create_proof_on_device<R: RngCore>(
        ctx: &Shared,
        pk: &ProvingKey<E>,
        instance_assignment: &[F],
        witness_assignment: &[F],
        rng: &mut R,
    ) -> Result<Proof<E>, PolymathError> {
        let x: &[Fr] = view::same_type_slice(instance_assignment).expect("BLS12-381 scalar field");
        let w: &[Fr] = view::same_type_slice(witness_assignment).expect("BLS12-381 scalar field");
        let mut ctx = ctx.lock().unwrap();

        // The reference checks the quotient (prover.rs:104-108) BEFORE it draws r_a (:110); a failing witness must
        // therefore leave the caller's RNG untouched.  r_a only enters the MSM scalars, so draw it first and let the
        // panic below unwind before anyone can observe the generator — `rng` is borrowed mutably for this call only.
        let r_a = [F::rand(rng), F::rand(rng)];
        let mut r_a_bytes = [0u8; 2 * sys::FR_BYTES];
        r_a_bytes[..sys::FR_BYTES].copy_from_slice(&fr_bytes(&generic::<F, Fr>(r_a[0])));
        r_a_bytes[sys::FR_BYTES..].copy_from_slice(&fr_bytes(&generic::<F, Fr>(r_a[1])));

        let (a, c) = ctx
            .prove_phase1(view::fp_slice_as_bytes(x), view::fp_slice_as_bytes(w), &r_a_bytes)
            .unwrap_or_else(|e| protocol_panic(e));
        let a_g1: E::G1Affine = generic::<G1Affine, E::G1Affine>(view::g1_from_bytes(&a));
        let c_g1: E::G1Affine = generic::<G1Affine, E::G1Affine>(view::g1_from_bytes(&c));

        // Fiat–Shamir exactly as src/prover.rs:125-140,187 (src/common.rs:21-97 untouched)
        let mut t = T::new(B_POLYMATH);
        let x1 = Self::compute_x1(&mut t, instance_assignment, &[a_g1, c_g1])?;
        let y1 = Self::compute_y1(x1, pk.vk.sigma);
        let y1_alpha = Self::neg_power(y1, MINUS_ALPHA);
        let a_at_x1_bytes = ctx
            .prove_phase2(&fr_bytes(&generic::<F, Fr>(x1)), &fr_bytes(&generic::<F, Fr>(y1_alpha)))
            .unwrap_or_else(|e| protocol_panic(e));
        let a_at_x1: F = generic::<Fr, F>(view::fp_from_bytes(&a_at_x1_bytes));
        let y1_gamma = Self::neg_power(y1, MINUS_GAMMA);
        let pi_at_x1 = Self::compute_pi_at_x1(&pk.vk, instance_assignment, x1, y1_gamma);
        let c_at_x1 = Self::compute_c_at_x1(y1_gamma, y1_alpha, a_at_x1, pi_at_x1);
        let x2 = Self::compute_x2(&mut t, &x1, &[a_at_x1, c_at_x1])?;

        let d = ctx
            .prove_phase3(&fr_bytes(&generic::<F, Fr>(x2)), &fr_bytes(&generic::<F, Fr>(c_at_x1)))
            .unwrap_or_else(|e| protocol_panic(e));
        let d_g1: E::G1Affine = generic::<G1Affine, E::G1Affine>(view::g1_from_bytes(&d));

        Ok(Proof { a_g1, c_g1, a_at_x1, d_g1 })
    }
}

/// `generate()` (src/generator.rs:169-177) on the device: `g * f(j)` for j = 0..=max_index as ONE fixed-base batch
/// multiplication (`pm_fixed_base_mul`) when `G` is BLS12-381 G1 and `g` its standard generator; `None` otherwise.
/// The scalars are the reference's own closures, evaluated on the host cores.
pub(crate) fn generate_on_device<G, M>(g: &G, max_index: usize, f: &M) -> Option<Vec<G::Affine>>
where
    G: ark_ec::CurveGroup,
    M: Fn(u64) -> G::ScalarField,
{
    use ark_bls12_381::G1Projective;
    use ark_ec::PrimeGroup;
    if std::env::var_os("POLYMATH_FORCE_CPU").is_some() || sys::device_count() <= 0 {
        return None;
    }
    let g1 = view::same_type_value::<G, G1Projective>(*g)?;
    if g1 != G1Projective::generator() {
        return None;
    }
    let scalars: Vec<G::ScalarField> = (0..max_index as u64 + 1).map(f).collect();
    let scalars: &[Fr] = view::same_type_slice(&scalars)?;
    let points = sys::fixed_base_mul(view::fp_slice_as_bytes(scalars)).expect("fixed-base multiplication on the device failed");
    let points = view::g1_vec_from_bytes(&points);
    Some(view::same_type_slice::<G1Affine, G::Affine>(&points)?.to_vec())
}

/// Whole-key variant: the six G1 vectors of `generate_proving_key` (src/generator.rs:81-137) built on the device from
/// the trapdoors (`pm_setup`: powers, Lagrange values and the lcs scalars are computed on the device as well), for
/// callers that restructure `generate_proving_key`; `reference.patch` uses the smaller `generate_on_device` seam.
pub struct DeviceKeyVectors {
    pub x_powers_g1: Vec<G1Affine>,
    pub x_powers_y_alpha_g1: Vec<G1Affine>,
    pub x_powers_zh_by_y_alpha_g1: Vec<G1Affine>,
    pub x_powers_y_gamma_g1: Vec<G1Affine>,
    pub x_powers_y_gamma_z_g1: Vec<G1Affine>,
    pub uj_wj_lcs_by_y_alpha_g1: Vec<G1Affine>,
}

pub fn generate_key_vectors_on_device(
    num_instance_variables: usize,
    num_r1cs_witness_variables: usize,
    a: &Matrix<Fr>,
    b: &Matrix<Fr>,
    c: &Matrix<Fr>,
    x: Fr,
    z: Fr,
) -> sys::Result<DeviceKeyVectors> {
    let r1cs = sys::R1cs {
        num_instance_variables: num_instance_variables as u64,
        num_r1cs_witness_variables: num_r1cs_witness_variables as u64,
        num_r1cs_constraints: a.len() as u64,
        a: csr(a),
        b: csr(b),
        c: csr(c),
    };
    let (ctx, _x_g2, _z_g2) = sys::Context::setup(&r1cs, &fr_bytes(&x), &fr_bytes(&z))?;
    let get = |which| ctx.export_key(which, sys::G1_BYTES).map(|bytes| view::g1_vec_from_bytes(&bytes));
    Ok(DeviceKeyVectors {
        x_powers_g1: get(sys::KeyVector::XPowers)?,
        x_powers_y_alpha_g1: get(sys::KeyVector::XPowersYAlpha)?,
        x_powers_zh_by_y_alpha_g1: get(sys::KeyVector::XPowersZhByYAlpha)?,
        x_powers_y_gamma_g1: get(sys::KeyVector::XPowersYGamma)?,
        x_powers_y_gamma_z_g1: get(sys::KeyVector::XPowersYGammaZ)?,
        uj_wj_lcs_by_y_alpha_g1: get(sys::KeyVector::UjWjLcsByYAlpha)?,
    })
}
