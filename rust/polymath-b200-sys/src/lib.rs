//! Bindings of `libpolymath_b200.so` — the B200 (sm_100a) prover backend for `sigma0-polymath`.
//!
//! `ffi` is generated from `include/polymath_b200.h` (tools/gen_rust_sys.py) and binds every entry point; this file is
//! the small safe layer the patched reference calls (`rust/reference-patch/cuda_backend.rs`): byte-slice in, byte-array
//! out, status codes mapped to [`Error`].  All field elements cross the boundary as the little-endian limbs of their
//! Montgomery form — `ark_ff::Fp<MontBackend<_, N>, N>` in memory — so the caller passes arkworks values without
//! conversion (see `cuda_backend.rs::as_bytes`).
//!
//! There is no CPU fallback: every compute call fails with [`Error::Cuda`] when no device is visible.

pub mod ffi;
#[cfg(feature = "arkworks")]
pub mod ark;

use core::ffi::c_int;
use std::ffi::CStr;
use std::ptr;

pub const FR_BYTES: usize = ffi::PM_FR_BYTES;
pub const FQ_BYTES: usize = ffi::PM_FQ_BYTES;
pub const G1_BYTES: usize = ffi::PM_G1_BYTES;
pub const G2_BYTES: usize = 192;
pub const VK_BYTES: usize = 392;
pub const PROOF_BYTES: usize = 176;

/// Status of a failed call.  The three protocol variants are the reference's panics:
/// `Unsatisfied` = `assert!(rem_poly.is_zero())` (src/prover.rs:108), `Degenerate` = the `h_poly` assertion
/// (src/prover.rs:107), `Remainder` = the opening remainder assertion (src/prover.rs:221).
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Error {
    Cuda(String),
    Arg(String),
    Unsatisfied(String),
    Degenerate(String),
    Remainder(String),
    State(String),
    Unknown(i32, String),
}

impl core::fmt::Display for Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "{self:?}")
    }
}
impl std::error::Error for Error {}

pub type Result<T> = core::result::Result<T, Error>;

fn last_error() -> String {
    // SAFETY: pm_last_error returns a NUL-terminated string owned by the library (thread-local), never null.
    unsafe {
        let p = ffi::pm_last_error();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

fn check(code: c_int) -> Result<()> {
    match code {
        ffi::PM_OK => Ok(()),
        ffi::PM_ERR_CUDA => Err(Error::Cuda(last_error())),
        ffi::PM_ERR_ARG => Err(Error::Arg(last_error())),
        ffi::PM_ERR_UNSATISFIED => Err(Error::Unsatisfied(last_error())),
        ffi::PM_ERR_DEGENERATE => Err(Error::Degenerate(last_error())),
        ffi::PM_ERR_REMAINDER => Err(Error::Remainder(last_error())),
        ffi::PM_ERR_STATE => Err(Error::State(last_error())),
        other => Err(Error::Unknown(other, last_error())),
    }
}

fn need(cond: bool, what: &str) -> Result<()> {
    if cond { Ok(()) } else { Err(Error::Arg(what.to_string())) }
}

/// ABI version of the loaded library (`pm_abi_version`); this crate was generated for version 1.
pub fn abi_version() -> i32 {
    unsafe { ffi::pm_abi_version() }
}

/// Number of visible CUDA devices (<= 0: the backend cannot run).
pub fn device_count() -> i32 {
    unsafe { ffi::pm_device_count() }
}

/// One R1CS matrix of `cs.to_matrices()` (src/generator.rs:46-54) flattened to CSR; `val` holds 32 bytes per entry.
#[derive(Clone, Debug, Default)]
pub struct Csr {
    pub row_ptr: Vec<u64>,
    pub col: Vec<u32>,
    pub val: Vec<u8>,
}

impl Csr {
    /// `rows` as arkworks stores them: `Vec<Vec<(F, usize)>>`, the coefficient given as its 32 Montgomery bytes.
    pub fn from_rows<'a, I, R>(rows: I) -> Self
    where
        I: IntoIterator<Item = R>,
        R: IntoIterator<Item = (&'a [u8], usize)>,
    {
        let mut m = Csr { row_ptr: vec![0], col: Vec::new(), val: Vec::new() };
        for row in rows {
            for (coeff, column) in row {
                assert_eq!(coeff.len(), FR_BYTES);
                m.col.push(column as u32);
                m.val.extend_from_slice(coeff);
            }
            m.row_ptr.push(m.col.len() as u64);
        }
        m
    }
}

/// The SAP source matrices (`SAPMatrices`, src/common.rs:113-127).
#[derive(Clone, Debug, Default)]
pub struct R1cs {
    pub num_instance_variables: u64,
    pub num_r1cs_witness_variables: u64,
    pub num_r1cs_constraints: u64,
    pub a: Csr,
    pub b: Csr,
    pub c: Csr,
}

impl R1cs {
    fn validate(&self) -> Result<()> {
        for m in [&self.a, &self.b, &self.c] {
            need(m.row_ptr.len() as u64 == self.num_r1cs_constraints + 1, "row_ptr length must be constraints + 1")?;
            need(m.val.len() == m.col.len() * FR_BYTES, "32 bytes per matrix coefficient")?;
            need(*m.row_ptr.last().unwrap() as usize == m.col.len(), "row_ptr must end at the entry count")?;
        }
        Ok(())
    }
    fn view(&self) -> ffi::pm_r1cs_view {
        ffi::pm_r1cs_view {
            num_instance_variables: self.num_instance_variables,
            num_r1cs_witness_variables: self.num_r1cs_witness_variables,
            num_r1cs_constraints: self.num_r1cs_constraints,
            a_row_ptr: self.a.row_ptr.as_ptr(), a_col: self.a.col.as_ptr(), a_val: self.a.val.as_ptr(),
            b_row_ptr: self.b.row_ptr.as_ptr(), b_col: self.b.col.as_ptr(), b_val: self.b.val.as_ptr(),
            c_row_ptr: self.c.row_ptr.as_ptr(), c_col: self.c.col.as_ptr(), c_val: self.c.val.as_ptr(),
        }
    }
}

/// Borrowed view of a `ProvingKey` (src/data_structures.rs:56-73).  Every slice holds `len * point_stride` bytes:
/// stride 96 (packed x‖y), `size_of::<G1Affine>()` = 104 (arkworks in memory) or 48 (compressed).
pub struct KeyView<'a> {
    pub r1cs: &'a R1cs,
    pub n: u64,
    pub sigma: u64,
    pub point_stride: usize,
    pub x_powers_g1: &'a [u8],
    pub x_powers_y_alpha_g1: &'a [u8],
    pub x_powers_zh_by_y_alpha_g1: &'a [u8],
    pub x_powers_y_gamma_g1: &'a [u8],
    pub x_powers_y_gamma_z_g1: &'a [u8],
    pub uj_wj_lcs_by_y_alpha_g1: &'a [u8],
}

/// Which key vector `export_key` copies back.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum KeyVector {
    XPowers = 0,
    XPowersYAlpha = 1,
    XPowersZhByYAlpha = 2,
    XPowersYGamma = 3,
    XPowersYGammaZ = 4,
    UjWjLcsByYAlpha = 5,
}

/// A proving key resident on the device (`pm_ctx`).  Created once per `ProvingKey`, reused by every `prove`.
pub struct Context {
    raw: *mut ffi::pm_ctx,
}

// SAFETY: the library serialises its entry points internally (include/polymath_b200.h, "Conventions"); the handle
// itself is an owning pointer.
unsafe impl Send for Context {}

impl Drop for Context {
    fn drop(&mut self) {
        if !self.raw.is_null() {
            unsafe { ffi::pm_ctx_destroy(self.raw) };
            self.raw = ptr::null_mut();
        }
    }
}

impl Context {
    /// Upload an existing proving key (`pm_ctx_create`).
    pub fn from_key(key: &KeyView<'_>) -> Result<Self> {
        key.r1cs.validate()?;
        need(key.point_stride == 48 || key.point_stride >= G1_BYTES, "point stride must be 48, 96 or >= 104")?;
        let len = |s: &[u8]| -> Result<u64> {
            need(s.len() % key.point_stride == 0, "key vector length is not a multiple of the stride")?;
            Ok((s.len() / key.point_stride) as u64)
        };
        let view = ffi::pm_pk_view {
            r1cs: key.r1cs.view(),
            n: key.n,
            sigma: key.sigma,
            point_stride: key.point_stride,
            x_powers_g1: key.x_powers_g1.as_ptr(),
            x_powers_g1_len: len(key.x_powers_g1)?,
            x_powers_y_alpha_g1: key.x_powers_y_alpha_g1.as_ptr(),
            x_powers_y_alpha_g1_len: len(key.x_powers_y_alpha_g1)?,
            x_powers_zh_by_y_alpha_g1: key.x_powers_zh_by_y_alpha_g1.as_ptr(),
            x_powers_zh_by_y_alpha_g1_len: len(key.x_powers_zh_by_y_alpha_g1)?,
            x_powers_y_gamma_g1: key.x_powers_y_gamma_g1.as_ptr(),
            x_powers_y_gamma_g1_len: len(key.x_powers_y_gamma_g1)?,
            x_powers_y_gamma_z_g1: key.x_powers_y_gamma_z_g1.as_ptr(),
            x_powers_y_gamma_z_g1_len: len(key.x_powers_y_gamma_z_g1)?,
            uj_wj_lcs_by_y_alpha_g1: key.uj_wj_lcs_by_y_alpha_g1.as_ptr(),
            uj_wj_lcs_by_y_alpha_g1_len: len(key.uj_wj_lcs_by_y_alpha_g1)?,
        };
        let mut raw = ptr::null_mut();
        // SAFETY: every pointer of `view` borrows from `key`, which outlives the call; the library copies what it keeps.
        check(unsafe { ffi::pm_ctx_create(&view, &mut raw) })?;
        Ok(Context { raw })
    }

    /// The six G1 vectors of `generate_proving_key` (src/generator.rs:81-137) from the trapdoors, built on the device
    /// (`pm_setup`).  Returns the context and `[x]_2`, `[z]_2` (src/generator.rs:144-145; x.c0, x.c1, y.c0, y.c1).
    pub fn setup(r1cs: &R1cs, x: &[u8; FR_BYTES], z: &[u8; FR_BYTES]) -> Result<(Self, [u8; G2_BYTES], [u8; G2_BYTES])> {
        r1cs.validate()?;
        let view = r1cs.view();
        let mut raw = ptr::null_mut();
        let (mut x_g2, mut z_g2) = ([0u8; G2_BYTES], [0u8; G2_BYTES]);
        check(unsafe { ffi::pm_setup(&view, x.as_ptr(), z.as_ptr(), &mut raw, x_g2.as_mut_ptr(), z_g2.as_mut_ptr()) })?;
        Ok((Context { raw }, x_g2, z_g2))
    }

    /// `(n, sigma, columns)` of the key.
    pub fn dims(&self) -> Result<(u64, u64, u64)> {
        let (mut n, mut sigma, mut cols) = (0u64, 0u64, 0u64);
        check(unsafe { ffi::pm_ctx_dims(self.raw, &mut n, &mut sigma, &mut cols) })?;
        Ok((n, sigma, cols))
    }

    /// One key vector copied back to the host with the given stride (96 packed, 104 arkworks layout, 48 compressed).
    pub fn export_key(&self, which: KeyVector, stride: usize) -> Result<Vec<u8>> {
        let mut len = 0u64;
        check(unsafe { ffi::pm_ctx_key_len(self.raw, which as c_int, &mut len) })?;
        let mut out = vec![0u8; len as usize * stride];
        check(unsafe { ffi::pm_ctx_export_key(self.raw, which as c_int, out.as_mut_ptr(), stride) })?;
        Ok(out)
    }

    /// src/prover.rs:73-123 — `[a]_1`, `[c]_1` from the assignment and the two blinding coefficients.
    pub fn prove_phase1(&mut self, instance: &[u8], witness: &[u8], r_a: &[u8; 2 * FR_BYTES]) -> Result<([u8; G1_BYTES], [u8; G1_BYTES])> {
        need(instance.len() % FR_BYTES == 0 && witness.len() % FR_BYTES == 0, "assignments are 32 bytes per element")?;
        let (mut a, mut c) = ([0u8; G1_BYTES], [0u8; G1_BYTES]);
        check(unsafe {
            ffi::pm_prove_phase1(self.raw, instance.as_ptr(), witness.as_ptr(), r_a.as_ptr(), a.as_mut_ptr(), c.as_mut_ptr())
        })?;
        Ok((a, c))
    }

    /// src/prover.rs:128-132 — `a(x1)`.
    pub fn prove_phase2(&mut self, x1: &[u8; FR_BYTES], y1_alpha: &[u8; FR_BYTES]) -> Result<[u8; FR_BYTES]> {
        let mut out = [0u8; FR_BYTES];
        check(unsafe { ffi::pm_prove_phase2(self.raw, x1.as_ptr(), y1_alpha.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// src/prover.rs:142-229 — `[d]_1`.
    pub fn prove_phase3(&mut self, x2: &[u8; FR_BYTES], c_at_x1: &[u8; FR_BYTES]) -> Result<[u8; G1_BYTES]> {
        let mut out = [0u8; G1_BYTES];
        check(unsafe { ffi::pm_prove_phase3(self.raw, x2.as_ptr(), c_at_x1.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// Raw handle for the entry points this layer does not wrap (sharded flow, measurement hooks).
    pub fn as_raw(&mut self) -> *mut ffi::pm_ctx {
        self.raw
    }
}

/// `VariableBaseMSM::msm_unchecked` (src/prover.rs:380-384): bases with `stride` bytes per point, 32-byte scalars.
pub fn msm_g1(bases: &[u8], stride: usize, scalars: &[u8]) -> Result<[u8; G1_BYTES]> {
    need(stride >= G1_BYTES && bases.len() % stride == 0 && scalars.len() % FR_BYTES == 0, "bad MSM operand sizes")?;
    let n = core::cmp::min(bases.len() / stride, scalars.len() / FR_BYTES);
    let mut out = [0u8; G1_BYTES];
    check(unsafe { ffi::pm_msm_g1(bases.as_ptr(), stride, scalars.as_ptr(), n, out.as_mut_ptr()) })?;
    Ok(out)
}

/// `Radix2EvaluationDomain::{fft, ifft_in_place}` (src/prover.rs:241,319,325) in place; `coset_gen` as in `pm_ntt_fr`.
pub fn ntt_fr(data: &mut [u8], inverse: bool, coset_gen: Option<&[u8; FR_BYTES]>) -> Result<()> {
    let n = data.len() / FR_BYTES;
    need(data.len() % FR_BYTES == 0 && n.is_power_of_two(), "NTT size must be a power of two")?;
    let g = coset_gen.map_or(ptr::null(), |g| g.as_ptr());
    check(unsafe { ffi::pm_ntt_fr(data.as_mut_ptr(), n.trailing_zeros(), inverse as c_int, g) })
}

/// `generate()` (src/generator.rs:169-177): `scalars[i] * G` as packed affine points.
pub fn fixed_base_mul(scalars: &[u8]) -> Result<Vec<u8>> {
    need(scalars.len() % FR_BYTES == 0, "32 bytes per scalar")?;
    let n = scalars.len() / FR_BYTES;
    let mut out = vec![0u8; n * G1_BYTES];
    check(unsafe { ffi::pm_fixed_base_mul(scalars.as_ptr(), n, out.as_mut_ptr()) })?;
    Ok(out)
}

/// Compressed (48-byte) G1 encodings to packed affine points; `validate` adds the subgroup check of
/// `deserialize_compressed` (src/data_structures.rs:55-73).
pub fn g1_decompress_batch(encoded: &[u8], validate: bool) -> Result<Vec<u8>> {
    need(encoded.len() % ffi::PM_G1_COMPRESSED_BYTES == 0, "48 bytes per compressed point")?;
    let n = encoded.len() / ffi::PM_G1_COMPRESSED_BYTES;
    let mut out = vec![0u8; n * G1_BYTES];
    check(unsafe { ffi::pm_g1_decompress_batch(encoded.as_ptr(), n, validate as c_int, out.as_mut_ptr()) })?;
    Ok(out)
}

#[cfg(test)]
mod tests {
    use super::*;

    /// Loads the library and checks the version; needs no GPU.
    #[test]
    fn library_loads() {
        assert_eq!(abi_version(), 1);
        assert_eq!(ffi::BOUND_SYMBOLS.len(), 80);
    }

    /// Without a device every compute entry fails loudly (there is no CPU fallback).
    #[test]
    fn no_device_is_an_error() {
        if device_count() > 0 {
            return;
        }
        let scalars = [0u8; FR_BYTES];
        assert!(matches!(fixed_base_mul(&scalars), Err(Error::Cuda(_))));
    }
}
