//! Zero-copy views of arkworks values for the C ABI (feature `arkworks`).
//!
//! `ark_ff::Fp<MontBackend<_, N>, N>` is `struct Fp(pub BigInt<N>, PhantomData<_>)`: the only non-zero-sized field is
//! `[u64; N]`, the little-endian limbs of the MONTGOMERY form — exactly the wire form of `include/polymath_b200.h`.
//! The layout of a `repr(Rust)` struct is not promised, but a struct whose size equals the size of its one
//! non-zero-sized field cannot have padding or an offset, so the size assertions below make the byte views sound.
//!
//! `Affine<P> { x, y, infinity: bool }` has padding after the flag, so points are passed by POINTER + stride (the
//! library reads x at 0, y at 48 and the flag byte at 96); `g1_layout_is_native` checks those offsets once and
//! `g1_pack` repacks when a compiler lays the struct out differently.
use core::any::TypeId;
use core::mem::size_of;

use ark_bls12_381::{Fq, Fr, G1Affine};
use ark_ec::AffineRepr;
use ark_ff::{BigInt, Fp, MontBackend, MontConfig};

use crate::{FQ_BYTES, FR_BYTES, G1_BYTES};

const _: () = assert!(size_of::<Fr>() == FR_BYTES);
const _: () = assert!(size_of::<Fq>() == FQ_BYTES);

/// `&[Fp<_, N>]` as the bytes of its Montgomery limbs.
pub fn fp_slice_as_bytes<P: MontConfig<N>, const N: usize>(s: &[Fp<MontBackend<P, N>, N>]) -> &[u8] {
    assert_eq!(size_of::<Fp<MontBackend<P, N>, N>>(), 8 * N);
    // SAFETY: the element is exactly N u64 limbs (asserted), u8 has alignment 1, the lifetime is inherited.
    unsafe { core::slice::from_raw_parts(s.as_ptr() as *const u8, s.len() * 8 * N) }
}

/// One element from its 8·N Montgomery bytes (no reduction: the library only writes reduced values).
pub fn fp_from_bytes<P: MontConfig<N>, const N: usize>(bytes: &[u8]) -> Fp<MontBackend<P, N>, N> {
    assert_eq!(bytes.len(), 8 * N);
    let mut limbs = [0u64; N];
    for (limb, chunk) in limbs.iter_mut().zip(bytes.chunks_exact(8)) {
        *limb = u64::from_le_bytes(chunk.try_into().unwrap());
    }
    Fp::new_unchecked(BigInt::new(limbs))
}

/// `&[T]` seen as `&[U]` when the two are the same type (how generic code over `F: PrimeField` reaches the
/// BLS12-381 path without `unsafe` of its own).
pub fn same_type_slice<T: 'static, U: 'static>(s: &[T]) -> Option<&[U]> {
    if TypeId::of::<T>() == TypeId::of::<U>() {
        // SAFETY: T and U are the same type.
        Some(unsafe { core::slice::from_raw_parts(s.as_ptr() as *const U, s.len()) })
    } else {
        None
    }
}

/// A value of type `T` as type `U` when the two are the same type.
pub fn same_type_value<T: 'static + Copy, U: 'static + Copy>(v: T) -> Option<U> {
    (&v as &dyn core::any::Any).downcast_ref::<U>().copied()
}

/// True when this compiler lays `G1Affine` out as x (48 B) | y (48 B) | infinity (1 B): then key vectors are passed
/// in place with `point_stride = size_of::<G1Affine>()`.
pub fn g1_layout_is_native() -> bool {
    let p = G1Affine::generator();
    let base = &p as *const G1Affine as usize;
    (&p.x as *const Fq as usize) - base == 0
        && (&p.y as *const Fq as usize) - base == FQ_BYTES
        && (&p.infinity as *const bool as usize) - base == 2 * FQ_BYTES
        && size_of::<G1Affine>() >= G1_BYTES + 1
}

/// Pointer view of a point vector for `KeyView` when the layout is native.  The bytes between the flag and the next
/// point are padding and never read by the library.
pub fn g1_slice_as_strided_bytes(points: &[G1Affine]) -> (&[u8], usize) {
    assert!(g1_layout_is_native());
    // SAFETY: used only as (pointer, length) by the C side, which reads the 97 initialised bytes of every element.
    let bytes = unsafe { core::slice::from_raw_parts(points.as_ptr() as *const u8, points.len() * size_of::<G1Affine>()) };
    (bytes, size_of::<G1Affine>())
}

/// Packed 96-byte form (x | y, (0, 0) = infinity) — the fallback when the layout is not native.
pub fn g1_pack(points: &[G1Affine]) -> Vec<u8> {
    let mut out = vec![0u8; points.len() * G1_BYTES];
    for (p, dst) in points.iter().zip(out.chunks_exact_mut(G1_BYTES)) {
        if !p.infinity {
            dst[..FQ_BYTES].copy_from_slice(fp_slice_as_bytes(core::slice::from_ref(&p.x)));
            dst[FQ_BYTES..].copy_from_slice(fp_slice_as_bytes(core::slice::from_ref(&p.y)));
        }
    }
    out
}

/// A point written by the library (96 bytes, canonical affine, (0, 0) = infinity).
pub fn g1_from_bytes(bytes: &[u8; G1_BYTES]) -> G1Affine {
    if bytes.iter().all(|b| *b == 0) {
        return G1Affine::identity();
    }
    G1Affine::new_unchecked(fp_from_bytes(&bytes[..FQ_BYTES]), fp_from_bytes(&bytes[FQ_BYTES..]))
}

/// Packed points (e.g. `Context::export_key(.., 96)`) as arkworks values.
pub fn g1_vec_from_bytes(bytes: &[u8]) -> Vec<G1Affine> {
    bytes.chunks_exact(G1_BYTES).map(|c| g1_from_bytes(c.try_into().unwrap())).collect()
}
