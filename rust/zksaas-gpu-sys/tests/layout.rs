//! Field offsets the C ABI relies on (sizes are guarded at compile time in src/lib.rs).
use ark_bn254::{Fr, G1Affine, G1Projective, G2Affine};
use ark_ec::{AffineRepr, CurveGroup, Group};
use ark_ff::{One, PrimeField, Zero};
use memoffset::offset_of;

#[test]
fn affine_and_projective_field_offsets() {
    assert_eq!(offset_of!(G1Affine, x), 0);
    assert_eq!(offset_of!(G1Affine, y), 32);
    assert_eq!(offset_of!(G1Affine, infinity), 64);
    assert_eq!(offset_of!(G2Affine, x), 0);
    assert_eq!(offset_of!(G2Affine, y), 64);
    assert_eq!(offset_of!(G2Affine, infinity), 128);
    assert_eq!(offset_of!(G1Projective, x), 0);
    assert_eq!(offset_of!(G1Projective, y), 32);
    assert_eq!(offset_of!(G1Projective, z), 64);
}

#[test]
fn fr_is_four_montgomery_limbs() {
    // one() in Montgomery form is R = 2^256 mod r
    let one = Fr::one();
    let raw: [u64; 4] = unsafe { core::mem::transmute(one) };
    assert_eq!(raw, [0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f]);
    assert_eq!(Fr::MODULUS_BIT_SIZE, 254);
    let zero: [u64; 4] = unsafe { core::mem::transmute(Fr::zero()) };
    assert_eq!(zero, [0; 4]);
}

#[test]
fn identity_images() {
    let id = G1Affine::identity();
    assert!(id.infinity);
    let p = G1Projective::zero();
    let raw: [u64; 12] = unsafe { core::mem::transmute(p) };
    assert_eq!(&raw[8..12], &[0u64; 4]); // z == 0
    let g = G1Projective::generator().into_affine();
    assert!(!g.infinity);
}

#[cfg(feature = "gpu")]
#[test]
fn msm_matches_arkworks() {
    use ark_ec::VariableBaseMSM;
    use ark_std::UniformRand;
    let rng = &mut ark_std::test_rng();
    let n = 1 << 10;
    let bases: Vec<G1Affine> = (0..n).map(|_| G1Projective::rand(rng).into_affine()).collect();
    let scalars: Vec<Fr> = (0..n).map(|_| Fr::rand(rng)).collect();
    let want = G1Projective::msm(&bases, &scalars).unwrap();
    let got = zksaas_gpu_sys::msm_g1(0, &bases, &scalars).unwrap();
    assert_eq!(want, got);
    assert_eq!(want.into_affine(), got.into_affine());
    assert_eq!(zksaas_gpu_sys::msm_g1(0, &bases, &scalars[1..]), Err(zksaas_gpu_sys::GpuError::LenMismatch(n - 1)));
}
