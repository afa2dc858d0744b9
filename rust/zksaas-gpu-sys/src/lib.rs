//! Raw bindings and thin safe wrappers for `libzksaas_gpu.so` (C ABI: `include/zksaas_gpu.h`).
//!
//! Every `extern` item cites the reference function whose body (or inner arkworks call) it replaces, as
//! `<file>:<line>` relative to the zk-SaaS tree.  Memory images are arkworks 0.4 in-memory forms: `Fr` / `Fq` are four
//! little-endian `u64` Montgomery limbs, `G1Affine` is `{x, y, infinity}` in 72 bytes, `G2Affine` 136 bytes,
//! `Projective` is Jacobian `(X, Y, Z)`; the guards at the end of this file fail the build if a future arkworks changes
//! them (the types are `repr(Rust)`).
#![allow(non_camel_case_types)]
#![allow(clippy::too_many_arguments)]

use ark_bn254::{Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use core::mem::size_of;
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

pub const ZKG_OK: i32 = 0;
pub const ZKG_ERR_LEN_MISMATCH: i32 = -1;
pub const ZKG_ERR_BAD_ARG: i32 = -2;
pub const ZKG_ERR_CUDA: i32 = -3;
pub const ZKG_ERR_OOM: i32 = -4;
pub const ZKG_ERR_UNSUPPORTED: i32 = -5;
pub const ZKG_ERR_NCCL: i32 = -6;

extern "C" {
    pub fn zkg_version() -> i32;
    pub fn zkg_device_count(count: *mut i32) -> i32;
    pub fn zkg_last_error() -> *const c_char;
    pub fn zkg_shutdown() -> i32;

    // G::msm(bases, scalars), dist-primitives/src/dmsm/mod.rs:73 (callers groth16/src/prove.rs:52,106,154,209,219)
    pub fn zkg_msm_bn254_g1(device: i32, bases: *const c_void, base_stride: usize, n_bases: usize,
                            scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> i32;
    pub fn zkg_msm_bn254_g2(device: i32, bases: *const c_void, base_stride: usize, n_bases: usize,
                            scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> i32;
    // the same over a device list (one call drives several GPUs of the box)
    pub fn zkg_msm_bn254_g1_sharded(devices: *const i32, n_devices: i32, bases: *const c_void, base_stride: usize,
                                    n_bases: usize, scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> i32;
    pub fn zkg_msm_bn254_g2_sharded(devices: *const i32, n_devices: i32, bases: *const c_void, base_stride: usize,
                                    n_bases: usize, scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> i32;
    // static CRS shares of PackedProvingKeyShare, groth16/src/proving_key.rs:15-45
    pub fn zkg_bases_register(device: i32, group: i32, bases: *const c_void, base_stride: usize, n: usize,
                              handle: *mut u64) -> i32;
    pub fn zkg_bases_register_sharded(devices: *const i32, n_devices: i32, group: i32, bases: *const c_void,
                                      base_stride: usize, n: usize, handle: *mut u64) -> i32;
    pub fn zkg_bases_release(handle: u64) -> i32;
    pub fn zkg_msm_bn254_registered(handle: u64, scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> i32;
    // pack_from_arkworks_proving_key, groth16/src/proving_key.rs:72-104
    pub fn zkg_crs_det_pack_bn254(device: i32, group: i32, bases: *const c_void, base_stride: usize, n: usize, l: u32,
                                  out_by_party: *const *mut c_void, out_stride: usize) -> i32;

    // king side of d_msm, dist-primitives/src/dmsm/mod.rs:85-87 (unpack_missing_shares over points + sum)
    pub fn zkg_pss_unpack2_bn254_g1(device: i32, l: u32, shares_xyz: *const u64, parties: *const u32, n_recv: u32,
                                    out_unpacked_xyz: *mut u64, out_sum_xyz: *mut u64) -> i32;
    pub fn zkg_pss_unpack2_bn254_g2(device: i32, l: u32, shares_xyz: *const u64, parties: *const u32, n_recv: u32,
                                    out_unpacked_xyz: *mut u64, out_sum_xyz: *mut u64) -> i32;
    // ark-serialize compressed points, mpc-net/src/ser_net.rs:25,40,119
    pub fn zkg_g1_to_wire_bn254(device: i32, points_xyz: *const u64, wire: *mut c_void, n: usize) -> i32;
    pub fn zkg_g1_from_wire_bn254(device: i32, wire: *const c_void, points_xyz: *mut u64, n: usize) -> i32;
    pub fn zkg_g2_to_wire_bn254(device: i32, points_xyz: *const u64, wire: *mut c_void, n: usize) -> i32;
    pub fn zkg_g2_from_wire_bn254(device: i32, wire: *const c_void, points_xyz: *mut u64, n: usize) -> i32;
    pub fn zkg_fr_from_wire_bn254(device: i32, wire: *const c_void, out_mont: *mut u64, n: usize) -> i32;
    pub fn zkg_fr_to_wire_bn254(device: i32, in_mont: *const u64, wire: *mut c_void, n: usize) -> i32;

    // fft1_in_place, dist-primitives/src/dfft/mod.rs:178-208 (+ :159 pre-scale, :254-258 in-mask fused)
    pub fn zkg_fft1_bn254(device: i32, px: *mut u64, mbyl: usize, l: u32, gen: *const u64, pre_scale: *const u64,
                          in_mask: *const u64) -> i32;
    pub fn zkg_fft1_bn254_sharded(devices: *const i32, n_devices: i32, px: *mut u64, mbyl: usize, l: u32,
                                  gen: *const u64, pre_scale: *const u64, in_mask: *const u64) -> i32;
    // king closure of fft2_with_rearrange, dist-primitives/src/dfft/mod.rs:264-304
    pub fn zkg_king_fft2_bn254(device: i32, shares_by_party: *const *const u64, parties: *const u32, n_recv: u32,
                               mbyl: usize, l: u32, gen: *const u64, g: *const u64, rearrange: i32, rand: *const u64,
                               out_by_party: *const *mut u64) -> i32;
    pub fn zkg_king_fft2_bn254_sharded(devices: *const i32, n_devices: i32, shares_by_party: *const *const u64,
                                       parties: *const u32, n_recv: u32, mbyl: usize, l: u32, gen: *const u64,
                                       g: *const u64, rearrange: i32, rand: *const u64,
                                       out_by_party: *const *mut u64) -> i32;
    // king closure of deg_red, dist-primitives/src/utils/deg_red.rs:103-111
    pub fn zkg_deg_red_king_bn254(device: i32, shares_by_party: *const *const u64, parties: *const u32, n_recv: u32,
                                  cols: usize, l: u32, rand: *const u64, out_by_party: *const *mut u64) -> i32;
    pub fn zkg_deg_red_king_bn254_sharded(devices: *const i32, n_devices: i32, shares_by_party: *const *const u64,
                                          parties: *const u32, n_recv: u32, cols: usize, l: u32, rand: *const u64,
                                          out_by_party: *const *mut u64) -> i32;
    // king closure of d_pp, dist-primitives/src/dpp/mod.rs:41-76
    pub fn zkg_dpp_king_bn254(device: i32, shares_by_party: *const *const u64, parties: *const u32, n_recv: u32,
                              cols: usize, l: u32, rand: *const u64, out_by_party: *const *mut u64) -> i32;
    // h = (a + ma)(b + mb) - (c + mc) [* factor], groth16/src/ext_wit.rs:82-86,173-177 (+ dfft/mod.rs:313-317)
    pub fn zkg_qap_h_bn254(device: i32, a: *const u64, b: *const u64, c: *const u64, mask_a: *const u64,
                           mask_b: *const u64, mask_c: *const u64, factor: *const u64, out: *mut u64, n: usize) -> i32;

    // PackedSharingParams over Fr, secret-sharing/src/pss.rs:69-166 (column-major batches)
    pub fn zkg_pss_pack_bn254_fr(device: i32, l: u32, secrets: *const u64, rand: *const u64, shares: *mut u64,
                                 cols: usize) -> i32;
    pub fn zkg_pss_unpack_bn254_fr(device: i32, l: u32, shares: *const u64, secrets: *mut u64, cols: usize) -> i32;
    pub fn zkg_pss_unpack2_bn254_fr(device: i32, l: u32, shares: *const u64, secrets: *mut u64, cols: usize) -> i32;
    // pack_from_witness (groth16/examples/sha256.rs:131-156, layout 0), QAP::pss (groth16/src/qap.rs:99-112, layout 1)
    pub fn zkg_pss_pack_vec_bn254_fr(device: i32, l: u32, layout: i32, x: *const u64, len: usize, rand: *const u64,
                                     out_by_party: *const *mut u64) -> i32;
    // FftMask::sample dist-primitives/src/dfft/mod.rs:30-85, DegRedMask::sample utils/deg_red.rs:40-66
    pub fn zkg_fft_mask_sample_bn254(device: i32, rearrange: i32, g: *const u64, gen: *const u64, m: usize, l: u32,
                                     mask_values: *const u64, rand_in: *const u64, rand_out: *const u64,
                                     in_by_party: *const *mut u64, out_by_party: *const *mut u64) -> i32;
    pub fn zkg_deg_red_mask_sample_bn254(device: i32, num: usize, l: u32, mask_values: *const u64, rand_in: *const u64,
                                         rand_out: *const u64, in_by_party: *const *mut u64,
                                         out_by_party: *const *mut u64) -> i32;
    // stand-alone pieces: fft2_in_place dfft/mod.rs:210-237, distribute_powers :49,:279, fft_in_place_rearrange :322-335
    pub fn zkg_fft2_bn254(device: i32, s1: *mut u64, m: usize, l: u32, gen: *const u64) -> i32;
    pub fn zkg_distribute_powers_bn254(device: i32, v: *mut u64, n: usize, g: *const u64) -> i32;
    pub fn zkg_bitrev_bn254(device: i32, v: *mut u64, n: usize) -> i32;
    pub fn zkg_fr_fft_bn254(device: i32, v: *mut u64, n: usize, offset: *const u64, inverse: i32) -> i32;
}

/// Error of a wrapped call.  `LenMismatch(min_len)` is `G::msm`'s `Err(usize)`; `Other` carries the library's message.
/// Both convert into the reference's `MpcNetError::Generic` through its blanket `impl<T: ToString> From<T>`.
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum GpuError {
    LenMismatch(usize),
    Other(i32, String),
}
impl core::fmt::Display for GpuError {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        match self {
            GpuError::LenMismatch(n) => write!(f, "{n}"),
            GpuError::Other(code, msg) => write!(f, "libzksaas_gpu error {code}: {msg}"),
        }
    }
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(zkg_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<(), GpuError> {
    if rc == ZKG_OK { Ok(()) } else { Err(GpuError::Other(rc, last_error())) }
}
#[inline]
pub fn limbs(x: &Fr) -> *const u64 {
    (x as *const Fr).cast()
}

/// `ark_bn254::G1Projective::msm(bases, scalars)` (dmsm/mod.rs:73) on GPU `device`.
pub fn msm_g1(device: i32, bases: &[G1Affine], scalars: &[Fr]) -> Result<G1Projective, GpuError> {
    let mut out = [0u64; 12];
    let rc = unsafe {
        zkg_msm_bn254_g1(device, bases.as_ptr().cast(), size_of::<G1Affine>(), bases.len(), scalars.as_ptr().cast(),
                         scalars.len(), out.as_mut_ptr())
    };
    if rc == ZKG_ERR_LEN_MISMATCH {
        return Err(GpuError::LenMismatch(bases.len().min(scalars.len())));
    }
    check(rc)?;
    // (X, Y, Z = 1) or (1, 1, 0): a valid Jacobian `Projective`
    Ok(unsafe { core::mem::transmute_copy::<[u64; 12], G1Projective>(&out) })
}
pub fn msm_g2(device: i32, bases: &[G2Affine], scalars: &[Fr]) -> Result<G2Projective, GpuError> {
    let mut out = [0u64; 24];
    let rc = unsafe {
        zkg_msm_bn254_g2(device, bases.as_ptr().cast(), size_of::<G2Affine>(), bases.len(), scalars.as_ptr().cast(),
                         scalars.len(), out.as_mut_ptr())
    };
    if rc == ZKG_ERR_LEN_MISMATCH {
        return Err(GpuError::LenMismatch(bases.len().min(scalars.len())));
    }
    check(rc)?;
    Ok(unsafe { core::mem::transmute_copy::<[u64; 24], G2Projective>(&out) })
}
/// The same MSM spread over `devices` (point-range split inside the library).
pub fn msm_g1_sharded(devices: &[i32], bases: &[G1Affine], scalars: &[Fr]) -> Result<G1Projective, GpuError> {
    let mut out = [0u64; 12];
    let rc = unsafe {
        zkg_msm_bn254_g1_sharded(devices.as_ptr(), devices.len() as i32, bases.as_ptr().cast(), size_of::<G1Affine>(),
                                 bases.len(), scalars.as_ptr().cast(), scalars.len(), out.as_mut_ptr())
    };
    if rc == ZKG_ERR_LEN_MISMATCH {
        return Err(GpuError::LenMismatch(bases.len().min(scalars.len())));
    }
    check(rc)?;
    Ok(unsafe { core::mem::transmute_copy::<[u64; 12], G1Projective>(&out) })
}

/// Device-resident CRS share (`PackedProvingKeyShare::{s,u,w,h}`): register once, one MSM per proof.
pub struct RegisteredBases {
    handle: u64,
    len: usize,
}
impl RegisteredBases {
    pub fn g1(devices: &[i32], bases: &[G1Affine]) -> Result<Self, GpuError> {
        let mut handle = 0u64;
        check(unsafe {
            zkg_bases_register_sharded(devices.as_ptr(), devices.len() as i32, 1, bases.as_ptr().cast(),
                                       size_of::<G1Affine>(), bases.len(), &mut handle)
        })?;
        Ok(Self { handle, len: bases.len() })
    }
    pub fn msm_g1(&self, scalars: &[Fr]) -> Result<G1Projective, GpuError> {
        let mut out = [0u64; 12];
        let rc = unsafe { zkg_msm_bn254_registered(self.handle, scalars.as_ptr().cast(), scalars.len(), out.as_mut_ptr()) };
        if rc == ZKG_ERR_LEN_MISMATCH {
            return Err(GpuError::LenMismatch(self.len.min(scalars.len())));
        }
        check(rc)?;
        Ok(unsafe { core::mem::transmute_copy::<[u64; 12], G1Projective>(&out) })
    }
}
impl Drop for RegisteredBases {
    fn drop(&mut self) {
        unsafe { zkg_bases_release(self.handle) };
    }
}

/// `fft1_in_place(px, pp, gen)` (dfft/mod.rs:178-208) with the optional fused `size_inv` pre-scale (:159) and in-mask (:254-258).
pub fn fft1_in_place(device: i32, px: &mut [Fr], l: usize, gen: &Fr, pre_scale: Option<&Fr>, in_mask: Option<&[Fr]>) -> Result<(), GpuError> {
    if let Some(m) = in_mask {
        assert_eq!(m.len(), px.len(), "in_mask length");
    }
    check(unsafe {
        zkg_fft1_bn254(device, px.as_mut_ptr().cast(), px.len(), l as u32, limbs(gen),
                       pre_scale.map_or(core::ptr::null(), limbs), in_mask.map_or(core::ptr::null(), |m| m.as_ptr().cast()))
    })
}

/// King closure of `fft2_with_rearrange` (dfft/mod.rs:264-304): `shares[r]` came from `parties[r]`; `rand` holds the
/// `m/l * t` packing draws of the host RNG; returns the n per-party vectors (what `transpose(out_shares)` builds).
pub fn king_fft2(devices: &[i32], shares: &[Vec<Fr>], parties: &[u32], l: usize, gen: &Fr, g: &Fr, rearrange: bool,
                 rand: &[Fr]) -> Result<Vec<Vec<Fr>>, GpuError> {
    use ark_ff::Zero;
    let (n, t) = (4 * l, l);
    let mbyl = shares[0].len();
    assert!(shares.iter().all(|v| v.len() == mbyl) && shares.len() == parties.len() && rand.len() == mbyl * t);
    let ins: Vec<*const u64> = shares.iter().map(|v| v.as_ptr().cast()).collect();
    let mut outs: Vec<Vec<Fr>> = vec![vec![Fr::zero(); mbyl]; n];
    let outp: Vec<*mut u64> = outs.iter_mut().map(|v| v.as_mut_ptr().cast()).collect();
    check(unsafe {
        zkg_king_fft2_bn254_sharded(devices.as_ptr(), devices.len() as i32, ins.as_ptr(), parties.as_ptr(),
                                    shares.len() as u32, mbyl, l as u32, limbs(gen), limbs(g), rearrange as i32,
                                    rand.as_ptr().cast(), outp.as_ptr())
    })?;
    Ok(outs)
}

/// King side of `d_msm` (dmsm/mod.rs:85-86): `pp.unpack_missing_shares(&shares, &parties).iter().sum()`.
pub fn dmsm_king_g1(device: i32, l: usize, shares: &[G1Projective], parties: &[u32]) -> Result<G1Projective, GpuError> {
    assert_eq!(shares.len(), parties.len());
    let mut out = [0u64; 12];
    check(unsafe {
        zkg_pss_unpack2_bn254_g1(device, l as u32, shares.as_ptr().cast(), parties.as_ptr(), shares.len() as u32,
                                 core::ptr::null_mut(), out.as_mut_ptr())
    })?;
    Ok(unsafe { core::mem::transmute_copy::<[u64; 12], G1Projective>(&out) })
}

// ---- layout guards: arkworks' structs are repr(Rust); the C side assumes these images -------------------------------
const _: () = assert!(size_of::<Fr>() == 32);
const _: () = assert!(size_of::<ark_bn254::Fq>() == 32);
const _: () = assert!(size_of::<G1Affine>() == 72);
const _: () = assert!(size_of::<G2Affine>() == 136);
const _: () = assert!(size_of::<G1Projective>() == 96);
const _: () = assert!(size_of::<G2Projective>() == 192);
const _: () = assert!(core::mem::align_of::<Fr>() == 8);
