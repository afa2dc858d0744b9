"""zksaas_b200 -- B200 (sm_100a) implementation of the data-parallel prover core of zk-SaaS.

The product is the C-ABI shared library `libzksaas_gpu.so` (include/zksaas_gpu.h); this package is
the thin host-side mirror of the reference's Rust interface for that path (same names, argument
meaning and error behaviour) used by the parity tests and the benchmark.  There is no CPU
fallback: importing works anywhere, but every compute call fails loudly without the CUDA library
or without a device.
"""
from .capi import ZkgError, lib, lib_path  # noqa: F401
from .api import (  # noqa: F401
    PackedSharingParams, MsmLengthMismatch, Radix2EvaluationDomain, FftMask, MsmMask, DegRedMask,
    msm_g1, msm_g2, fft1_in_place, fft2_in_place, fft_in_place_rearrange, distribute_powers,
    king_fft2, deg_red_king, dpp_king, pack_vec, pack_from_witness, qap_pss_pack, qap_pss, group_generator, transpose, d_fft, d_ifft, d_msm, deg_red, d_pp, LocalTestNet,
    pss_unpack2_group, group_to_wire, group_from_wire, qap_h,
)
