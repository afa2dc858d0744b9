"""ctypes binding of libzksaas_gpu.so (declarations mirror include/zksaas_gpu.h one to one)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
pp_u64 = C.POINTER(u64p)
ctx_p = C.c_void_p

ZKG_OK = 0
ZKG_ERR_LEN_MISMATCH = -1
ZKG_ERR_BAD_ARG = -2
ZKG_ERR_CUDA = -3
ZKG_ERR_OOM = -4
ZKG_ERR_UNSUPPORTED = -5
ZKG_ERR_NCCL = -6

# name -> (restype, argtypes); every symbol include/zksaas_gpu.h declares
SIGNATURES = {
    "zkg_version": (C.c_int32, []),
    "zkg_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "zkg_last_error": (C.c_char_p, []),
    "zkg_ctx_create": (C.c_int32, [C.c_int32, C.c_void_p, C.POINTER(ctx_p)]),
    "zkg_ctx_destroy": (C.c_int32, [ctx_p]),
    "zkg_ctx_sync": (C.c_int32, [ctx_p]),
    "zkg_ctx_stream": (C.c_void_p, [ctx_p]),
    "zkg_ctx_launch_count": (C.c_int32, [ctx_p, u64p]),
    "zkg_ctx_set_profiling": (C.c_int32, [ctx_p, C.c_int32]),
    "zkg_ctx_phase_ms": (C.c_int32, [ctx_p, C.c_int32, C.POINTER(C.c_float)]),
    "zkg_shutdown": (C.c_int32, []),
    "zkg_msm_bn254_g1": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_msm_bn254_g2": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_crs_det_pack_bn254": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32,
                                            C.POINTER(C.c_void_p), C.c_size_t]),
    "zkg_bases_register": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, u64p]),
    "zkg_bases_register_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_void_p, C.c_size_t, u64p]),
    "zkg_bases_release": (C.c_int32, [C.c_uint64]),
    "zkg_msm_bn254_registered_dev": (C.c_int32, [ctx_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32]),
    "zkg_msm_bn254_registered": (C.c_int32, [C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_pack_bases_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "zkg_msm_bn254_g1_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_msm_bn254_g2_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_msm_bn254_partial_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_msm_combine_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_fixed_base_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_fft1_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zkg_fft1_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zkg_fft1_shard_local_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                                    C.c_void_p, C.c_void_p]),
    "zkg_fft1_shard_outer_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32,
                                                    C.c_void_p, C.c_void_p]),
    "zkg_king_fft2_bn254": (C.c_int32, [C.c_int32, pp_u64, u32p, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p, pp_u64]),
    "zkg_king_fft2_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, u32p, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p,
                                             C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "zkg_king_stage1_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, u32p, C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t,
                                               C.c_uint32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "zkg_king_stage2_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
    "zkg_deg_red_king_bn254": (C.c_int32, [C.c_int32, pp_u64, u32p, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p, pp_u64]),
    "zkg_deg_red_king_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, u32p, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "zkg_dpp_king_bn254": (C.c_int32, [C.c_int32, pp_u64, u32p, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p, pp_u64]),
    "zkg_pss_pack_vec_bn254_fr": (C.c_int32, [C.c_int32, C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, pp_u64]),
    "zkg_pss_pack_bn254_fr": (C.c_int32, [C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_pss_unpack_bn254_fr": (C.c_int32, [C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_pss_unpack2_bn254_fr": (C.c_int32, [C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_fft2_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
    "zkg_distribute_powers_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_bitrev_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t]),
    "zkg_fr_fft_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32]),
    "zkg_fr_from_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_fr_to_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_pss_unpack2_bn254_g1": (C.c_int32, [C.c_int32, C.c_uint32, C.c_void_p, u32p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "zkg_pss_unpack2_bn254_g2": (C.c_int32, [C.c_int32, C.c_uint32, C.c_void_p, u32p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "zkg_g1_to_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_g1_from_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_g2_to_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_g2_from_wire_bn254": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_qap_h_bn254": (C.c_int32, [C.c_int32] + [C.c_void_p] * 8 + [C.c_size_t]),
    "zkg_qap_h_bn254_dev": (C.c_int32, [ctx_p] + [C.c_void_p] * 8 + [C.c_size_t]),
    "zkg_fft_mask_sample_bn254": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p,
                                               C.c_void_p, C.c_void_p, pp_u64, pp_u64]),
    "zkg_deg_red_mask_sample_bn254": (C.c_int32, [C.c_int32, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, pp_u64, pp_u64]),
    "zkg_msm_bn254_g1_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_msm_bn254_g2_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "zkg_bases_register_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, u64p]),
    "zkg_king_fft2_bn254_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, pp_u64, u32p, C.c_uint32, C.c_size_t, C.c_uint32,
                                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, pp_u64]),
    "zkg_deg_red_king_bn254_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, pp_u64, u32p, C.c_uint32, C.c_size_t, C.c_uint32,
                                                    C.c_void_p, pp_u64]),
    "zkg_fft1_bn254_sharded": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zkg_shared_alloc": (C.c_int32, [ctx_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "zkg_shared_open": (C.c_int32, [ctx_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "zkg_shared_close": (C.c_int32, [ctx_p, C.c_void_p]),
    "zkg_shared_free": (C.c_int32, [ctx_p, C.c_void_p]),
    "zkg_king_stage1_scatter_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, u32p, C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32,
                                                       C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_uint32]),
    "zkg_fft1_shard_local_scatter_bn254_dev": (C.c_int32, [ctx_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                                            C.c_void_p, C.POINTER(C.c_void_p)]),
    "zkg_field_op_dev": (C.c_int32, [ctx_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "zkg_field_op": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
}


class ZkgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libzksaas_gpu error {code}: {msg}")
        self.code = code
        self.msg = msg


def lib_path():
    return os.path.join(_HERE, "libzksaas_gpu.so")


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built: there is no other path."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ZkgError(ZKG_ERR_CUDA, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                         "(zksaas_b200 has no CPU fallback)")
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc):
    if rc != ZKG_OK:
        raise ZkgError(rc, lib().zkg_last_error().decode(errors="replace"))
    return rc
