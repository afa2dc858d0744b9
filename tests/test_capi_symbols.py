"""CPU: the C-ABI library loads and exports every symbol include/zksaas_gpu.h declares, and the
product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zksaas_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import zksaas_b200
    from zksaas_b200 import capi
    if not os.path.exists(zksaas_b200.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(zksaas_b200.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/zksaas_gpu.h but not exported"
    # the ctypes binding covers the whole header, nothing more
    assert sorted(capi.SIGNATURES) == syms


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import zksaas_b200 as z
    with pytest.raises(z.ZkgError) as ei:
        z.msm_g1(np.zeros((1, 72), dtype=np.uint8), np.zeros((1, 4), dtype=np.uint64))
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(z.ZkgError):
        z.PackedSharingParams.new(2).det_pack(np.zeros((2, 4), dtype=np.uint64))
    with pytest.raises(z.ZkgError):
        z.fft1_in_place(np.zeros((8, 4), dtype=np.uint64), z.PackedSharingParams.new(2),
                        z.Radix2EvaluationDomain.new(16).group_gen())


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "zk-saas_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "zkoracle" not in text and "pyref" not in text and "oracle_lib" not in text, f
    out = os.popen(f"ldd {os.path.join(pkg, 'libzksaas_gpu.so')} 2>/dev/null").read()
    assert "zkoracle" not in out


def test_host_mirror_argument_checks():
    import zksaas_b200 as z
    with pytest.raises(ValueError):
        z.PackedSharingParams.new(3)
    dom = z.Radix2EvaluationDomain.new(1000)
    assert dom.size() == 1024                       # Radix2EvaluationDomain::new rounds up
    from zksaas_b200.api import fr_value, fr_image, R_MOD
    w = fr_value(dom.group_gen())
    assert pow(w, 1024, R_MOD) == 1 and pow(w, 512, R_MOD) != 1
    assert fr_value(dom.size_inv()) * 1024 % R_MOD == 1
    assert fr_value(fr_image(R_MOD - 5)) == R_MOD - 5


def _c_prototypes():
    """name -> number of parameters, from the header."""
    text = open(os.path.join(ROOT, "include", "zksaas_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(zkg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_rust_crate_binds_only_exported_symbols():
    """rust/zksaas-gpu-sys/src/lib.rs cannot be compiled here (no cargo); keep it honest mechanically: every extern fn
    it declares exists in the header with the same parameter count, and is exported by the built library."""
    import zksaas_b200
    src = open(os.path.join(ROOT, "rust", "zksaas-gpu-sys", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//[^\n]*", "", block)
    protos = _c_prototypes()
    lib = C.CDLL(zksaas_b200.lib_path())
    found = re.findall(r"pub fn (zkg_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S)
    assert len(found) >= 35
    for name, args in found:
        assert name in protos, f"{name} bound by the Rust crate but not declared in the header"
        n_args = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n_args == protos[name], f"{name}: Rust declares {n_args} parameters, the header {protos[name]}"
        assert hasattr(lib, name)
    # every host-pointer entry point of the header is bound (the `_dev` / ctx ones are for device-resident callers)
    host_api = {s for s in protos if not s.endswith("_dev") and not s.startswith(("zkg_ctx_", "zkg_shared_")) and s not in
                ("zkg_field_op",)}
    assert host_api <= {name for name, _ in found}, sorted(host_api - {name for name, _ in found})
    # the patch touches the five functions it claims to
    patch = open(os.path.join(ROOT, "rust", "patches", "zk-saas-gpu.patch")).read()
    for f in ("dmsm/mod.rs", "dfft/mod.rs", "utils/deg_red.rs"):
        assert f in patch
    for sym in ("msm_g1_sharded", "zkg_fft1_bn254_sharded", "king_fft2", "zkg_deg_red_king_bn254", "dmsm_king_g1"):
        assert sym in patch and sym in src
