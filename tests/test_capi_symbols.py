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
