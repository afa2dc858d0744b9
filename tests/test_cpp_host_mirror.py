"""The C++ host-side mirror of the reference interface (include/zksaas_host.hpp) and its protocol-level test program
(tests/cpp/host_mirror_test.cpp: the reference's d_ifft / d_fft / deg_red / d_msm / pss tests written against the C++
names, as they are written against the Rust ones).  CPU: the program compiles and links against the C ABI, and the
host-side domain constants equal the oracle's.  GPU: the program runs green."""
import os
import subprocess

import pytest

import oracle_lib as ol
from oracle_lib import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
R = pyref.R_MOD


def build_host_mirror_test():
    """g++ build of tests/cpp/host_mirror_test.cpp (also called by __graft_entry__.build(), so the binary travels)."""
    ol.oracle()                                                    # makes sure oracle/libzkoracle.so exists
    src = os.path.join(HERE, "cpp", "host_mirror_test.cpp")
    out_dir = os.path.join(HERE, "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "host_mirror_test")
    deps = [src, os.path.join(ROOT, "include", "zksaas_host.hpp"), os.path.join(ROOT, "include", "zksaas_gpu.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L" + os.path.join(ROOT, "zk-saas_b200"), "-lzksaas_gpu", "-L" + os.path.join(ROOT, "oracle"), "-lzkoracle",
                               "-Wl,-rpath,$ORIGIN/../../../zk-saas_b200", "-Wl,-rpath,$ORIGIN/../../../oracle", "-pthread"], env=env)
    return exe


def test_host_mirror_compiles_and_host_constants_match_the_oracle():
    exe = build_host_mirror_test()
    out = subprocess.run([exe, "--host-only"], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    dom = pyref.Radix2Domain(1000)                                 # ceil to 1024, ark_std::log2
    mont = lambda v: "%064x" % (v % R * (1 << 256) % R)
    assert int(got["size"]) == dom.size == 1024
    assert got["group_gen"] == mont(dom.group_gen)
    assert got["group_gen_inv"] == mont(dom.group_gen_inv)
    assert got["size_inv"] == mont(dom.size_inv)
    assert got["element5"] == mont(dom.element(5))
    assert got["generator"] == mont(pyref.FR_GENERATOR)
    assert got["minus_one"] == mont(R - 1)
    assert got["gen_times_inv"] == mont(1)


@pytest.mark.gpu
def test_host_mirror_protocol_tests_on_gpu():
    exe = build_host_mirror_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0 and "ALL PASS" in r.stdout and "FAIL" not in r.stdout
