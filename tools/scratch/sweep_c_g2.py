"""Sweep the MSM window size at a given log2 n (generic path) and print phase timings."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); cs = [int(x) for x in sys.argv[2].split(",")]
n = 1 << lg
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = capi.ctx_p(); capi.check(lib.zkg_ctx_create(0, C.c_void_p(st.cuda_stream), C.byref(ctx)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
a, s = rnd(n), rnd(n)
b = torch.empty((n, 128), dtype=torch.uint8, device="cuda"); o = torch.zeros(24, dtype=torch.int64, device="cuda")
capi.check(lib.zkg_fixed_base_dev(ctx, 2, C.c_void_p(s.data_ptr()), n, C.c_void_p(b.data_ptr())))
lib.zkg_ctx_set_profiling(ctx, 1)
for c in cs:
    os.environ["ZKG_MSM_C"] = str(c)
    for _ in range(2):
        capi.check(lib.zkg_msm_bn254_g2_dev(ctx, C.c_void_p(b.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr())))
    capi.check(lib.zkg_ctx_sync(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        capi.check(lib.zkg_msm_bn254_g2_dev(ctx, C.c_void_p(b.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr())))
    e1.record(st); e1.synchronize()
    ph = []
    for k in range(3):
        f = C.c_float(0); lib.zkg_ctx_phase_ms(ctx, k, C.byref(f)); ph.append(round(f.value, 3))
    print(f"n=2^{lg} c={c}: {e0.elapsed_time(e1)/5:.3f} ms  phases(sort,acc,reduce)={ph}", flush=True)
