"""Experiment: one 2^lg-point registered MSM as TWO half-range partial MSMs on two streams (second one high priority)
+ combine, against the plain single-stream call.  Measures how much of the sort / reduction can hide under the
other half's bucket accumulation."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); reps = 5
n = 1 << lg; h = n // 2
sA = torch.cuda.Stream(); sB = torch.cuda.Stream(priority=-1)
torch.cuda.set_stream(sA)
ctxA, ctxB = capi.ctx_p(), capi.ctx_p()
capi.check(lib.zkg_ctx_create(0, C.c_void_p(sA.cuda_stream), C.byref(ctxA)))
capi.check(lib.zkg_ctx_create(0, C.c_void_p(sB.cuda_stream), C.byref(ctxB)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
a, s = rnd(n), rnd(n)
b = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
capi.check(lib.zkg_fixed_base_dev(ctxA, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(b.data_ptr())))
hf, h1, h2 = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
os.environ["ZKG_MSM_PREP_C"] = "20"
capi.check(lib.zkg_bases_register_dev(ctxA, 1, C.c_void_p(b.data_ptr()), n, C.byref(hf)))
capi.check(lib.zkg_bases_register_dev(ctxA, 1, C.c_void_p(b[:h].data_ptr()), h, C.byref(h1)))
capi.check(lib.zkg_bases_register_dev(ctxA, 1, C.c_void_p(b[h:].data_ptr()), h, C.byref(h2)))
capi.check(lib.zkg_ctx_sync(ctxA))
o = torch.zeros(12, dtype=torch.int64, device="cuda")
o2 = torch.zeros(12, dtype=torch.int64, device="cuda")
parts = torch.zeros((2, 16), dtype=torch.int64, device="cuda")
def single():
    capi.check(lib.zkg_msm_bn254_registered_dev(ctxA, hf.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr()), 0))
evA, evB = torch.cuda.Event(), torch.cuda.Event()
def split():
    evA.record(sA); sB.wait_event(evA)
    capi.check(lib.zkg_msm_bn254_registered_dev(ctxA, h1.value, C.c_void_p(a.data_ptr()), h, C.c_void_p(parts[0].data_ptr()), 1))
    capi.check(lib.zkg_msm_bn254_registered_dev(ctxB, h2.value, C.c_void_p(a[h:].data_ptr()), h, C.c_void_p(parts[1].data_ptr()), 1))
    evB.record(sB); sA.wait_event(evB)
    capi.check(lib.zkg_msm_combine_dev(ctxA, 1, C.c_void_p(parts.data_ptr()), 2, C.c_void_p(o2.data_ptr())))
def timed(f):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sA)
    for _ in range(reps): f()
    e1.record(sA); e1.synchronize()
    return e0.elapsed_time(e1) / reps
t1 = timed(single); t2 = timed(split)
print(f"n=2^{lg}: single {t1:.3f} ms, two half-range partials on two streams + combine {t2:.3f} ms, equal={bool((o == o2).all())}", flush=True)
