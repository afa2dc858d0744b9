"""Device-resident G1 MSM at 2^lg points with skewed scalars (witness-like: 30 % zero, 30 % one, 40 % uniform; all equal),
registered and generic path: total and phase times."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); reps = 3
n = 1 << lg
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = capi.ctx_p(); capi.check(lib.zkg_ctx_create(0, C.c_void_p(st.cuda_stream), C.byref(ctx)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
s = rnd(n)
b = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(b.data_ptr())))
h = C.c_uint64(0)
capi.check(lib.zkg_bases_register_dev(ctx, 1, C.c_void_p(b.data_ptr()), n, C.byref(h)))
o = torch.zeros(12, dtype=torch.int64, device="cuda")
lib.zkg_ctx_set_profiling(ctx, 1)
one_mont = torch.tensor([x - (1 << 64) if x >= (1 << 63) else x for x in
                         (0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f)], dtype=torch.int64, device="cuda")
for kind in ("uniform", "witness_like", "all_equal"):
    a = rnd(n)
    if kind == "witness_like":
        u = torch.rand(n, device="cuda", generator=g)
        a[u < 0.6] = one_mont
        a[u < 0.3] = 0
    elif kind == "all_equal":
        a[:] = a[0].clone()
    for name, run in (("registered", lambda: capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr()), 0))),
                      ("generic", lambda: capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(b.data_ptr()), C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr()))))):
        run(); capi.check(lib.zkg_ctx_sync(ctx))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): run()
        e1.record(st); e1.synchronize()
        ph = []
        for k in range(3):
            f = C.c_float(0); lib.zkg_ctx_phase_ms(ctx, k, C.byref(f)); ph.append(round(f.value, 3))
        print(f"n=2^{lg} {kind:13s} {name:10s}: {e0.elapsed_time(e1)/reps:.3f} ms  phases(sort,acc,reduce)={ph}", flush=True)
