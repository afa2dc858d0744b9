"""Window sweep of the device-resident REGISTERED MSM (ZKG_MSM_PREP_C read at registration).  argv: g1|g2 lg:c_lo:c_hi ..."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
g2 = sys.argv[1] == "g2"
grp = 2 if g2 else 1
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = capi.ctx_p(); capi.check(lib.zkg_ctx_create(0, C.c_void_p(st.cuda_stream), C.byref(ctx)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
lib.zkg_ctx_set_profiling(ctx, 1)
for spec in sys.argv[2:]:
    lg, clo, chi = (int(x) for x in spec.split(":"))
    n = 1 << lg
    a, s = rnd(n), rnd(n)
    b = torch.empty((n, 128 if g2 else 64), dtype=torch.uint8, device="cuda")
    capi.check(lib.zkg_fixed_base_dev(ctx, grp, C.c_void_p(s.data_ptr()), n, C.c_void_p(b.data_ptr())))
    o = torch.zeros(24, dtype=torch.int64, device="cuda")
    ref = None
    for c in [0] + list(range(clo, chi + 1)):
        if c: os.environ["ZKG_MSM_PREP_C"] = str(c)
        else: os.environ.pop("ZKG_MSM_PREP_C", None)
        h = C.c_uint64(0)
        rc = lib.zkg_bases_register_dev(ctx, grp, C.c_void_p(b.data_ptr()), n, C.byref(h))
        if rc != 0:
            print(f"n=2^{lg} c={c}: register failed rc={rc}", flush=True); continue
        def run():
            capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h.value, C.c_void_p(a.data_ptr()), n, C.c_void_p(o.data_ptr()), 0))
        run(); capi.check(lib.zkg_ctx_sync(ctx))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3): run()
        e1.record(st); e1.synchronize()
        ph = []
        for k in range(3):
            f = C.c_float(0); lib.zkg_ctx_phase_ms(ctx, k, C.byref(f)); ph.append(round(f.value, 3))
        res = o.cpu().numpy().copy()
        if ref is None: ref = res
        print(f"{'G2' if g2 else 'G1'} registered n=2^{lg} c={c or 'auto'}: {e0.elapsed_time(e1)/3:.3f} ms phases={ph} same={bool((res == ref).all())}", flush=True)
        lib.zkg_bases_release(h.value)
