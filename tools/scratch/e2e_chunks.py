"""e2e host-pointer MSM (zkg_msm_bn254_g1, pinned buffers, bases re-shipped) under different chunk plans (ZKG_MSM_BOUNDS)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); n = 1 << lg
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = capi.ctx_p(); capi.check(lib.zkg_ctx_create(0, C.c_void_p(st.cuda_stream), C.byref(ctx)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
a, s = rnd(n), rnd(n)
b = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s.data_ptr()), n, C.c_void_p(b.data_ptr())))
capi.check(lib.zkg_ctx_sync(ctx))
hb = np.zeros((n, 72), dtype=np.uint8); hb[:, :64] = b.cpu().numpy()
pb = torch.from_numpy(hb).pin_memory(); pa = a.cpu().pin_memory()
po = torch.zeros(12, dtype=torch.int64).pin_memory()
h = C.c_uint64(0)
capi.check(lib.zkg_bases_register(0, 1, C.c_void_p(hb.ctypes.data), 72, n, C.byref(h)))
ref = None
plans = ["", "4,16,32,48,64", "2,8,20,34,49,64", "1,4,12,24,37,50,64", "2,6,14,26,38,51,64", "1,3,8,16,28,40,52,64", "3,12,28,46,64", "2,10,26,45,64", "2,8,24,44,64"]
for cenv in ([""] if len(sys.argv) < 3 else sys.argv[2].split(",")):
    for plan in plans:
        os.environ["ZKG_MSM_BOUNDS"] = plan
        if cenv: os.environ["ZKG_MSM_C"] = cenv
        for _ in range(3):
            capi.check(lib.zkg_msm_bn254_g1(0, C.c_void_p(pb.data_ptr()), 72, n, C.c_void_p(pa.data_ptr()), n, C.c_void_p(po.data_ptr())))
        t0 = time.perf_counter()
        for _ in range(8):
            capi.check(lib.zkg_msm_bn254_g1(0, C.c_void_p(pb.data_ptr()), 72, n, C.c_void_p(pa.data_ptr()), n, C.c_void_p(po.data_ptr())))
        dt = (time.perf_counter() - t0) / 8
        cur = po.numpy().copy()
        if ref is None: ref = cur
        print(f"c={cenv or 'auto'} plan=[{plan or 'default'}]: {dt*1e3:.3f} ms  {n/dt/1e6:.1f} Mpts/s  same={bool((cur == ref).all())}", flush=True)

rref = None
for plan in ["", "1,4,16,32,48,64", "1,3,8,16,28,40,52,64", "1,4,12,24,37,50,64", "2,8,24,44,64", "1,5,21,42,64", "1,9,32,64", "1,17,64", "2,64", "1,64"]:
    os.environ["ZKG_MSM_BOUNDS"] = plan
    for _ in range(3):
        capi.check(lib.zkg_msm_bn254_registered(h.value, C.c_void_p(pa.data_ptr()), n, C.c_void_p(po.data_ptr())))
    t0 = time.perf_counter()
    for _ in range(8):
        capi.check(lib.zkg_msm_bn254_registered(h.value, C.c_void_p(pa.data_ptr()), n, C.c_void_p(po.data_ptr())))
    dt = (time.perf_counter() - t0) / 8
    cur = po.numpy().copy()
    if rref is None: rref = cur
    print(f"registered plan=[{plan or 'default'}]: {dt*1e3:.3f} ms  {n/dt/1e6:.1f} Mpts/s  same={bool((cur == rref).all())} same_as_strict={bool((cur == ref).all())}", flush=True)
