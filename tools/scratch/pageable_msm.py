"""Host-pointer MSM with pageable (numpy) vs pinned (torch) buffers."""
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
ha = a.cpu().numpy().view(np.uint64).copy()
out = np.zeros(12, dtype=np.uint64)
pb = torch.from_numpy(hb).pin_memory(); pa = torch.from_numpy(ha.view(np.int64)).pin_memory()
po = torch.zeros(12, dtype=torch.int64).pin_memory()
def run(bp, ap, op, label):
    for _ in range(2):
        capi.check(lib.zkg_msm_bn254_g1(0, C.c_void_p(bp), 72, n, C.c_void_p(ap), n, C.c_void_p(op)))
    t0 = time.perf_counter()
    for _ in range(5):
        capi.check(lib.zkg_msm_bn254_g1(0, C.c_void_p(bp), 72, n, C.c_void_p(ap), n, C.c_void_p(op)))
    dt = (time.perf_counter() - t0) / 5
    print(f"{label}: {dt*1e3:.2f} ms  {n/dt/1e6:.1f} Mpts/s", flush=True)
run(pb.data_ptr(), pa.data_ptr(), po.data_ptr(), "pinned")
run(hb.ctypes.data, ha.ctypes.data, out.ctypes.data, "pageable")
print("agree", bool((po.numpy().view(np.uint64) == out).all()))
h = C.c_uint64(0)
capi.check(lib.zkg_bases_register(0, 1, C.c_void_p(hb.ctypes.data), 72, n, C.byref(h)))
def runr(ap, op, label):
    for _ in range(2):
        capi.check(lib.zkg_msm_bn254_registered(h.value, C.c_void_p(ap), n, C.c_void_p(op)))
    t0 = time.perf_counter()
    for _ in range(5):
        capi.check(lib.zkg_msm_bn254_registered(h.value, C.c_void_p(ap), n, C.c_void_p(op)))
    dt = (time.perf_counter() - t0) / 5
    print(f"registered {label}: {dt*1e3:.2f} ms  {n/dt/1e6:.1f} Mpts/s", flush=True)
runr(pa.data_ptr(), po.data_ptr(), "pinned")
runr(ha.ctypes.data, out.ctypes.data, "pageable")
