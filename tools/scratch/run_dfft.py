"""Run the d_fft device pieces (client fft1 + king pipeline) a few times at m = 2^lg (profiling target)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
l = 2; m = 1 << lg; mbyl = m // l
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = capi.ctx_p(); capi.check(lib.zkg_ctx_create(0, C.c_void_p(st.cuda_stream), C.byref(ctx)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
def rnd(k):
    t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device="cuda", generator=g); t[:, 3] &= (1 << 61) - 1; return t
dom = z.Radix2EvaluationDomain.new(m)
gen = dom.group_gen(); gcos = z.Radix2EvaluationDomain.new(2 * m).element(1)
px, shares, rd = rnd(mbyl), rnd(8 * mbyl), rnd(2 * mbyl)
outp = torch.empty((8 * mbyl, 4), dtype=torch.int64, device="cuda")
for _ in range(reps):
    capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(px.data_ptr()), mbyl, l, gen.ctypes.data, None, None))
    capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, 8, mbyl, l, gen.ctypes.data, gcos.ctypes.data, 1,
                                           C.c_void_p(rd.data_ptr()), C.c_void_p(outp.data_ptr())))
capi.check(lib.zkg_ctx_sync(ctx))
print("done", lg)
