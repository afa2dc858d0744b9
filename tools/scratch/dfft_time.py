"""Time the d_fft device pieces (client fft1, king stage 1, king pack) at m = 2^lg; CUDA events + phase marks."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import zksaas_b200 as z
from zksaas_b200 import capi
lib = z.lib()
lg = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
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
lib.zkg_ctx_set_profiling(ctx, 1)
def fft1(): capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(px.data_ptr()), mbyl, l, gen.ctypes.data, None, None))
def king(gg, re): capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, 8, mbyl, l, gen.ctypes.data, gg.ctypes.data, re,
                                                    C.c_void_p(rd.data_ptr()), C.c_void_p(outp.data_ptr())))
one = z.Radix2EvaluationDomain.new(m).element(0)
def timed(f):
    for _ in range(3): f()
    capi.check(lib.zkg_ctx_sync(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): f()
    e1.record(st); e1.synchronize()
    return e0.elapsed_time(e1) / reps
def phases():
    ph = []
    for k in range(4):
        f = C.c_float(0); lib.zkg_ctx_phase_ms(ctx, k, C.byref(f)); ph.append(round(f.value, 4))
    return ph
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ZKG_"))
t = timed(fft1); print(f"[{tag}] m=2^{lg} fft1 {t:.4f} ms phases {phases()}", flush=True)
t = timed(lambda: king(gcos, 1)); print(f"[{tag}] m=2^{lg} king(g=coset,rearrange) {t:.4f} ms phases {phases()}  chk={outp[12345 % (8*mbyl)].cpu().numpy()[:2]}", flush=True)
t = timed(lambda: king(one, 0)); print(f"[{tag}] m=2^{lg} king(g=1,consecutive) {t:.4f} ms phases {phases()}  chk={outp[12345 % (8*mbyl)].cpu().numpy()[:2]}", flush=True)
