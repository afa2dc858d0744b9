#!/usr/bin/env python3
"""bench.py -- headline benchmark of the zk-SaaS hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 22]

Metric (BASELINE.json): BN254 G1 MSM Mpts/s.  Workload: the per-party local MSM of d_msm
(dist-primitives/src/dmsm/mod.rs:73) at 2^22 points per GPU ("Large BN254 G1 d_msm 2^22-2^24",
BASELINE.json configs[3]); synthetic uniform scalars and bases with random discrete logs.  A step is
one MSM.  With N GPUs every rank owns its own 2^22-point range of one N*2^22-point MSM (weak
scaling); the partial sums are exchanged with one NCCL all-gather and added on the device.

One JSON line is printed by rank 0; see README/DESIGN.md for the keys.  The d_fft leg (BASELINE
configs[1], m = 2^16) and a 2^20-constraint d_fft are reported under "secondary".

--impl reference times the CPU restatement of the reference's arkworks path (oracle/, kind "port":
the Rust reference cannot be built in this image) on the host cores, same metric and unit.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INT_PEAK_TIMAD = 18.51              # profiles/r01_intpipe_microbench_v2.json: 32-bit IMAD issue rate, B200, 148 SMs @1965 MHz
WIDE_MAD_PEAK_T = 9.27              # profiles/r01_widemad_microbench.json: IMAD.WIDE.U32 (any form: RZ / addend / .X) issue rate, 10^12/s
WIDE_MADS_PER_MADD = 6 * 128 + 2 * 100 + 192   # XYZZ mixed add as executed: 6 products, 2 dedicated squarings, one 2-term dot
ACC_TRAFFIC_BYTES = 7.336e9         # profiles/r01_ncu_k_accumulate_v2.json: dram read+write of one k_accumulate launch at 2^22 (prepared path)
HBM_PEAK_FALLBACK_GBS = 6650.0      # B200_PROFILING.md fallback


def measured_peaks():
    hbm, how = HBM_PEAK_FALLBACK_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm, how = float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        pass
    imad, ihow = INT_PEAK_TIMAD, "measured (tools/microbench/intpipe.cu -> profiles/r01_intpipe_microbench.json)"
    return hbm, how, imad, ihow


def ark_window(k):
    """arkworks' own rule, used only to state the ALGORITHMIC work per point."""
    lg = (k - 1).bit_length()
    c = 3 if k < 32 else lg * 69 // 100 + 2
    return c, (254 + c - 1) // c


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, reasons, smax, pmax = [], set(), None, 0.0
        for r in self.rows:
            try:
                clocks.append(float(r[1])); smax = float(r[2]); pmax = max(pmax, float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        clocks.sort()
        # "under load": upper half of the samples (the sampler also sees the idle gaps between steps)
        load = clocks[len(clocks) // 2:] if clocks else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "power_w_max": pmax, "samples": len(clocks), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the arkworks path) -- reported beside the GPU number, never the target
# ---------------------------------------------------------------------------------------------
def cpu_msm_sample(log2n, reps, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol
    from oracle_lib import _p
    lib = ol.oracle()
    th = threads or min(lib.zko_max_threads(), os.cpu_count() or 1)
    n = 1 << log2n
    rng = np.random.default_rng(0x7A6B)
    bases = np.zeros((n, 72), dtype=np.uint8)
    lib.zko_g1_sequence(_p(ol.rand_fr(rng, 1)), _p(ol.rand_fr(rng, 1)), n, bases.ctypes.data, 72)
    scalars = ol.rand_fr(rng, n)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ol.o_g1_msm(bases, scalars, threads=th)
        times.append(time.perf_counter() - t0)
    return n, th, times


def cpu_dfft_sample(log2m):
    """client fft1 + king closure of one d_fft on the host (oracle literal loops, single thread: what
    dist-primitives runs -- it has no `parallel` feature)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import numpy as np
    import oracle_lib as ol
    from oracle_lib import _p, pyref
    o = ol.oracle()
    l, m = 2, 1 << log2m
    mbyl, n = m // l, 8
    rng = np.random.default_rng(5)
    gen = ol.fr_np([pyref.Radix2Domain(m).group_gen])
    g = ol.fr_np([pyref.Radix2Domain(2 * m).element(1)])
    px = ol.rand_fr(rng, mbyl)
    t0 = time.perf_counter()
    o.zko_fft1_in_place(_p(px), mbyl, l, _p(gen))
    t_fft1 = time.perf_counter() - t0
    shares = [ol.rand_fr(rng, mbyl) for _ in range(n)]
    outs = [np.zeros((mbyl, 4), dtype=np.uint64) for _ in range(n)]
    rand = ol.rand_fr(rng, mbyl * l)
    par = (C.c_uint32 * n)(*range(n))
    t0 = time.perf_counter()
    o.zko_king_fft2(ol.ptr_array(shares), par, n, mbyl, l, _p(gen), _p(g), 1, _p(rand), ol.ptr_array(outs))
    t_king = time.perf_counter() - t0
    return {"fft1_ms": round(t_fft1 * 1e3, 3), "king_ms": round(t_king * 1e3, 3),
            "d_fft_elems_per_s": round(m / (t_fft1 + t_king), 1), "cores": 1, "kind": "port"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log2n = min(args.log2n, 20)
    n, th, times = cpu_msm_sample(log2n, args.warmup + args.steps)
    timed = times[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    val = n / (ms * 1e-3) / 1e6
    c, W = ark_window(n)
    sample = f"G1 MSM of 2^{log2n} points per step (arkworks window rule c={c}, W={W}), {th} OpenMP threads over windows"
    line = {
        "impl": "reference", "metric": "BN254 G1 MSM Mpts/s", "value": round(val, 4), "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic",
        "config": {"workload": f"d_msm local G1 MSM, BN254, 2^{args.log2n} points per GPU (bounded CPU sample 2^{log2n})",
                   "curve": "BN254 G1", "l": 2},
        "cpu_baseline": {"value": round(val, 4), "unit": "Mpts/s", "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = C restatement of ark-ec 0.4.2 msm_bigint_wnaf (oracle/zkoracle.c); the Rust reference "
                "cannot be built here (no cargo/rustc, arkworks crates un-vendored)",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import zksaas_b200 as z
    from zksaas_b200 import capi
    from zksaas_b200.api import fr_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = None
    if world > 1 and not os.environ.get("ZKG_BENCH_NO_AFFINITY"):
        # one process per GPU: keep the rank (and the pinned host buffers it first-touches) on the CPUs next to its GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local)
            bus = "%08x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
            ncpu = os.cpu_count() or 1
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
                affinity = len(cpus)
        except Exception:
            affinity = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = z.lib()
    # One explicit (non-default) stream for everything: the library launches on it, torch's events
    # and NCCL collectives are recorded on it, so CUDA-event timings see the kernels they bracket.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.ctx_p()
    capi.check(lib.zkg_ctx_create(local, C.c_void_p(stream.cuda_stream), C.byref(ctx)))

    n = 1 << args.log2n
    # ---- synthetic inputs, generated on the device: uniform scalars, bases = s_i * G ---------------
    g = torch.Generator(device=dev)
    g.manual_seed(0x7A6B53616153 ^ (3 + rank))

    def rand_fr_dev(k):
        # uniform below 2^253 (< r): 253-bit uniform values used directly as Montgomery images
        t = torch.randint(-2**63, 2**63 - 1, (k, 4), dtype=torch.int64, device=dev, generator=g)
        t[:, 3] &= (1 << 61) - 1
        return t

    scalars = rand_fr_dev(n)
    dlogs = rand_fr_dev(n)
    bases = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(dlogs.data_ptr()), n, C.c_void_p(bases.data_ptr())))
    torch.cuda.synchronize()
    del dlogs
    # The CRS shares are static across proofs: register them once (window-shifted table in HBM).  The
    # device-resident `value` leg runs against the handle; the e2e leg below does NOT (it ships the
    # arkworks base images over PCIe every step, as the unmodified d_msm signature would).
    handle = C.c_uint64(0)
    t_reg0 = time.perf_counter()
    capi.check(lib.zkg_bases_register_dev(ctx, 1, C.c_void_p(bases.data_ptr()), n, C.byref(handle)))
    torch.cuda.synchronize()
    register_s = time.perf_counter() - t_reg0
    partial = torch.zeros(16, dtype=torch.int64, device=dev)           # XYZZ, 128 B
    gathered = torch.zeros(16 * world, dtype=torch.int64, device=dev)
    out_xyz = torch.zeros(12, dtype=torch.int64, device=dev)

    def step_device():
        if world == 1:
            capi.check(lib.zkg_msm_bn254_registered_dev(ctx, handle.value, C.c_void_p(scalars.data_ptr()), n,
                                                        C.c_void_p(out_xyz.data_ptr()), 0))
        else:
            capi.check(lib.zkg_msm_bn254_registered_dev(ctx, handle.value, C.c_void_p(scalars.data_ptr()), n,
                                                        C.c_void_p(partial.data_ptr()), 1))
            dist.all_gather_into_tensor(gathered, partial)
            capi.check(lib.zkg_msm_combine_dev(ctx, 1, C.c_void_p(gathered.data_ptr()), world,
                                                C.c_void_p(out_xyz.data_ptr())))

    def step_device_unprepared():
        capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(scalars.data_ptr()), n,
                                             C.c_void_p(out_xyz.data_ptr())))

    def barrier(collective=True):
        if world > 1 and collective:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collective=True):
        """K steps bracketed by barrier + synchronize; device time via CUDA events; max over ranks.
        collective=False: rank-local timing (the secondary legs run on rank 0 only)."""
        barrier(collective)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier(collective)
        ms = e0.elapsed_time(e1)
        if world > 1 and collective:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident leg ("value") ----------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    launches0 = C.c_uint64(0)
    lib.zkg_ctx_launch_count(ctx, C.byref(launches0))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_device, args.steps)
    launches1 = C.c_uint64(0)
    lib.zkg_ctx_launch_count(ctx, C.byref(launches1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e6
    dev_result = out_xyz.cpu().numpy().copy()

    # ---- same workload with two MSMs in flight (two contexts / streams): the latency-bound tail of one
    #      (bucket reduction) overlaps the bucket accumulation of the next -- how a prover that issues its
    #      d_msm calls concurrently (prove.rs:227 try_join!) drives the library.  Reported beside `value`.
    two_stream_ms = None
    if world == 1:
        st2 = torch.cuda.Stream(device=dev)
        ctx2 = capi.ctx_p()
        capi.check(lib.zkg_ctx_create(local, C.c_void_p(st2.cuda_stream), C.byref(ctx2)))
        out2 = torch.zeros(12, dtype=torch.int64, device=dev)
        st2.wait_stream(stream)

        def run_pair_steps(k):
            for i in range(k):
                if i % 2 == 0:
                    capi.check(lib.zkg_msm_bn254_registered_dev(ctx, handle.value, C.c_void_p(scalars.data_ptr()), n, C.c_void_p(out_xyz.data_ptr()), 0))
                else:
                    capi.check(lib.zkg_msm_bn254_registered_dev(ctx2, handle.value, C.c_void_p(scalars.data_ptr()), n, C.c_void_p(out2.data_ptr()), 0))
            stream.wait_stream(st2)
        run_pair_steps(4)
        torch.cuda.synchronize()
        k2 = max(4, args.steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st2.wait_stream(stream)
        e0.record()
        run_pair_steps(k2)
        e1.record()
        torch.cuda.synchronize()
        two_stream_ms = e0.elapsed_time(e1) / k2
        two_stream_same = bool((out2.cpu().numpy() == dev_result).all())
        lib.zkg_ctx_destroy(ctx2)

    # ---- dominant kernel (bucket accumulation) timed live with CUDA events on the launch stream -------
    capi.check(lib.zkg_ctx_set_profiling(ctx, 1))
    phase = [[], [], []]
    for _ in range(max(3, min(args.steps, 10))):
        capi.check(lib.zkg_msm_bn254_registered_dev(ctx, handle.value, C.c_void_p(scalars.data_ptr()), n,
                                                    C.c_void_p(partial.data_ptr()), 1))
        for ph in range(3):
            f = C.c_float(0)
            capi.check(lib.zkg_ctx_phase_ms(ctx, ph, C.byref(f)))
            phase[ph].append(f.value)
    capi.check(lib.zkg_ctx_set_profiling(ctx, 0))
    acc_ms = sum(phase[1]) / len(phase[1])
    sort_ms = sum(phase[0]) / len(phase[0])
    red_ms = sum(phase[2]) / len(phase[2])
    # the same MSM without the prepared table (generic path: per-window buckets + Horner), for reference
    unprepared_ms = None
    if world == 1:
        for _ in range(2):
            step_device_unprepared()
        unprepared_ms = timed(step_device_unprepared, max(3, min(args.steps, 10)), collective=False) / max(3, min(args.steps, 10))
        unprepared_same = bool((out_xyz.cpu().numpy() == dev_result).all())

    # ---- end-to-end leg: the reference-facing C-ABI call with HOST buffers ------------------------------
    # arkworks Affine images (72 B/point) + Fr images in pinned host memory; every step copies them in.
    h_bases = torch.zeros((n, 72), dtype=torch.uint8).pin_memory()
    h_bases[:, :64] = bases.cpu()
    h_scal = scalars.cpu().pin_memory()
    h_out = torch.zeros(12, dtype=torch.int64).pin_memory()
    xyzz_dev = torch.zeros(16, dtype=torch.int64, device=dev)
    one_fq = torch.tensor([x - (1 << 64) if x >= (1 << 63) else x
                           for x in (0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f)],
                          dtype=torch.int64)
    e2e_result = {}

    def step_e2e():
        capi.check(lib.zkg_msm_bn254_g1(local, C.c_void_p(h_bases.data_ptr()), 72, n, C.c_void_p(h_scal.data_ptr()), n,
                                        C.c_void_p(h_out.data_ptr())))
        if world > 1:
            # the result is back on the host; combine the ranks' points with one more tiny exchange
            xyzz = torch.zeros(16, dtype=torch.int64)
            if bool((h_out[8:12] != 0).any()):
                xyzz[0:8] = h_out[0:8]; xyzz[8:12] = one_fq; xyzz[12:16] = one_fq
            xyzz_dev.copy_(xyzz)
            dist.all_gather_into_tensor(gathered, xyzz_dev)
            capi.check(lib.zkg_msm_combine_dev(ctx, 1, C.c_void_p(gathered.data_ptr()), world, C.c_void_p(out_xyz.data_ptr())))
            e2e_result["xyz"] = out_xyz.cpu().numpy().copy()
        else:
            e2e_result["xyz"] = h_out.numpy().copy()

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = n * world * e2e_steps / e2e_s / 1e6
    same = bool((e2e_result["xyz"] == dev_result).all())
    # e2e against the registered handle: only the scalars cross PCIe each step
    e2e_reg_ms = None
    if world == 1:
        for _ in range(2):
            capi.check(lib.zkg_msm_bn254_registered(handle.value, C.c_void_p(h_scal.data_ptr()), n, C.c_void_p(h_out.data_ptr())))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            capi.check(lib.zkg_msm_bn254_registered(handle.value, C.c_void_p(h_scal.data_ptr()), n, C.c_void_p(h_out.data_ptr())))
        e2e_reg_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        same = same and bool((h_out.numpy() == dev_result).all())
    # the same two calls with PAGEABLE host buffers (a Rust Vec<F> as the reference would pass it): the
    # library stages them through its own pinned slots with parallel memcpy (csrc/staging.cu)
    e2e_pageable = None
    if world == 1:
        pg_bases, pg_scal = h_bases.numpy().copy(), h_scal.numpy().copy()
        pg_out = np.zeros(12, dtype=np.int64)
        res = {}
        for name, call in (("unregistered", lambda: lib.zkg_msm_bn254_g1(local, C.c_void_p(pg_bases.ctypes.data), 72, n,
                                                                          C.c_void_p(pg_scal.ctypes.data), n, C.c_void_p(pg_out.ctypes.data))),
                           ("registered", lambda: lib.zkg_msm_bn254_registered(handle.value, C.c_void_p(pg_scal.ctypes.data), n,
                                                                                C.c_void_p(pg_out.ctypes.data)))):
            for _ in range(2):
                capi.check(call())
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                capi.check(call())
            ms = (time.perf_counter() - t0) / e2e_steps * 1e3
            res[name + "_ms_per_step"] = round(ms, 3)
            res[name + "_Mpts_per_s"] = round(n / ms / 1e3, 2)
            same = same and bool((pg_out.view(dev_result.dtype) == dev_result).all())
        e2e_pageable = res
        del pg_bases, pg_scal

    # ---- secondary: d_fft pieces (configs[1]: m = 2^16; and the 2^20-constraint size), rank 0 only -------
    secondary = {}
    if rank == 0 and not args.no_secondary:
        pp_l = 2
        for lg in (16, 20, 24):
            m = 1 << lg
            mbyl = m // pp_l
            dom = z.Radix2EvaluationDomain.new(m)
            px = rand_fr_dev(mbyl)
            shares = rand_fr_dev(8 * mbyl)
            rnd = rand_fr_dev(2 * mbyl)
            outp = torch.empty((8 * mbyl, 4), dtype=torch.int64, device=dev)
            gen = dom.group_gen()
            gcos = z.Radix2EvaluationDomain.new(2 * m).element(1)

            def f1():
                capi.check(lib.zkg_fft1_bn254_dev(ctx, C.c_void_p(px.data_ptr()), mbyl, pp_l, gen.ctypes.data, None, None))

            def fk():
                capi.check(lib.zkg_king_fft2_bn254_dev(ctx, C.c_void_p(shares.data_ptr()), None, 8, mbyl, pp_l,
                                                        gen.ctypes.data, gcos.ctypes.data, 1, C.c_void_p(rnd.data_ptr()),
                                                        C.c_void_p(outp.data_ptr())))
            for fn in (f1, fk):
                for _ in range(3):
                    fn()
            t1 = timed(f1, 10, collective=False) / 10
            tk = timed(fk, 10, collective=False) / 10
            if lg > 20:                      # top of the north_star range: device-resident figures only
                secondary[f"d_fft_m2^{lg}"] = {
                    "fft1_ms": round(t1, 4), "king_ms": round(tk, 4),
                    "d_fft_elems_per_s": round(m / ((t1 + tk) * 1e-3), 1),
                    "king_hbm_gbs": round((256 + 32) * m / (tk * 1e-3) / 1e9, 1),
                    "fft1_hbm_gbs": round(64 * mbyl / (t1 * 1e-3) / 1e9, 1),
                }
                del px, shares, rnd, outp
                continue
            # e2e of the king call with host buffers (the reference-facing entry point): pageable numpy
            # arrays (what a Rust Vec<F> is) and pinned buffers (the contract's e2e convention)
            hs = [np.ascontiguousarray(shares.cpu().numpy().view(np.uint64).reshape(8, mbyl, 4)[p]) for p in range(8)]
            hr = rnd.cpu().numpy().view(np.uint64)
            pp = z.PackedSharingParams.new(pp_l, device=local)
            z.king_fft2(hs, list(range(8)), pp, gen, gcos, True, hr)
            t0 = time.perf_counter()
            for _ in range(3):
                z.king_fft2(hs, list(range(8)), pp, gen, gcos, True, hr)
            tke = (time.perf_counter() - t0) / 3
            pin_in = torch.empty((8, mbyl, 4), dtype=torch.int64).pin_memory()
            pin_in.copy_(shares.reshape(8, mbyl, 4))
            pin_rnd = torch.empty((2 * mbyl, 4), dtype=torch.int64).pin_memory()
            pin_rnd.copy_(rnd)
            pin_out = torch.empty((8, mbyl, 4), dtype=torch.int64).pin_memory()
            u64p = C.POINTER(C.c_uint64)
            in_arr = (u64p * 8)(*[C.cast(pin_in[p].data_ptr(), u64p) for p in range(8)])
            out_arr = (u64p * 8)(*[C.cast(pin_out[p].data_ptr(), u64p) for p in range(8)])

            def fke():
                capi.check(lib.zkg_king_fft2_bn254(local, in_arr, None, 8, mbyl, pp_l, gen.ctypes.data, gcos.ctypes.data, 1,
                                                   C.c_void_p(pin_rnd.data_ptr()), out_arr))
            fke()
            t0 = time.perf_counter()
            for _ in range(3):
                fke()
            tkp = (time.perf_counter() - t0) / 3
            pinned_ok = bool((pin_out.numpy().view(np.uint64)[3] == z.king_fft2(hs, list(range(8)), pp, gen, gcos, True, hr)[3]).all())
            secondary[f"d_fft_m2^{lg}"] = {
                "fft1_ms": round(t1, 4), "king_ms": round(tk, 4),
                "d_fft_elems_per_s": round(m / ((t1 + tk) * 1e-3), 1),
                "king_e2e_host_ms": round(tke * 1e3, 3), "king_e2e_pinned_ms": round(tkp * 1e3, 3),
                "king_e2e_pinned_elems_per_s": round(m / tkp, 1), "king_e2e_paths_agree": pinned_ok,
                "king_hbm_gbs": round((256 + 32) * m / (tk * 1e-3) / 1e9, 1),
                "fft1_hbm_gbs": round(64 * mbyl / (t1 * 1e-3) / 1e9, 1),
            }
            del px, shares, rnd, outp

    # ---- secondary: G2 MSM at the config-5 size (V query: 2^19 points), rank 0 ----------------------------
    if rank == 0 and not args.no_secondary:
        n2 = 1 << 19
        a2, s2 = rand_fr_dev(n2), rand_fr_dev(n2)
        b2 = torch.empty((n2, 128), dtype=torch.uint8, device=dev)
        o2 = torch.zeros(24, dtype=torch.int64, device=dev)
        capi.check(lib.zkg_fixed_base_dev(ctx, 2, C.c_void_p(s2.data_ptr()), n2, C.c_void_p(b2.data_ptr())))

        def fg2():
            capi.check(lib.zkg_msm_bn254_g2_dev(ctx, C.c_void_p(b2.data_ptr()), C.c_void_p(a2.data_ptr()), n2, C.c_void_p(o2.data_ptr())))
        for _ in range(2):
            fg2()
        tg2 = timed(fg2, 5, collective=False) / 5
        secondary["msm_g2_2^19"] = {"ms": round(tg2, 3), "Mpts_per_s": round(n2 / (tg2 * 1e-3) / 1e6, 2)}
        n1 = 1 << 20
        a1 = rand_fr_dev(n1)

        def fg1():
            capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(bases.data_ptr()), C.c_void_p(a1.data_ptr()), n1, C.c_void_p(out_xyz.data_ptr())))
        for _ in range(2):
            fg1()
        tg1 = timed(fg1, 5, collective=False) / 5
        # bottom of the north_star range with registered (prepared) bases: the first 2^20 points of the bench set
        h20 = C.c_uint64(0)
        capi.check(lib.zkg_bases_register_dev(ctx, 1, C.c_void_p(bases.data_ptr()), n1, C.byref(h20)))
        o20 = torch.zeros(12, dtype=torch.int64, device=dev)

        def fr20():
            capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h20.value, C.c_void_p(a1.data_ptr()), n1, C.c_void_p(o20.data_ptr()), 0))
        for _ in range(2):
            fr20()
        tr20 = timed(fr20, 5, collective=False) / 5
        secondary["msm_g1_2^20"] = {"ms": round(tg1, 3), "Mpts_per_s": round(n1 / (tg1 * 1e-3) / 1e6, 2),
                                    "registered_ms": round(tr20, 3), "registered_Mpts_per_s": round(n1 / (tr20 * 1e-3) / 1e6, 2),
                                    "paths_agree": bool((o20 == out_xyz).all())}
        capi.check(lib.zkg_bases_release(h20.value))
        del a2, s2, b2, a1
        # top of the north_star range: 2^24 points on one GPU (generic path and registered bases)
        n24 = 1 << 24
        a24, s24 = rand_fr_dev(n24), rand_fr_dev(n24)
        b24 = torch.empty((n24, 64), dtype=torch.uint8, device=dev)
        capi.check(lib.zkg_fixed_base_dev(ctx, 1, C.c_void_p(s24.data_ptr()), n24, C.c_void_p(b24.data_ptr())))
        del s24
        h24 = C.c_uint64(0)
        capi.check(lib.zkg_bases_register_dev(ctx, 1, C.c_void_p(b24.data_ptr()), n24, C.byref(h24)))
        o24 = torch.zeros((2, 12), dtype=torch.int64, device=dev)

        def fg24():
            capi.check(lib.zkg_msm_bn254_g1_dev(ctx, C.c_void_p(b24.data_ptr()), C.c_void_p(a24.data_ptr()), n24, C.c_void_p(o24[0].data_ptr())))

        def fr24():
            capi.check(lib.zkg_msm_bn254_registered_dev(ctx, h24.value, C.c_void_p(a24.data_ptr()), n24, C.c_void_p(o24[1].data_ptr()), 0))
        fg24(); fr24()
        tg24 = timed(fg24, 3, collective=False) / 3
        tr24 = timed(fr24, 3, collective=False) / 3
        secondary["msm_g1_2^24"] = {"generic_ms": round(tg24, 3), "generic_Mpts_per_s": round(n24 / (tg24 * 1e-3) / 1e6, 2),
                                    "registered_ms": round(tr24, 3), "registered_Mpts_per_s": round(n24 / (tr24 * 1e-3) / 1e6, 2),
                                    "paths_agree": bool((o24[0] == o24[1]).all())}
        capi.check(lib.zkg_bases_release(h24.value))
        del a24, b24

    # ---- secondary: ONE king pipeline (m = 2^20) sharded over all ranks: stage 1 -> reduce-scatter -> stage 2 ----
    if world > 1 and not args.no_secondary:
        from zksaas_b200 import sharding
        l_, mbyl_ = 2, 1 << 19
        dom_ = z.Radix2EvaluationDomain.new(mbyl_ * l_)
        gen_, g_ = dom_.group_gen(), z.Radix2EvaluationDomain.new(2 * mbyl_ * l_).element(1)
        lo_, hi_ = sharding.shard_range(mbyl_, world, rank)
        loc_ = rand_fr_dev(8 * (hi_ - lo_)).reshape(8, hi_ - lo_, 4)
        rl_ = rand_fr_dev((hi_ - lo_) * 2)

        def fks():
            sharding.king_fft2_sharded_cuda(ctx, lib, torch, dist, loc_, mbyl_, l_, gen_, g_, True, rl_, rank, world)
        for _ in range(3):
            fks()
        tks = timed(fks, 10) / 10
        if rank == 0:
            secondary["king_sharded_m2^20"] = {"ms": round(tks, 4), "ranks": world,
                                               "elems_per_s": round(mbyl_ * l_ / (tks * 1e-3), 1),
                                               "collective": "one NCCL reduce_scatter (sum) of the pack-order buffer"}

    # ---- secondary: ONE fft1 lane (m = 2^24, l = 2) sharded over all ranks: inner NTT -> all-to-all -> outer DFT ----
    if world > 1 and not args.no_secondary and (world & (world - 1)) == 0:
        from zksaas_b200 import sharding
        l_, mbyl_ = 2, 1 << 23
        gen_ = z.Radix2EvaluationDomain.new(mbyl_ * l_).group_gen()
        blk_ = rand_fr_dev(mbyl_ // world)

        def ffs():
            sharding.fft1_sharded_cuda(ctx, lib, torch, dist, blk_, mbyl_, l_, gen_, rank, world)
        for _ in range(3):
            ffs()
        tfs = timed(ffs, 10) / 10
        if rank == 0:
            secondary["fft1_sharded_m2^24"] = {"ms": round(tfs, 4), "ranks": world,
                                               "share_elems_per_s": round(mbyl_ / (tfs * 1e-3), 1),
                                               "collective": "one NCCL all_to_all_single of the twiddled inner transforms (m/l x 32 B in total)"}
        del blk_

    # ---- secondary: the SURVEY 8(f) rows built so far, through their host-pointer entry points (PCIe included) ----
    if rank == 0 and not args.no_secondary:
        from zksaas_b200 import api as zapi
        rngn = np.random.default_rng(9)

        def np_fr(k):
            a = rngn.integers(0, 2**64, size=(k, 4), dtype=np.uint64)
            a[:, 3] &= np.uint64((1 << 61) - 1)
            return a
        widened = {}
        pp2 = z.PackedSharingParams.new(2, device=local)
        cols = 1 << 17                                              # d_pp over m = 2^18 secrets
        shares_pp = [np_fr(2 * cols) for _ in range(8)]
        rnd_pp = np_fr(cols * 2)
        z.dpp_king(shares_pp, list(range(8)), pp2, rnd_pp)
        t0 = time.perf_counter()
        z.dpp_king(shares_pp, list(range(8)), pp2, rnd_pp)
        t_dpp = time.perf_counter() - t0
        widened["dpp_king_m2^18_host_ms"] = round(t_dpp * 1e3, 3)
        ncrs = 1 << 14                                              # CRS query of 2^14 G1 points -> 8 x 2^13 share points
        crs = np.zeros((ncrs, 72), dtype=np.uint8)
        crs[:, :64] = bases[:ncrs].cpu().numpy()
        zapi.crs_det_pack(crs, pp2)
        t0 = time.perf_counter()
        zapi.crs_det_pack(crs, pp2)
        t_crs = time.perf_counter() - t0
        widened["crs_det_pack_g1_2^14_host_ms"] = round(t_crs * 1e3, 3)
        widened["crs_det_pack_g1_points_per_s"] = round(ncrs / t_crs, 1)
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            from oracle_lib import _p
            o = ol.oracle()
            ccols = 1 << 12
            outs_ = [np.zeros((ccols, 4), dtype=np.uint64) for _ in range(8)]
            sh_ = [np.ascontiguousarray(np.concatenate([x[:ccols], x[cols:cols + ccols]])) for x in shares_pp]
            par_ = (C.c_uint32 * 8)(*range(8))
            t0 = time.perf_counter()
            o.zko_dpp_king(ol.ptr_array(sh_), par_, 8, ccols, 2, _p(rnd_pp), ol.ptr_array(outs_))
            widened["cpu_dpp_king_ms_scaled_to_m2^18"] = round((time.perf_counter() - t0) * 1e3 * cols / ccols, 1)
            sec_ = np.zeros((2, 12), dtype=np.uint64)
            one_q = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f], dtype=np.uint64)
            for k_ in range(2):
                sec_[k_, 0:8] = crs[k_, :64].view(np.uint64)
                sec_[k_, 8:12] = one_q
            exp_ = np.zeros(8 * 12, dtype=np.uint64)
            t0 = time.perf_counter()
            for _ in range(8):
                o.zko_pss_pack_g1(2, _p(sec_.reshape(-1)), None, _p(exp_))
            widened["cpu_crs_det_pack_g1_points_per_s"] = round(16 / (time.perf_counter() - t0), 1)
        secondary["survey_8f_rows"] = widened

    # ---- secondary: emulated distributed Groth16 prove (BASELINE configs[4]: 2^20 constraints, l = 2, n = 8 parties) ----
    # Dataflow of groth16/examples/sha256.rs:32-129 with synthetic CRS shares (PackedProvingKeyShare::rand sizes,
    # groth16/src/proving_key.rs:125-176): per party circom_h = 3 d_ifft + 3 d_fft + deg_red (ext_wit.rs:104-181) then
    # 5 d_msm (prove.rs:52,106,154,209,219).  Parties are dealt round-robin to the ranks; the king closures run on rank 0.
    # Device-resident, network excluded; correctness of this dataflow is pinned in tests/test_gpu_protocol.py.
    if not args.no_secondary and not args.no_prove:
        lg_m, l_ = 20, 2
        m_ = 1 << lg_m
        mbyl_ = m_ // l_
        dom_ = z.Radix2EvaluationDomain.new(m_)
        gen_i, gen_f, sinv = dom_.group_gen_inv(), dom_.group_gen(), dom_.size_inv()
        zeta = z.Radix2EvaluationDomain.new(2 * m_).element(1)
        one_ = fr_image(1)
        my_parties = [p for p in range(8) if p % world == rank]
        hs = {}
        for name, grp, cnt in (("S", 1, mbyl_), ("H", 1, mbyl_), ("W", 1, mbyl_), ("U", 1, 2 * mbyl_), ("V", 2, mbyl_)):
            bb = torch.empty((cnt, 64 * grp), dtype=torch.uint8, device=dev)
            capi.check(lib.zkg_fixed_base_dev(ctx, grp, C.c_void_p(rand_fr_dev(cnt).data_ptr()), cnt, C.c_void_p(bb.data_ptr())))
            hh = C.c_uint64(0)
            capi.check(lib.zkg_bases_register_dev(ctx, grp, C.c_void_p(bb.data_ptr()), cnt, C.byref(hh)))
            torch.cuda.synchronize()
            hs[name] = (hh, grp, cnt)
            del bb
        va, vb, vc = rand_fr_dev(mbyl_), rand_fr_dev(mbyl_), rand_fr_dev(mbyl_)
        mask = rand_fr_dev(mbyl_)
        hbuf = rand_fr_dev(2 * mbyl_)
        kin = rand_fr_dev(8 * mbyl_)
        kout = torch.empty((8 * mbyl_, 4), dtype=torch.int64, device=dev)
        krand = rand_fr_dev(2 * mbyl_)
        P = lambda t: C.c_void_p(t.data_ptr())
        # The 5 MSMs of a party (and the parties themselves) are independent: in the reference they are
        # concurrent tokio tasks (prove.rs:227 try_join!, multi.rs:320-325), i.e. concurrent C-ABI calls on
        # pooled contexts.  Here: 4 contexts on 4 streams, so one MSM's latency-bound tail overlaps the
        # bucket accumulation of the next.
        n_lanes = 4
        lanes = []
        for _i in range(n_lanes):
            st_ = torch.cuda.Stream(device=dev)
            cx_ = capi.ctx_p()
            capi.check(lib.zkg_ctx_create(local, C.c_void_p(st_.cuda_stream), C.byref(cx_)))
            lanes.append((st_, cx_, torch.zeros(24, dtype=torch.int64, device=dev)))

        def prove_round():
            for _p in my_parties:                                   # clients: 3 x d_ifft first halves
                for v in (va, vb, vc):
                    capi.check(lib.zkg_fft1_bn254_dev(ctx, P(v), mbyl_, l_, gen_i.ctypes.data, sinv.ctypes.data, P(mask)))
            if rank == 0:                                           # king: 3 x (unpack2, fft2, powers, rearranged pack)
                for _ in range(3):
                    capi.check(lib.zkg_king_fft2_bn254_dev(ctx, P(kin), None, 8, mbyl_, l_, gen_i.ctypes.data, zeta.ctypes.data, 1, P(krand), P(kout)))
            for _p in my_parties:                                   # clients: 3 x d_fft first halves
                for v in (va, vb, vc):
                    capi.check(lib.zkg_fft1_bn254_dev(ctx, P(v), mbyl_, l_, gen_f.ctypes.data, None, P(mask)))
            if rank == 0:
                for _ in range(3):
                    capi.check(lib.zkg_king_fft2_bn254_dev(ctx, P(kin), None, 8, mbyl_, l_, gen_f.ctypes.data, one_.ctypes.data, 0, P(krand), P(kout)))
            for _p in my_parties:                                   # h = a*b - c on shares, then deg_red
                capi.check(lib.zkg_field_op_dev(ctx, 0, 0, P(va), P(vb), P(hbuf), mbyl_))
                capi.check(lib.zkg_field_op_dev(ctx, 0, 2, P(hbuf), P(vc), P(hbuf), mbyl_))
            if rank == 0:
                capi.check(lib.zkg_deg_red_king_bn254_dev(ctx, P(kin), None, 8, mbyl_, l_, P(krand), P(kout)))
            for st_, _, _ in lanes:                                 # 5 d_msm local MSMs against the registered CRS shares
                st_.wait_stream(stream)
            k_ = 0
            for _p in my_parties:
                for name, sc in (("V", va), ("U", hbuf), ("S", va), ("H", va), ("W", vb)):
                    hh, grp, cnt = hs[name]
                    st_, cx_, mo_ = lanes[k_ % n_lanes]
                    k_ += 1
                    capi.check(lib.zkg_msm_bn254_registered_dev(cx_, hh.value, P(sc), cnt, P(mo_), 0))
            for st_, _, _ in lanes:
                stream.wait_stream(st_)
        prove_round()
        t_prove = timed(prove_round, 3) / 3
        for hh, _, _ in hs.values():
            lib.zkg_bases_release(hh.value)
        for _, cx_, _ in lanes:
            lib.zkg_ctx_destroy(cx_)
        if rank == 0:
            secondary["groth16_prove_emulated_m2^20"] = {
                "prove_sec": round(t_prove * 1e-3, 5), "parties": 8, "ranks": world, "l": 2,
                "msm_streams": n_lanes,
                "what": "6 client fft1 + 5 registered MSMs (S,H,W: 2^19 G1; U: 2^20 G1; V: 2^19 G2) per party, 6 king d_fft "
                        "closures + 1 deg_red king on rank 0; device-resident, network and pairing check excluded"}
        del va, vb, vc, mask, hbuf, kin, kout, krand

    if rank == 0:
        hbm_peak, hbm_how, int_peak, int_how = measured_peaks()
        c_ark, W_ark = ark_window(n)
        imad_per_point = 11 * W_ark * 272                        # 11*W modmul x 272 IMAD-class instructions
        achieved = n * imad_per_point / (acc_ms * 1e-3) / 1e12
        my_c = int(os.environ.get("ZKG_MSM_PREP_C", "0")) or {19: 20, 20: 20, 21: 20, 22: 20, 23: 20, 24: 20}.get(args.log2n, 20)
        my_W = 254 // my_c + 1
        wide_rate = n * my_W * WIDE_MADS_PER_MADD / (acc_ms * 1e-3) / 1e12   # 32x32->64 multiply-adds the kernel actually executes
        if not args.no_cpu and not args.no_secondary:
            secondary["cpu_d_fft_m2^16"] = cpu_dfft_sample(16)
        cpu_n, cpu_th, cpu_t = cpu_msm_sample(18, 3) if not args.no_cpu else (0, 0, [1.0])
        cpu_val = cpu_n / min(cpu_t) / 1e6
        line = {
            "metric": "BN254 G1 MSM Mpts/s", "value": round(value, 2), "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit Montgomery limbs, IMAD.WIDE)", "data": "synthetic",
            "config": {"workload": f"d_msm local G1 MSM (dist-primitives/src/dmsm/mod.rs:73), BN254, 2^{args.log2n} points per GPU",
                       "points_per_gpu": n, "total_points": n * world, "curve": "BN254 G1", "l": 2,
                       "bases": "registered once (zkg_bases_register_dev: window-shifted table, W copies per base, "
                                f"{register_s:.2f} s one-time); the e2e leg re-ships unregistered bases every step",
                       "window_bits": my_c, "windows": my_W,
                       "l2_policy": "inputs larger than L2 (bases 256 MiB + scalars 128 MiB per step at 2^22)",
                       "multi_gpu": "point-range sharding; one NCCL all-gather of 128 B partial sums + device add",
                       "rank_cpu_affinity": affinity},
            "clocks": clocks,
            "e2e": {"value": round(e2e_val, 2), "unit": "Mpts/s", "h2d_bytes_per_step": n * (72 + 32),
                    "d2h_bytes_per_step": 96, "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3),
                    "call": "zkg_msm_bn254_g1 (host pointers, pinned; arkworks 72-B affine images + Fr images)",
                    "matches_device_leg": same,
                    "registered_bases_ms_per_step": round(e2e_reg_ms, 3) if e2e_reg_ms else None,
                    "registered_bases_Mpts_per_s": round(n / (e2e_reg_ms * 1e-3) / 1e6, 2) if e2e_reg_ms else None,
                    "pageable_host_buffers": e2e_pageable},
            "value_two_streams": {"Mpts_per_s": round(n / (two_stream_ms * 1e-3) / 1e6, 2), "ms_per_step": round(two_stream_ms, 4),
                                  "what": "same K registered MSMs alternating over two contexts/streams (tail of one overlaps "
                                          "accumulation of the next)", "matches": two_stream_same} if two_stream_ms else None,
            "value_unprepared": {"Mpts_per_s": round(n / (unprepared_ms * 1e-3) / 1e6, 2), "ms_per_step": round(unprepared_ms, 4),
                                 "what": "same MSM through zkg_msm_bn254_g1_dev (no prepared table)",
                                 "matches": unprepared_same} if unprepared_ms else None,
            "gpu_launches": int(launches1.value - launches0.value),
            "roofline": {"bound": "int (fmaheavy integer multiply-add pipe; not hbm, not tensor)", "kernel": "k_accumulate<Fq>",
                         "achieved": round(achieved, 3), "peak": int_peak, "unit": "TIMAD/s",
                         "frac": round(achieved / int_peak, 4), "peak_source": int_how,
                         "algorithmic_imad_per_point": imad_per_point,
                         "note": "SURVEY 8(d) formula k*11*W*272/T with arkworks' W; frac exceeds 1 because the kernel does less "
                                 "work than the formula charges: 13 windows instead of 15 (prepared table), and an XYZZ mixed "
                                 "add of 1160 wide multiply-adds (6 products, 2 squarings of 100, one 2-term inner product) "
                                 "where arkworks' Jacobian madd is 11 products of 128",
                         "executed_wide_mad_T_per_s": round(wide_rate, 3), "wide_mad_peak_T_per_s": WIDE_MAD_PEAK_T,
                         "wide_mad_frac": round(wide_rate / WIDE_MAD_PEAK_T, 4),
                         "wide_mad_note": "the binding unit: IMAD.WIDE.U32 issues at half the 32-bit IMAD rate in every form "
                                          "(tools/microbench/widemad.cu), so a 254-bit Montgomery product is 128 of them at best",
                         "kernel_ms": round(acc_ms, 4),
                         "kernel_share_of_step": round(acc_ms / ms_per_step, 4),
                         "phase_ms": {"digits_sort": round(sort_ms, 4), "accumulate": round(acc_ms, 4),
                                      "reduce_final": round(red_ms, 4),
                                      "note": "CUDA events around the phases of one un-pipelined call; the two launch-heavy "
                                              "phases include host launch latency (visible when N ranks share the host cores)"},
                         "hbm": {"achieved_gbs": round(96 * n / (acc_ms * 1e-3) / 1e9, 1), "peak_gbs": hbm_peak,
                                 "frac": round(96 * n / (acc_ms * 1e-3) / 1e9 / hbm_peak, 4), "peak_source": hbm_how},
                         "traffic": ACC_TRAFFIC_BYTES if args.log2n == 22 else None,
                         "traffic_note": "ncu dram bytes per launch; 18x the 96 B/point because Pippenger gathers every base once per window (13 gathers that each pull 128 B) -- 0.92 TB/s, 14% of HBM peak, not the bound"},
            "cpu_baseline": {"value": round(cpu_val, 4), "unit": "Mpts/s", "cores": cpu_th, "kind": "port",
                             "sample": "G1 MSM of 2^18 points, best of 3, oracle/zkoracle.c (arkworks msm_bigint_wnaf restated), OpenMP over windows"},
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    lib.zkg_ctx_destroy(ctx)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=22)
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
