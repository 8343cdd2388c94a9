#!/usr/bin/env python
"""bench.py -- fft1 + power + wide graph + mix1 throughput on B200 (BASELINE.json metric).

A "step" is `passes` passes of the hot path over one batch of synthetic timf1 input per receiver
stream (passes is chosen so that a step is >= 25 ms of GPU work):
    lb200_fft1_dev  (unpack / window / fft1 / filtercorr / |z|^2 -> fft1_sumsq)
    lb200_update_fft1_slowsum_dev + lb200_fft1_waterfall_dev   (the wide-graph consumers)
    lb200_mix1_dev  (bin selection, inverse transform, overlap -> timf3)
then, with more than one stream or GPU, the sum of the streams' averaged power spectra
(lb200_reduce_* : copy-engine push over NVLink + one add kernel on rank 0).
value  : new input samples per second (Msamples/s), whole job, inputs resident in HBM.
e2e    : same metric through the host-buffer C-ABI calls (lb200_fft1 + lb200_mix1) with
         host<->device copies inside the timed region.
roofline: algorithmic bytes of the fft1 kernel / its CUDA-event duration.
per_config: the other BASELINE.json shapes, same measurement, shorter.
cpu_baseline / --impl reference: the reference's own C path (oracle/_ref) on the host cores.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per
GPU, independent receiver streams per rank = weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from linrad_b200 import sizing  # noqa: E402
from linrad_b200.synth import make_timf1  # noqa: E402

IQ, DW, TWO = sizing.IQ_DATA, sizing.DWORD_INPUT, sizing.TWO_CHANNELS

CFG4_SEL = [8192.0 * (1 + c) + 0.25 * c for c in range(16)]
WORKLOADS = {
    # name: (PathSetup kwargs, reference fft_cntrl row, selections (bins), transforms per pass and stream)
    "cfg1": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=13, mix1_red_n=4), 6, [3000.37], 5920),
    "cfg2": (dict(input_mode=IQ | DW | TWO, rf_channels=2, ad_speed=192000, fft1_n=14, mix1_red_n=4), 7, [6000.74], 2960),
    "cfg3": (dict(input_mode=0, rf_channels=1, ad_speed=2400000, fft1_n=15, mix1_red_n=5), 2, [], 740),
    # 240 transforms per call and stream = 1.6 s of a 20 MS/s stream.  The persistent four-step kernel fills and drains
    # its queue over ~8 transforms at either end of a launch: measured 0.155 ms per 60 transforms at 60 per call, 0.130 at
    # 120, 0.124 at 240 (--batch N to try others; round 1 and the first round-2 lines used 60)
    "cfg4": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=18, mix1_red_n=6), 20, CFG4_SEL, 240),
    # configs[4]: 64 independent cfg4 streams on 8 GPUs = 8 streams per GPU (any --gpus N runs 8 per GPU)
    "cfg5": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=18, mix1_red_n=6), 20, CFG4_SEL, 240),
}
STREAMS_PER_GPU = {"cfg5": 8}
WORKLOAD_TEXT = {
    "cfg1": "configs[0]: 1-ch complex IQ 96 kS/s int16, fft1 N=8192 sin^2 window, mix1 M=512 one signal",
    "cfg2": "configs[1]: 2-ch complex IQ 192 kS/s 24-bit (int32), fft1 N=16384 sin^2 window, mix1 M=1024 one signal",
    "cfg3": "configs[2]: real 1-ch int16 2.4 MS/s, fft1_re N=32768 bins (65536 reals), power-spectrum averaging, no mix1",
    "cfg4": "configs[3]: 1-ch complex IQ 20 MS/s int16, fft1 N=262144 four-step, mix1 M=4096 x 16 selections",
    "cfg5": "configs[3] as configs[4] runs it: 8 independent 20 MS/s int16 IQ streams per GPU (64 on 8 GPUs), each fft1 N=262144 "
            "four-step + mix1 M=4096 x 16 selections; sum of the streams' averaged power spectra",
}
DEFAULT_WORKLOAD = "cfg5"
STEP_MS = 25.0

# cfg4's N = 2^18 exceeds every float CPU version of the reference (N <= 65536, buf.c:285-290) and its
# double-precision version 20 is 2-channel only: the CPU arm times the nearest legal size instead
# (version 6, N = 65536, same M = 4096 and 16 selections) and says so.
CPU_OVERRIDE = {
    "cfg4": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=16, mix1_red_n=4), 6,
             "reference float path stops at N=65536: timed at N=65536 (version 6), M=4096, 16 selections"),
}
CPU_OVERRIDE["cfg5"] = CPU_OVERRIDE["cfg4"]


def config_dict(name):
    """the workload as both arms name it (byte-identical in the two JSON lines)"""
    kw, _, selbins, batch = WORKLOADS[name]
    s = sizing.PathSetup(**kw)
    return {"workload": WORKLOAD_TEXT[name], "name": name, "fft1_size": s.fft1_size, "rx_channels": s.rf_channels,
            "mix1_size": s.mix1_size if selbins else 0, "mix1_selections": len(selbins), "fft_avg1num": s.avg1num,
            "streams_per_gpu": STREAMS_PER_GPU.get(name, 1), "parallelism": "independent receiver streams per GPU (weak scaling)"}


def cpu_workload(name):
    kw, version, selbins, _ = WORKLOADS[name]
    note = ""
    if name in CPU_OVERRIDE:
        kw, version, note = CPU_OVERRIDE[name]
        scale = (1 << kw["fft1_n"]) / (1 << WORKLOADS[name][0]["fft1_n"])
        selbins = [b * scale for b in selbins]
    return kw, version, selbins, note


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def samples_per_transform(s):
    """new input samples per transform (real input: 2P real samples)"""
    return s.fft1_new_points * (1 if s.input_mode & IQ else 2)


def kernel_name(s):
    if s.fft1_n > 14:
        k = "fft1_large_pipe_kernel (persistent four-step, one launch per lb200_fft1_dev call)"
        return k if s.input_mode & IQ else k + " + fft1_real_post_kernel"
    if not s.input_mode & IQ:
        return "fft1 real input: packed transform kernel + fft1_real_post_kernel (one lb200_fft1_dev call)"
    return "fft1_fused_kernel" if s.fft1_n >= 10 else "fft1_small_kernel"


def pow2_at_least(x):
    p = 1
    while p < x:
        p *= 2
    return p


def alg_bytes(s, nsel):
    """algorithmic bytes per transform, split by kernel."""
    N, C = s.fft1_size, s.rf_channels
    b_in = s.timf1_blockbytes
    b_fft1 = 8 * C * N
    b_pow = 4.0 * N / s.avg1num
    M, Mi, Mn = s.mix1_size, s.mix1_interleave_points, s.mix1_new_points
    b_mix = nsel * (8 * C * M + 8 * C * Mn + (8 * C * Mi if Mi == Mn else 0))
    return dict(fft1=b_in + b_fft1 + b_pow, mix1=b_mix, total=b_in + b_fft1 + b_pow + b_mix)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
            try:
                pw.append(float(r[6]))
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


# ------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores.
    One process per core, each an independent Linrad-style pipeline on its own block range
    (the reference's own parallel model is independent time blocks per fft1b thread,
    wcw.c:974-1000)."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import refwrap
    kw, version, selbins, note = cpu_workload(args.workload)
    s = sizing.PathSetup(**kw)
    # the compiled reference is mapped by this (parent) process too: the workers are its forks
    parent = refwrap.RefOracle(fft1_version=version, n_sel=len(selbins), **kw)
    raw0 = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, 2, s.fft1_new_points, seed=1)
    parent.process_timed(raw0, 2)
    cores = args.cpu_procs or max(1, (os.cpu_count() or 2))
    blocks = args.cpu_blocks or max(8, int(48 * 8192 * 13 / (s.fft1_size * s.fft1_n * s.rf_channels)))
    ctx = mp.get_context("fork")

    def worker(q_in, q_out, seed):
        from oracle.refwrap import RefOracle
        r = RefOracle(fft1_version=version, n_sel=len(selbins), **kw)
        for i, fb in enumerate(selbins):
            r.set_selfreq(i, s.selfreq_for_bin(fb))
        raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, blocks, s.fft1_new_points, seed=seed)
        q_out.put("ready")
        while True:
            cmd = q_in.get()
            if cmd is None:
                return
            t0 = time.perf_counter()
            r.process_timed(raw, blocks)
            q_out.put(time.perf_counter() - t0)

    procs = []
    for i in range(cores):
        qi, qo = ctx.Queue(), ctx.Queue()
        p = ctx.Process(target=worker, args=(qi, qo, 100 + i), daemon=True)
        p.start()
        procs.append((p, qi, qo))
    for _, _, qo in procs:
        qo.get()
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _, qi, _ in procs:
            qi.put(1)
        for _, _, qo in procs:
            qo.get()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    for p, qi, _ in procs:
        qi.put(None)
    tot = sum(times)
    samples = blocks * samples_per_transform(s) * cores * len(times)
    value = samples / tot / 1e6
    line = {
        "impl": "reference", "metric": "fft1+mix1 IQ Msamples/s", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                         "cpu_model": cpu_model(), "nproc": os.cpu_count(),
                         "algorithmic_GBps": alg_bytes(s, len(selbins))["total"] * value * 1e6 / samples_per_transform(s) / 1e9,
                         "reference_fft1_version": version, "blocks_per_step_per_core": blocks,
                         "sample": f"{blocks} transforms per core per step, {cores} independent pipelines, "
                                   f"fft1_b(v{version})+fft1_c+fft1_waterfall+fft1_mix1_fixed compiled from the reference C files (-O2 -ffast-math)"
                                   + (f"; {note}" if note else "")},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_quick(workload, seconds=12.0):
    """Bounded single-core sample of the compiled reference for the default run."""
    try:
        from oracle import refwrap
        if not refwrap.available():
            return None
        kw, version, selbins, note = cpu_workload(workload)
        s = sizing.PathSetup(**kw)
        r = refwrap.RefOracle(fft1_version=version, n_sel=len(selbins), **kw)
        for i, fb in enumerate(selbins):
            r.set_selfreq(i, s.selfreq_for_bin(fb))
        blocks = 16
        raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, blocks, s.fft1_new_points, seed=9)
        r.process_timed(raw, blocks)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            r.process_timed(raw, blocks)
            n += blocks
        dt = time.perf_counter() - t0
        v = n * samples_per_transform(s) / dt / 1e6
        return {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "reference", "cpu_model": cpu_model(),
                "nproc": os.cpu_count(),
                "algorithmic_GBps": alg_bytes(s, len(selbins))["total"] * v * 1e6 / samples_per_transform(s) / 1e9,
                "sample": f"{n} transforms in {dt:.1f} s, one thread: fft1_b(v{version})+fft1_c+fft1_waterfall+"
                          f"fft1_mix1_fixed from the reference C files (-O2 -ffast-math); nproc={os.cpu_count()}"
                          + (f"; {note}" if note else "")}
    except Exception as e:  # the baseline must never take the GPU line down
        return {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}


def cufft_reference_path(workload, seconds=6.0, batch_n=4):
    """Informational: the reference's OWN GPU support (fft_cntrl row 19 "CUDA", fft1.c:3531-3553: window on the
    CPU, pageable cudaMemcpy in, cufftExecC2C of 2^gpu.fft1_batch_n transforms, cudaMemcpy out, CPU post-processing,
    fft1_c and mix1 on the CPU), compiled from the reference's files with -DHAVE_CUFFT=1 and run on this GPU, one
    thread.  Only 1-channel IQ set-ups have that row (fft1var.c:78); other workloads report the 1-channel case."""
    try:
        from oracle import refwrap
        if not (refwrap.available() and refwrap.cufft_available()):
            return None
        kw, _, selbins, _ = WORKLOADS["cfg4" if workload == "cfg5" else workload]
        note = ""
        if kw["input_mode"] != sizing.IQ_DATA or kw["rf_channels"] != 1:
            kw, _, selbins, _ = WORKLOADS["cfg1"]
            note = "; this workload's input format has no CUDA row in the reference: configs[0] timed instead"
        s = sizing.PathSetup(**kw)
        r = refwrap.RefOracle(fft1_version=19, n_sel=len(selbins), cufft=True, gpu_batch_n=batch_n, **kw)
        for i, fb in enumerate(selbins):
            r.set_selfreq(i, s.selfreq_for_bin(fb))
        calls = 2
        blocks = calls << batch_n
        raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, blocks, s.fft1_new_points, seed=9)
        r.process_timed(raw, calls)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            r.process_timed(raw, calls)
            n += blocks
        dt = time.perf_counter() - t0
        v = n * samples_per_transform(s) / dt / 1e6
        return {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "reference, built with its own -DHAVE_CUFFT=1",
                "fft1_size": s.fft1_size, "mix1_selections": len(selbins), "batch": 1 << batch_n,
                "sample": f"{n} transforms in {dt:.1f} s: fft1_b row 19 (fft1win_gpu + cudaMemcpy + cufftExecC2C x{1 << batch_n} + "
                          f"cudaMemcpy + swap) + fft1_c + fft1_waterfall + fft1_mix1_fixed, one host thread" + note}
    except Exception as e:
        return {"value": None, "unit": "Msamples/s", "sample": f"unavailable: {e}"}


# ------------------------------------------------------------------------------------------
class GpuWorkload:
    """device-resident rings of S receiver streams on one GPU and one pass of the hot path over them"""

    def __init__(self, name, rank, local_rank, world, batch=0):
        import torch
        from linrad_b200 import api
        self.torch, self.api = torch, api
        self.name, self.rank, self.world = name, rank, world
        kw, _, selbins, default_batch = WORKLOADS[name]
        self.s = s = sizing.PathSetup(**kw)
        self.selbins = selbins
        self.B = B = batch or default_batch
        self.S = S = STREAMS_PER_GPU.get(name, 1)
        self.dev = dev = torch.device("cuda", local_rank)
        self.plan = api.Plan(s, device=local_rank)
        self.stream = torch.cuda.ExternalStream(self.plan.stream, device=dev)
        N, C = s.fft1_size, s.rf_channels
        self.timf1_bytes = pow2_at_least((B + 2) * s.timf1_blockbytes)
        self.fft1_floats = pow2_at_least(B * s.fft1_block)
        self.rows = rows = (B + s.avg1num - 1) // s.avg1num
        self.sumsq_floats = pow2_at_least((rows + s.avg2num + 2) * N)
        self.timf3_size = pow2_at_least((B + 2) * s.timf3_block + 2 * C * s.mix1_size)
        self.nsel = nsel = len(selbins)
        # wide graph (fft1.c:4607-4651, wide_graph.c:1374-1389): one pixel per bin, 8 waterfall lines
        wg_first = s.first_xpoint
        wg_last = min(s.first_xpoint + s.xpoints, N - 1)
        xpix = min(s.xpoints, N - s.first_xpoint)
        self.wf_size = xpix * 8
        self.wgc = api.WgConfig(s.avg2num, s.waterfall_avgnum, s.first_xpoint, s.xpoints, wg_first, wg_last, xpix, 1, 1, 100)
        yfac = torch.from_numpy(s.waterfall_yfac(1)).to(dev)
        self.streams = []
        self.host_in = None
        for si in range(S):
            raw = make_timf1(s.input_mode, C, N, 64, s.fft1_new_points, seed=100 + rank * S + si)
            raw_bytes = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
            reps = (B * s.timf1_blockbytes + raw_bytes.size - 1) // raw_bytes.size
            hin = np.tile(raw_bytes, reps)[: B * s.timf1_blockbytes]
            if self.host_in is None:
                self.host_in = hin
            t1 = torch.zeros(self.timf1_bytes, dtype=torch.uint8, device=dev)
            t1[: hin.size].copy_(torch.from_numpy(hin))
            st = dict(
                timf1=t1,
                fft1=torch.empty(self.fft1_floats, dtype=torch.float32, device=dev),
                sumsq=torch.zeros(self.sumsq_floats, dtype=torch.float32, device=dev),
                timf3=torch.zeros(max(nsel, 1) * 2 * self.timf3_size, dtype=torch.float32, device=dev),
                states=api.new_states([s.selfreq_for_bin(b) for b in selbins]),
                slowsum=torch.zeros(N, dtype=torch.float32, device=dev),
                wsum=torch.full((N,), 0.00001, dtype=torch.float32, device=dev),
                yfac=yfac,
                waterf=torch.full((self.wf_size + xpix + 64,), -32768, dtype=torch.int16, device=dev),
                wgstate=api.WgState(0, s.fft1_first_point, 0, 0, 0, 0),
            )
            self.streams.append(st)
        self.specsum = torch.zeros(rows * N, dtype=torch.float32, device=dev) if (S > 1 or world > 1) else None
        self.reducer = None
        torch.cuda.synchronize()

    def one_pass(self, events=None):
        api, s, B = self.api, self.s, self.B
        N = s.fft1_size
        for st in self.streams:
            if events is not None:
                e0 = self.torch.cuda.Event(enable_timing=True)
                e1 = self.torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
            self.plan.fft1_dev(timf1=st["timf1"].data_ptr(), timf1_bytes=self.timf1_bytes, ref=0, nblocks=B,
                               fft1=st["fft1"].data_ptr(), fft1_floats=self.fft1_floats, fft1_pa=0, apply_fc=True,
                               sumsq=st["sumsq"].data_ptr(), sumsq_floats=self.sumsq_floats, sumsq_pa=0, counter=0)
            if events is not None:
                e1.record(self.stream)
                events.append((e0, e1))
            # the wide-graph consumers of the new fft1_sumsq rows (wcw.c:1067-1072: fft1_c, then fft1_waterfall)
            st["wgstate"].fft1_sumsq_pwg = 0
            api.wide_graph_dev(self.plan, self.wgc, st["wgstate"], sumsq=st["sumsq"].data_ptr(), sumsq_floats=self.sumsq_floats,
                               sumsq_pa=0, nrows=B // s.avg1num, slowsum=st["slowsum"].data_ptr(), wsum=st["wsum"].data_ptr(),
                               yfac=st["yfac"].data_ptr(), waterf=st["waterf"].data_ptr(), waterf_size=self.wf_size)
            if self.nsel:
                self.plan.mix1_dev(fft1=st["fft1"].data_ptr(), fft1_floats=self.fft1_floats, fft1_px=0, nblocks=B,
                                   states=st["states"], timf3=st["timf3"].data_ptr(), timf3_floats=self.timf3_size, timf3_pa=0)
        if self.specsum is not None:
            # the averaged power spectrum is the only thing that crosses streams / GPUs
            with self.torch.cuda.stream(self.stream):
                if self.reducer is not None:
                    self.reducer.wait_consumed(self.stream)
                n = self.rows * N
                self.specsum.copy_(self.streams[0]["sumsq"][:n])
                for st in self.streams[1:]:
                    self.specsum.add_(st["sumsq"][:n])
            if self.reducer is not None:
                self.reducer.reduce(self.stream)

    def close(self):
        if self.reducer is not None and hasattr(self.reducer, "close"):
            self.reducer.close()
        self.plan.close()


class NcclReducer:
    """sum of the per-GPU spectra on rank 0 through torch.distributed (NCCL), on its own stream"""

    def __init__(self, wl):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.wl = torch, dist, wl
        self.comm_stream = torch.cuda.Stream(device=wl.dev)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.done.record(self.comm_stream)
        self.kind = "nccl reduce to rank 0 on a side stream"

    def wait_consumed(self, stream):
        stream.wait_event(self.done)

    def reduce(self, stream):
        self.ready.record(stream)
        with self.torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(self.ready)
            self.dist.reduce(self.wl.specsum, dst=0)
            self.done.record(self.comm_stream)

    def finish(self, stream):
        stream.wait_event(self.done)


class P2PReducer:
    """the product's own reduction (lb200_reduce_*): every rank pushes its rows into rank 0's mailbox
    with the copy engine over NVLink, rank 0 adds them with one small kernel; side stream, no NCCL"""

    def __init__(self, wl):
        import torch
        import torch.distributed as dist
        n = wl.rows * wl.s.fft1_size

        def exchange(mine):
            out = [None] * wl.world
            dist.all_gather_object(out, mine)
            return out

        self.wl = wl
        self.r = wl.api.Reducer(wl.plan, wl.rank, wl.world, n, exchange=exchange, root=0, depth=2)
        self.out = torch.zeros(n, dtype=torch.float32, device=wl.dev) if wl.rank == 0 else None
        self.kind = "lb200_reduce_push/sum: copy-engine push over NVLink into rank 0's mailbox + one add kernel, side stream"

    def wait_consumed(self, stream):
        self.r.rows_released()

    def reduce(self, stream):
        self.r.push(self.wl.specsum.data_ptr())
        if self.wl.rank == 0:
            self.r.sum(self.out.data_ptr())

    def finish(self, stream):
        self.r.rows_released()
        if self.wl.rank == 0:
            self.r.result_ready()

    def close(self):
        self.r.synchronize()
        self.r.close()


def make_reducer(kind, wl):
    if wl.world <= 1 or kind == "none":
        return None
    if kind in ("auto", "p2p"):
        try:
            return P2PReducer(wl)
        except Exception as e:          # e.g. IPC not permitted in this container
            if kind == "p2p":
                raise
            sys.stderr.write(f"[bench] lb200_reduce unavailable ({e}); falling back to NCCL\n")
    return NcclReducer(wl)


def measure(wl, steps, warmup, dist, step_ms=STEP_MS, clocks_index=None):
    """timed region: `steps` steps of `passes` passes each; CUDA events on the plan's stream, max over ranks"""
    torch = wl.torch
    world = wl.world
    for _ in range(max(1, warmup // 2)):
        wl.one_pass()
    wl.plan.synchronize()
    # passes per step so that a step is >= step_ms of GPU work (same on every rank)
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    c0.record(wl.stream)
    for _ in range(2):
        wl.one_pass()
    c1.record(wl.stream)
    wl.plan.synchronize()
    t_pass = c0.elapsed_time(c1) / 2
    passes = max(1, int(np.ceil(step_ms / max(t_pass, 1e-3))))
    if world > 1:
        t = torch.tensor([passes], dtype=torch.int64, device=wl.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        passes = int(t.item())

    def step(events=None):
        for i in range(passes):
            wl.one_pass(events if i == 0 else None)

    for _ in range(warmup):
        step()
    wl.plan.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = wl.plan.launches()
    ev_pairs = []
    sampler = ClockSampler(clocks_index) if clocks_index is not None else None
    if sampler:
        sampler.__enter__()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(wl.stream)
    for _ in range(steps):
        step(ev_pairs)
    if wl.reducer is not None:
        wl.reducer.finish(wl.stream)                 # the last reduction belongs to the timed region
    t_end.record(wl.stream)
    wl.plan.synchronize()
    torch.cuda.synchronize()
    if sampler:
        sampler.__exit__()
    if world > 1:
        dist.barrier()
    ms = t_start.elapsed_time(t_end)
    launches = wl.plan.launches() - launches0
    fft1_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=wl.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    s = wl.s
    spt = samples_per_transform(s)
    samples = wl.B * wl.S * spt * passes * steps * world
    value = samples / (ms * 1e-3) / 1e6
    ab = alg_bytes(s, wl.nsel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = ab["fft1"] * wl.B / (fft1_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get("cfg4" if wl.name == "cfg5" else wl.name)
        if tj:
            traffic = tj["dram_bytes_per_transform"] * wl.B
            traffic_src = tj["source"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name(s), "kernel_ms": fft1_ms,
                "algorithmic_bytes_per_launch": ab["fft1"] * wl.B, "transforms_per_launch": wl.B, "peak_source": peak_src,
                "whole_step_frac": ab["total"] * wl.B * wl.S * passes * steps / (ms * 1e-3) / 1e9 / peak}
    return {"value": value, "ms": ms, "ms_per_step": ms / steps, "passes": passes, "launches": launches, "roofline": roofline,
            "clocks": sampler.summary() if sampler else None, "timed_region_s": ms * 1e-3}


def measure_timf2(wl, reps=5):
    """make_timf2 (strong/weak split, back transform, fft1back_fp_finish) over the spectra the
    pass has just left in fft1_float, device resident, timed with events on the plan's stream."""
    torch, api, s = wl.torch, wl.api, wl.s
    if s.fft1_n > 14 or not (s.input_mode & sizing.IQ_DATA):
        return None
    N, C, B = s.fft1_size, s.rf_channels, wl.B
    newp = s.fft1_new_points
    sf = 4 * C
    ring = pow2_at_least(sf * (B * newp + N))
    st = wl.streams[0]
    timf2 = torch.zeros(ring, dtype=torch.float32, device=wl.dev)
    pwr = torch.zeros(ring // sf, dtype=torch.float32, device=wl.dev)
    lim = np.zeros(N, np.float32)
    rng = np.random.default_rng(1)
    for _ in range(12):                                     # a dozen strong carriers, as fft1_update_liminfo leaves them
        a = int(rng.integers(0, N - 40))
        lim[a: a + int(rng.integers(1, 40))] = float(rng.uniform(0.01, 1.0))
    liminfo = torch.from_numpy(lim).to(wl.dev)
    wl.one_pass()
    times = []
    for i in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(wl.stream)
        api.make_timf2_dev(wl.plan, fft1=st["fft1"].data_ptr(), fft1_floats=wl.fft1_floats, fft1_px=0, nblocks=B,
                           liminfo=liminfo.data_ptr(), timf2=timf2.data_ptr(), timf2_floats=ring, timf2_pwr=pwr.data_ptr(), timf2_pa=0)
        e1.record(wl.stream)
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    alg = B * (8 * C * N + 16 * C * newp + 4 * newp)       # spectra in, [weak, strong] samples and |weak|^2 out
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    out = {"what": "lb200_make_timf2_dev (timf2_back_kernel + timf2_finish_kernel), transforms of the pass, device resident",
           "transforms": B, "ms": ms, "algorithmic_bytes": alg, "achieved_GBps": alg / (ms * 1e-3) / 1e9,
           "frac": alg / (ms * 1e-3) / 1e9 / peak, "Msamples_per_s": B * samples_per_transform(s) / (ms * 1e-3) / 1e6}
    del timf2, pwr
    return out


def measure_fft3(wl, reps=5):
    """the transforms of make_fft3_all over the baseband the pass has just left in timf3 (a second
    plan with float IQ input on that ring): fft3_size = mix1.size, sin^2 window, 50 % overlap, every selection."""
    torch, api, s = wl.torch, wl.api, wl.s
    if not wl.nsel or not (s.input_mode & sizing.IQ_DATA) or s.mix1_n < 7 or s.mix1_n > 14:
        return None
    C, B = s.rf_channels, wl.B
    two = sizing.TWO_CHANNELS if C == 2 else 0
    s3 = sizing.PathSetup(input_mode=sizing.IQ_DATA | sizing.DWORD_INPUT | sizing.FLOAT_INPUT | two, rf_channels=C, ad_speed=96000,
                          fft1_n=s.mix1_n, mix1_red_n=3, sinpow=2)
    N3, newp = s3.fft1_size, s3.fft1_new_points
    per_sel = (B * s.timf3_block // (2 * C) - N3) // newp + 1          # whole transforms inside what one pass wrote
    if per_sel < 1:
        return None
    plan3 = api.Plan(s3, device=wl.dev.index)
    st = wl.streams[0]
    out_floats = pow2_at_least(wl.nsel * per_sel * s3.fft1_block)
    out = torch.empty(out_floats, dtype=torch.float32, device=wl.dev)
    stream3 = torch.cuda.ExternalStream(plan3.stream, device=wl.dev)
    wl.one_pass()
    torch.cuda.synchronize()
    times = []
    for i in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream3)
        # one call: the selections are rings 2*timf3_size floats apart (fft3.c:232), their blocks side by side in the output ring
        plan3.fft1_dev(timf1=st["timf3"].data_ptr(), timf1_bytes=wl.timf3_size * 4, ref=s3.fft1_interleave_points * s3.frame_bytes,
                       nblocks=per_sel, fft1=out.data_ptr(), fft1_floats=out_floats, fft1_pa=0, apply_fc=False,
                       rings=wl.nsel, ring_stride=2 * wl.timf3_size * 4, pa_stride=per_sel * s3.fft1_block)
        e1.record(stream3)
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    ntr = per_sel * wl.nsel
    alg = ntr * (8 * C * newp + 8 * C * N3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    res = {"what": "transform half of make_fft3_all (lb200_fft1_dev on a float-input plan over timf3, fft1_small_kernel), "
                   "every selection of the pass, device resident",
           "fft3_size": N3, "transforms": ntr, "calls": 1, "rings": wl.nsel, "ms": ms, "algorithmic_bytes": alg,
           "achieved_GBps": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak}
    plan3.close()
    del out
    return res


def run_e2e(wl, args, dist):
    """end to end through the host-buffer C ABI: the call a Linrad-side host makes"""
    torch, api, s = wl.torch, wl.api, wl.s
    world = wl.world
    N, C = s.fft1_size, s.rf_channels
    nsel = wl.nsel
    spt = samples_per_transform(s)
    Be = args.e2e_batch or 16 * max(1, (16 << 20) // (4 * s.fft1_block))
    Be = min(Be, wl.B)
    h_timf1 = torch.zeros(pow2_at_least((Be + 2) * s.timf1_blockbytes), dtype=torch.uint8).pin_memory()
    h_timf1[: Be * s.timf1_blockbytes].copy_(torch.from_numpy(wl.host_in[: Be * s.timf1_blockbytes]))
    h_fft1 = torch.zeros(pow2_at_least(Be * s.fft1_block), dtype=torch.float32).pin_memory()
    h_sumsq = torch.zeros(pow2_at_least((Be // s.avg1num + 2) * N), dtype=torch.float32).pin_memory()
    t3s = pow2_at_least((Be + 2) * s.timf3_block + 2 * C * s.mix1_size)
    h_timf3 = torch.zeros(max(nsel, 1) * 2 * t3s, dtype=torch.float32).pin_memory()
    plan2 = api.Plan(s, device=wl.dev.index)
    st2 = api.new_states([s.selfreq_for_bin(b) for b in wl.selbins])
    os.environ["LB200_NO_HOSTREGISTER"] = "1"      # buffers are already pinned

    def e2e_step(keep=False):
        plan2.fft1_host(timf1=h_timf1.numpy(), ref=0, nblocks=Be, fft1=h_fft1.numpy(), fft1_pa=0, apply_fc=True,
                        sumsq=h_sumsq.numpy(), sumsq_pa=0, counter=0, keep_on_device=keep)
        if nsel:
            plan2.mix1_host(fft1=h_fft1.numpy(), fft1_px=0, nblocks=Be, states=st2, timf3=h_timf3.numpy(),
                            timf3_floats=t3s, timf3_pa=0)

    def e2e_run(keep):
        for _ in range(3):
            e2e_step(keep)
        h0, d0 = plan2.h2d_bytes(), plan2.d2h_bytes()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        nrep = max(3, args.steps // 2)
        for _ in range(nrep):
            e2e_step(keep)
        plan2.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=wl.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return {"value": Be * spt * nrep * world / dt / 1e6, "unit": "Msamples/s",
                "h2d_bytes_per_step": (plan2.h2d_bytes() - h0) // nrep, "d2h_bytes_per_step": (plan2.d2h_bytes() - d0) // nrep,
                "batch": Be, "streams": 1}

    # the drop-in call: everything the reference's fft1_b / fft1_c / fft1_mix1_fixed leave in host
    # memory comes back (fft1_float, fft1_sumsq, timf3)
    e2e = e2e_run(False)
    e2e["api"] = "lb200_fft1 + lb200_mix1 on pinned host rings (one receiver stream per GPU)"
    if nsel:
        lazy = e2e_run(True)
        lazy["api"] = "same, fft1_float kept in the device mirror (LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE); fft1_sumsq and timf3 come back"
        e2e["spectrum_on_device"] = lazy

    # what the host interface itself can carry with every GPU of the run copying at the same time:
    # bare pinned-memory copies, both directions at once (e2e above is bound by this, not by the kernels)
    try:
        nbytes = 256 << 20
        hp_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        hp_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dv_in = torch.empty(nbytes, dtype=torch.uint8, device=wl.dev)
        dv_out = torch.empty(nbytes, dtype=torch.uint8, device=wl.dev)
        s_a, s_b = torch.cuda.Stream(device=wl.dev), torch.cuda.Stream(device=wl.dev)

        def both(reps):
            for _ in range(reps):
                with torch.cuda.stream(s_a):
                    dv_in.copy_(hp_in, non_blocking=True)
                with torch.cuda.stream(s_b):
                    hp_out.copy_(dv_out, non_blocking=True)
            s_a.synchronize()
            s_b.synchronize()

        both(1)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        both(4)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=wl.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        gbs = 4 * nbytes / dt / 1e9
        bytes_per_sample = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / (Be * spt)
        # the two directions run side by side: the busier one bounds the step
        busier = max(e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"]) / (Be * spt)
        e2e["pcie_probe"] = {"concurrent_gpus": world, "GBps_per_gpu_each_direction": gbs, "GBps_all_gpus_both_directions": 2 * gbs * world,
                             "e2e_bytes_per_sample": bytes_per_sample, "e2e_bytes_per_sample_busier_direction": busier,
                             "e2e_ceiling_Msamples_per_s": gbs * 1e9 * world / busier / 1e6,
                             "e2e_GBps_busier_direction_per_gpu": e2e["value"] * 1e6 * busier / world / 1e9,
                             "what": "256 MiB pinned copies H2D and D2H at the same time on every GPU of the run, slowest rank; "
                                     "ceiling = that rate / bytes per sample of the busier direction (D2H: fft1_float comes back)"}
        del hp_in, hp_out, dv_in, dv_out
    except Exception as ex:
        e2e["pcie_probe"] = {"error": str(ex)}

    # Linrad-sized calls: one transform per call, as the shim issues them (real-time use)
    def one_block(i):
        plan2.fft1_host(timf1=h_timf1.numpy(), ref=(i % Be) * s.timf1_blockbytes, nblocks=1, fft1=h_fft1.numpy(),
                        fft1_pa=(i % Be) * s.fft1_block, apply_fc=True, sumsq=h_sumsq.numpy(), sumsq_pa=0, counter=0)
        if nsel:
            plan2.mix1_host(fft1=h_fft1.numpy(), fft1_px=(i % Be) * s.fft1_block, nblocks=1, states=st2,
                            timf3=h_timf3.numpy(), timf3_floats=t3s, timf3_pa=0)
    for i in range(10):
        one_block(i)
    plan2.synchronize()
    t0 = time.perf_counter()
    for i in range(50):
        one_block(10 + i)
    plan2.synchronize()
    lat = (time.perf_counter() - t0) / 50
    e2e["single_block_call_us"] = lat * 1e6
    e2e["single_block_realtime_margin"] = (spt / s.ad_speed) / lat      # sample time of one block / time to process it
    plan2.close()
    return e2e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="transforms per pass per stream (0 = workload default)")
    ap.add_argument("--step-ms", type=float, default=STEP_MS, help="minimum GPU time of one step (passes are repeated)")
    ap.add_argument("--e2e-batch", type=int, default=0, help="transforms per host-ring call (0 = sixteen 16 MB sub-batches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl", "p2p", "none"])
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--cpu-blocks", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    # ONE JSON line on stdout: libraries that print to file descriptor 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = GpuWorkload(args.workload, rank, local_rank, world, batch=args.batch)
    reduce_kind = "none (one GPU)"
    wl.reducer = make_reducer(args.reduce, wl)
    if wl.reducer is not None:
        reduce_kind = wl.reducer.kind
    m = measure(wl, args.steps, args.warmup, dist, step_ms=args.step_ms, clocks_index=local_rank)
    e2e = None if args.no_e2e else run_e2e(wl, args, dist)
    wl.close()
    del wl
    torch.cuda.empty_cache()

    per_config = None
    if not args.no_per_config:
        per_config = {}
        for name in ("cfg1", "cfg2", "cfg3", "cfg4"):
            w2 = GpuWorkload(name, rank, local_rank, world)
            w2.reducer = make_reducer(args.reduce, w2)
            r = measure(w2, max(3, args.steps // 4), 3, dist, step_ms=args.step_ms)
            per_config[name] = {"workload": WORKLOAD_TEXT[name], "value": r["value"], "unit": "Msamples/s",
                                "ms_per_step": r["ms_per_step"], "passes_per_step": r["passes"],
                                "transforms_per_pass_per_gpu": w2.B * w2.S,
                                "frac": r["roofline"]["frac"], "kernel_ms": r["roofline"]["kernel_ms"],
                                "kernel": r["roofline"]["kernel"], "achieved_GBps": r["roofline"]["achieved"],
                                "traffic": r["roofline"]["traffic"], "whole_step_frac": r["roofline"]["whole_step_frac"]}
            try:
                t2 = measure_timf2(w2)
                if t2:
                    per_config[name]["make_timf2"] = t2
            except Exception as ex:                        # a widened row must not take the headline down
                per_config[name]["make_timf2"] = {"error": str(ex)}
            try:
                t3 = measure_fft3(w2)
                if t3:
                    per_config[name]["make_fft3"] = t3
            except Exception as ex:
                per_config[name]["make_fft3"] = {"error": str(ex)}
            w2.close()
            del w2
            torch.cuda.empty_cache()

    cpu = None
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_quick(args.workload)
        ref_gpu = cufft_reference_path(args.workload)

    if rank == 0:
        kw, _, selbins, default_batch = WORKLOADS[args.workload]
        s = sizing.PathSetup(**kw)
        S = STREAMS_PER_GPU.get(args.workload, 1)
        B = args.batch or default_batch
        line = {
            "metric": "fft1+mix1 IQ Msamples/s", "value": m["value"], "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload),
            "detail": {"transforms_per_pass_per_stream": B, "passes_per_step": m["passes"], "timed_region_s": m["timed_region_s"],
                       "step": "fft1 (+fft1_c power) -> slowsum + waterfall -> mix1 per stream, then the sum of the streams' spectra",
                       "l2": f"working set per pass {(S * B * (s.timf1_blockbytes + 4 * s.fft1_block)) >> 20} MiB > 126 MiB L2, no flush needed",
                       "spectrum_reduction": reduce_kind},
            "roofline": m["roofline"], "cpu_baseline": cpu, "cufft_reference_path": ref_gpu, "e2e": e2e, "gpu_launches": m["launches"],
            "clocks": m["clocks"], "per_config": per_config,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
