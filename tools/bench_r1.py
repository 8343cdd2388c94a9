#!/usr/bin/env python
"""bench.py -- fft1 + power + mix1 throughput on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic timf1 input:
    lb200_fft1_dev (fused unpack/window/fft1/filtercorr/|z|^2/fft1_sumsq) + lb200_mix1_dev.
value  : new input samples per second (Msamples/s), whole job, inputs resident in HBM.
e2e    : same metric through the host-buffer C-ABI calls (lb200_fft1 + lb200_mix1) with
         host<->device copies inside the timed region.
roofline: algorithmic bytes of the fft1 kernel / its CUDA-event duration.
cpu_baseline / --impl reference: the reference's own C path (oracle/_ref) on the host cores.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per
GPU, independent receiver streams per rank = weak scaling; the only collective is the
all-reduce of the averaged power spectrum).
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

WORKLOADS = {
    # name: (PathSetup kwargs, reference fft_cntrl row, selections (bins), batch per step)
    "cfg1": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=13, mix1_red_n=4), 6, [3000.37], 5920),
    "cfg2": (dict(input_mode=IQ | DW | TWO, rf_channels=2, ad_speed=192000, fft1_n=14, mix1_red_n=4), 7, [6000.74], 2960),
    "cfg3": (dict(input_mode=0, rf_channels=1, ad_speed=2400000, fft1_n=15, mix1_red_n=5), 2, [], 740),
    "cfg4": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=18, mix1_red_n=6), 20,
             [8192.0 * (1 + c) + 0.25 * c for c in range(16)], 60),
    # configs[4]: 64 independent cfg4 streams on 8 GPUs = 8 streams per GPU (any --gpus N runs 8 per GPU)
    "cfg5": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=18, mix1_red_n=6), 20,
             [8192.0 * (1 + c) + 0.25 * c for c in range(16)], 60),
}
STREAMS_PER_GPU = {"cfg5": 8}
WORKLOAD_TEXT = {
    "cfg1": "configs[0]: 1-ch complex IQ 96 kS/s int16, fft1 N=8192 sin^2 window, mix1 M=512 one signal",
    "cfg2": "configs[1]: 2-ch complex IQ 192 kS/s 24-bit (int32), fft1 N=16384 sin^2 window, mix1 M=1024 one signal",
    "cfg3": "configs[2]: real 1-ch int16 2.4 MS/s, fft1_re N=32768 bins (65536 reals), power-spectrum averaging, no mix1",
    "cfg4": "configs[3]: 1-ch complex IQ 20 MS/s int16, fft1 N=262144 four-step, mix1 M=4096 x 16 selections",
    "cfg5": "configs[4]: 8 independent 20 MS/s IQ streams per GPU (64 on 8 GPUs), each as configs[3]; per-GPU sum + all-reduce of the averaged power spectra",
}


# cfg4's N = 2^18 exceeds every float CPU version of the reference (N <= 65536, buf.c:285-290) and its
# double-precision version 20 is 2-channel only: the CPU arm times the nearest legal size instead
# (version 6, N = 65536, same M = 4096 and 16 selections) and says so.
CPU_OVERRIDE = {
    "cfg4": (dict(input_mode=IQ, rf_channels=1, ad_speed=20000000, fft1_n=16, mix1_red_n=4), 6,
             "reference float path stops at N=65536: timed at N=65536 (version 6), M=4096, 16 selections"),
}
CPU_OVERRIDE["cfg5"] = CPU_OVERRIDE["cfg4"]


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
    if not s.input_mode & IQ:
        return "fft1 real input: packed transform kernel(s) + fft1_real_post_kernel (one lb200_fft1_dev call)"
    if s.fft1_n > 14:
        return "fft1_large_cols_kernel + fft1_large_rows_kernel (four-step, one lb200_fft1_dev call)"
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
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores.
    One process per core, each an independent Linrad-style pipeline on its own block range
    (the reference's own parallel model is independent time blocks per fft1b thread,
    wcw.c:974-1000)."""
    if rank != 0:
        return
    import multiprocessing as mp
    kw, version, selbins, note = cpu_workload(args.workload)
    s = sizing.PathSetup(**kw)
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
        "config": {"workload": WORKLOAD_TEXT[args.workload], "blocks_per_step_per_core": blocks,
                   "reference_fft1_version": version},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                         "cpu_model": cpu_model(), "nproc": os.cpu_count(),
                         "algorithmic_GBps": alg_bytes(s, len(selbins))["total"] * value * 1e6 / samples_per_transform(s) / 1e9,
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


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="transforms per step per GPU (0 = workload default)")
    ap.add_argument("--e2e-batch", type=int, default=0, help="transforms per host-ring call (0 = sixteen 16 MB sub-batches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--cpu-blocks", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from linrad_b200 import api

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kw, version, selbins, default_batch = WORKLOADS[args.workload]
    s = sizing.PathSetup(**kw)
    B = args.batch or default_batch
    N, C = s.fft1_size, s.rf_channels
    plan = api.Plan(s, device=local_rank)
    stream = torch.cuda.ExternalStream(plan.stream, device=dev)

    # ---- device-resident rings (same layouts as Linrad's host rings), one set per receiver stream
    S = STREAMS_PER_GPU.get(args.workload, 1)
    timf1_bytes = pow2_at_least((B + 2) * s.timf1_blockbytes)
    fft1_floats = pow2_at_least(B * s.fft1_block)
    rows = (B + s.avg1num - 1) // s.avg1num
    sumsq_floats = pow2_at_least((rows + 1) * N)
    timf3_size = pow2_at_least((B + 2) * s.timf3_block + 2 * C * s.mix1_size)
    nsel = len(selbins)
    d_timf1, d_fft1, d_sumsq, d_timf3, states = [], [], [], [], []
    host_in = None
    for si in range(S):
        raw = make_timf1(s.input_mode, C, N, 64, s.fft1_new_points, seed=100 + rank * S + si)
        raw_bytes = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
        reps = (B * s.timf1_blockbytes + raw_bytes.size - 1) // raw_bytes.size
        hin = np.tile(raw_bytes, reps)[: B * s.timf1_blockbytes]
        if host_in is None:
            host_in = hin
        t1 = torch.zeros(timf1_bytes, dtype=torch.uint8, device=dev)
        t1[: hin.size].copy_(torch.from_numpy(hin))
        d_timf1.append(t1)
        d_fft1.append(torch.empty(fft1_floats, dtype=torch.float32, device=dev))
        d_sumsq.append(torch.zeros(sumsq_floats, dtype=torch.float32, device=dev))
        d_timf3.append(torch.zeros(max(nsel, 1) * 2 * timf3_size, dtype=torch.float32, device=dev))
        states.append(api.new_states([s.selfreq_for_bin(b) for b in selbins]))
    d_specsum = torch.zeros(rows * N, dtype=torch.float32, device=dev) if (S > 1 or world > 1) else None
    torch.cuda.synchronize()

    ev_pairs = []

    def step(record=False):
        # block 0 of the batch starts at byte 0; its overlap half is the ring's tail (zeros/old data)
        for si in range(S):
            if record:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            plan.fft1_dev(timf1=d_timf1[si].data_ptr(), timf1_bytes=timf1_bytes, ref=0, nblocks=B, fft1=d_fft1[si].data_ptr(),
                          fft1_floats=fft1_floats, fft1_pa=0, apply_fc=True, sumsq=d_sumsq[si].data_ptr(),
                          sumsq_floats=sumsq_floats, sumsq_pa=0, counter=0)
            if record:
                e1.record(stream)
                ev_pairs.append((e0, e1))
            if nsel:
                plan.mix1_dev(fft1=d_fft1[si].data_ptr(), fft1_floats=fft1_floats, fft1_px=0, nblocks=B, states=states[si],
                              timf3=d_timf3[si].data_ptr(), timf3_floats=timf3_size, timf3_pa=0)
        if d_specsum is not None:
            # the averaged power spectrum is the only thing that crosses GPUs:
            # per-GPU sum over its streams, then one reduction to rank 0 (the instance that
            # draws the wide graph), on its own stream so that the next batch's kernels do not wait
            with torch.cuda.stream(stream):
                if world > 1:
                    stream.wait_event(comm_done)          # the previous reduction is done with d_specsum
                d_specsum.copy_(d_sumsq[0][: rows * N])
                for si in range(1, S):
                    d_specsum.add_(d_sumsq[si][: rows * N])
                if world > 1:
                    spec_ready.record(stream)
            if world > 1:
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(spec_ready)
                    dist.reduce(d_specsum, dst=0)
                    comm_done.record(comm_stream)

    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    spec_ready = torch.cuda.Event()
    comm_done = torch.cuda.Event()
    if world > 1:
        comm_done.record(comm_stream)

    for _ in range(args.warmup):
        step()
    plan.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = plan.launches()
    with ClockSampler(local_rank) as clk:
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        for _ in range(args.steps):
            step(record=True)
        if world > 1:
            stream.wait_event(comm_done)                  # the last reduction belongs to the timed region
        t_end.record(stream)
        plan.synchronize()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t_start.elapsed_time(t_end)
    launches = plan.launches() - launches0
    fft1_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    spt = samples_per_transform(s)
    samples = B * S * spt * args.steps * world
    value = samples / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (fft1_small_kernel) -----------------------------------
    ab = alg_bytes(s, nsel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = ab["fft1"] * B / (fft1_ms * 1e-3) / 1e9
    # DRAM bytes of the same launch from the committed `ncu --set full` capture (per transform,
    # scaled to this launch's batch); null when no capture exists for the workload
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get("cfg4" if args.workload == "cfg5" else args.workload)
        if tj:
            traffic = tj["dram_bytes_per_transform"] * B
            traffic_src = tj["source"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name(s), "kernel_ms": fft1_ms,
                "algorithmic_bytes_per_launch": ab["fft1"] * B, "peak_source": peak_src,
                "whole_step_frac": ab["total"] * B * S * args.steps / (ms * 1e-3) / 1e9 / peak}

    # ---- end to end through the host-buffer C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Be = args.e2e_batch or 16 * max(1, (16 << 20) // (4 * s.fft1_block))
        Be = min(Be, B)
        h_timf1 = torch.zeros(pow2_at_least((Be + 2) * s.timf1_blockbytes), dtype=torch.uint8).pin_memory()
        h_timf1[: Be * s.timf1_blockbytes].copy_(torch.from_numpy(host_in[: Be * s.timf1_blockbytes]))
        h_fft1 = torch.zeros(pow2_at_least(Be * s.fft1_block), dtype=torch.float32).pin_memory()
        h_sumsq = torch.zeros(pow2_at_least((Be // s.avg1num + 2) * N), dtype=torch.float32).pin_memory()
        t3s = pow2_at_least((Be + 2) * s.timf3_block + 2 * C * s.mix1_size)
        h_timf3 = torch.zeros(max(nsel, 1) * 2 * t3s, dtype=torch.float32).pin_memory()
        plan2 = api.Plan(s, device=local_rank)
        st2 = api.new_states([s.selfreq_for_bin(b) for b in selbins])
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
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return {"value": Be * spt * nrep * world / dt / 1e6, "unit": "Msamples/s",
                    "h2d_bytes_per_step": (plan2.h2d_bytes() - h0) // nrep, "d2h_bytes_per_step": (plan2.d2h_bytes() - d0) // nrep,
                    "batch": Be}

        # the drop-in call: everything the reference's fft1_b / fft1_c / fft1_mix1_fixed leave in host
        # memory comes back (fft1_float, fft1_sumsq, timf3)
        e2e = e2e_run(False)
        e2e["api"] = "lb200_fft1 + lb200_mix1 on pinned host rings"
        if nsel:
            # same calls with LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE: for set-ups where mix1 is the only
            # reader of fft1_float (second FFT / AFC / network output off), informational
            lazy = e2e_run(True)
            lazy["api"] = "same, fft1_float kept in the device mirror (LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE); fft1_sumsq and timf3 come back"
            e2e["spectrum_on_device"] = lazy
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_quick(args.workload)

    if rank == 0:
        line = {
            "metric": "fft1+mix1 IQ Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "transforms_per_step_per_gpu": B,
                       "l2": f"working set per step {(S * B * (s.timf1_blockbytes + 4 * s.fft1_block)) >> 20} MiB > 126 MiB L2, no flush needed",
                       "mix1_selections": nsel, "fft_avg1num": s.avg1num,
                       "streams_per_gpu": S,
                       "parallelism": f"{world * S} independent receiver streams, {S} per GPU; only the averaged power spectrum is all-reduced"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
