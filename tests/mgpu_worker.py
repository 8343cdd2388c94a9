"""Worker of tests/test_multigpu_gpu.py, launched under torch.distributed.run with one rank per GPU.

  1. independent receiver streams, one per rank: fft1 + fft1_c through the C ABI, the ranks' fft1_sumsq
     rows summed on rank 0 by lb200_reduce_* (copy-engine push over NVLink + add kernel) over several
     rounds (more rounds than mailbox slots: flow control) == the float64 sum of every rank's own rows,
     == the sum of the compiled reference's rows of the same inputs (oracle/_ref) when it is built;
  2. one stream cut into per-rank time-block ranges (shard.block_ranges), each rank on its own device:
     gathered spectra, power rows and baseband == the one-pass run on rank 0.
Prints MGPU_OK on rank 0 when everything held."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from linrad_b200 import api, shard, sizing  # noqa: E402
from linrad_b200.synth import make_timf1  # noqa: E402
from tests.helpers import pow2_at_least, rel_rms, IQ_DATA  # noqa: E402


def rings(s, nblocks, nsel):
    timf1 = np.zeros(pow2_at_least((nblocks + 2) * s.timf1_blockbytes), np.uint8)
    fft1 = np.zeros(pow2_at_least(nblocks * s.fft1_block), np.float32)
    sumsq = np.zeros(pow2_at_least((nblocks // s.avg1num + 2) * s.fft1_size), np.float32)
    t3size = pow2_at_least((nblocks + 2) * s.timf3_block + 2 * s.rf_channels * s.mix1_size)
    timf3 = np.zeros(max(nsel, 1) * 2 * t3size, np.float32)
    return timf1, fft1, sumsq, timf3, t3size


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def exchange(mine):
        out = [None] * world
        dist.all_gather_object(out, mine)
        return out

    # ---------------------------------------------------------------- 1. reduction of power spectra
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=12, mix1_red_n=3)
    N, nblocks, rounds = s.fft1_size, 10, 5
    rows = nblocks // s.avg1num
    plan = api.Plan(s, device=local)
    red = api.Reducer(plan, rank, world, rows * N, exchange=exchange, root=0, depth=2)
    mine, sums = [], []
    d_rows = torch.zeros(rows * N, dtype=torch.float32, device=dev)
    d_out = torch.zeros(rows * N, dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(plan.stream, device=dev)
    raws = []
    timf1, fft1, sumsq, _, _ = rings(s, nblocks, 0)          # Linrad-style: the rings live as long as the plan
    for rd in range(rounds):
        raw = make_timf1(s.input_mode, 1, N, nblocks, s.fft1_new_points, seed=1000 * rd + rank)
        rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]
        raws.append(rawb)
        timf1[:] = 0
        timf1[: rawb.size] = rawb
        plan.fft1_host(timf1=timf1, ref=0, nblocks=nblocks, fft1=fft1, fft1_pa=0, apply_fc=True, sumsq=sumsq, sumsq_pa=0, counter=0)
        mine.append(sumsq[: rows * N].copy())
        red.rows_released()                     # the previous round's copy has left d_rows
        with torch.cuda.stream(stream):
            d_rows.copy_(torch.from_numpy(mine[-1]), non_blocking=False)
        red.push(d_rows.data_ptr())
        if rank == 0:
            red.sum(d_out.data_ptr())
            red.result_ready()
            with torch.cuda.stream(stream):
                sums.append(d_out.clone())
    red.synchronize()
    plan.synchronize()
    torch.cuda.synchronize()
    ok = True
    for rd in range(rounds):
        t = torch.from_numpy(mine[rd]).to(dev)
        parts = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        if rank == 0:
            want = np.sum([p.cpu().numpy().astype(np.float64) for p in parts], axis=0)
            got = sums[rd].cpu().numpy().astype(np.float64)
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err <= 1e-6, f"round {rd}: reduced rows differ from the sum of the ranks' rows ({err})"
            # rank order summation: bit-exact against float32 adds in rank order
            acc = parts[0].cpu().numpy().copy()
            for p in parts[1:]:
                acc = acc + p.cpu().numpy()
            assert np.array_equal(sums[rd].cpu().numpy(), acc), f"round {rd}: not the rank-order float32 sum"
    # against the compiled reference: every stream's rows from oracle/_ref, summed
    try:
        from oracle import refwrap
        have_ref = refwrap.available()
    except Exception:
        have_ref = False
    if have_ref:
        from tests.helpers import run_reference
        kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=12, mix1_red_n=3, version=6)
        ref = run_reference(kw, raws[0], [], nblocks)
        t = torch.from_numpy(ref["sumsq"][: rows * N].astype(np.float32)).to(dev)
        parts = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        if rank == 0:
            want = np.sum([p.cpu().numpy().astype(np.float64) for p in parts], axis=0)
            got = sums[0].cpu().numpy().astype(np.float64)
            strong = want > 1e-4 * want.max()
            worst = (np.abs(got - want)[strong] / want[strong]).max()
            assert worst <= 1e-4, f"reduced spectrum vs the sum of the reference's rows: {worst}"
            print(f"reduce: {world} ranks, {rounds} rounds, worst per-bin error vs reference rows {worst:.2e}")
    red.close()
    plan.close()

    # ---------------------------------------------------------------- 2. time-block sharding on real devices
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3)
    nblocks, sel = 38, [300.37, 1234.0]
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nblocks, s.fft1_new_points, seed=5)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]
    hz = s.ad_speed / s.fft1_size
    br = shard.block_ranges(nblocks, world, avg1num=s.avg1num)[rank]
    timf1, fft1, sumsq, timf3, t3size = rings(s, nblocks, len(sel))
    timf1[: rawb.size] = rawb
    plan = api.Plan(s, device=local)
    states = api.new_states([b * hz for b in sel])
    if br.count:
        start = br.first - br.warmup
        api.advance_mix1_states(plan, states, start)
        if br.warmup:
            plan.fft1_host(timf1=timf1, ref=start * s.timf1_blockbytes, nblocks=br.warmup, fft1=fft1, fft1_pa=start * s.fft1_block,
                           apply_fc=True, sumsq=None)
        plan.fft1_host(timf1=timf1, ref=br.first * s.timf1_blockbytes, nblocks=br.count, fft1=fft1, fft1_pa=br.first * s.fft1_block,
                       apply_fc=True, sumsq=sumsq, sumsq_pa=(br.first // s.avg1num) * s.fft1_size, counter=0)
        plan.mix1_host(fft1=fft1, fft1_px=start * s.fft1_block, nblocks=br.warmup + br.count, states=states, timf3=timf3,
                       timf3_floats=t3size, timf3_pa=start * s.timf3_block)
        plan.synchronize()
    plan.close()
    # gather: every rank contributes its own range, zero elsewhere
    f = np.zeros((nblocks, s.fft1_block), np.float32)
    p = np.zeros((nblocks // s.avg1num, s.fft1_size), np.float32)
    t3 = np.zeros((len(sel), nblocks, s.timf3_block), np.float32)
    sl = slice(br.first, br.first + br.count)
    f[sl] = fft1[: nblocks * s.fft1_block].reshape(nblocks, -1)[sl]
    r0, r1 = br.first // s.avg1num, min((br.first + br.count) // s.avg1num, p.shape[0])
    p[r0:r1] = sumsq[: p.size].reshape(p.shape)[r0:r1]
    for i in range(len(sel)):
        t3[i, sl] = timf3[i * 2 * t3size: i * 2 * t3size + nblocks * s.timf3_block].reshape(nblocks, -1)[sl]
    tf, tp, tt = (torch.from_numpy(x).to(dev) for x in (f, p, t3))
    for t in (tf, tp, tt):
        dist.reduce(t, dst=0)                  # disjoint supports: the sum is the union
    if rank == 0:
        timf1, fft1, sumsq, timf3, t3size = rings(s, nblocks, len(sel))
        timf1[: rawb.size] = rawb
        plan = api.Plan(s, device=local)
        states = api.new_states([b * hz for b in sel])
        plan.fft1_host(timf1=timf1, ref=0, nblocks=nblocks, fft1=fft1, fft1_pa=0, apply_fc=True, sumsq=sumsq, sumsq_pa=0, counter=0)
        plan.mix1_host(fft1=fft1, fft1_px=0, nblocks=nblocks, states=states, timf3=timf3, timf3_floats=t3size, timf3_pa=0)
        plan.synchronize()
        plan.close()
        assert np.array_equal(tf.cpu().numpy(), fft1[: nblocks * s.fft1_block].reshape(nblocks, -1)), "sharded fft1_float != one pass"
        assert np.array_equal(tp.cpu().numpy(), sumsq[: p.size].reshape(p.shape)), "sharded fft1_sumsq != one pass"
        for i in range(len(sel)):
            one = timf3[i * 2 * t3size: i * 2 * t3size + nblocks * s.timf3_block].reshape(nblocks, -1)
            assert rel_rms(tt[i].cpu().numpy(), one) <= 1e-6, "sharded timf3 != one pass"
        print(f"time-block sharding on {world} devices == one pass")
        print("MGPU_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
