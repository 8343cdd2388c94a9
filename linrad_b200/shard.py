"""Multi-GPU sharding of the path: one process per GPU, no data-path collective.

Receiver streams are independent (each keeps its own fft1_sumsq_counter, mix1_phase[] and timf3
overlap state), so they are dealt out to ranks in contiguous groups; a single stream can instead
be cut into contiguous time-block ranges, each rank re-reading the fft1_interleave_points halo in
front of its range (the same independence the reference uses when it farms time blocks to its
fft1b worker threads, wcw.c:974-1000).  The only thing that crosses GPUs is the averaged power
spectrum: one all-reduce(sum, float32) of fft1_sumsq rows per averaging period -- NCCL on the
GPU box, gloo in the CPU tests.
"""
from dataclasses import dataclass


def stream_assignment(n_streams, world):
    """Contiguous, balanced: rank r owns streams [lo_r, hi_r)."""
    if world < 1 or n_streams < 0:
        raise ValueError("bad world/stream count")
    base, extra = divmod(n_streams, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(lo, lo + n))
        lo += n
    return out


@dataclass
class BlockRange:
    first: int          # first transform of the rank's range
    count: int          # transforms in the range
    warmup: int         # transforms to run (and discard) before `first` so that mix1's overlap
                        # half and phase state at the seam are those of the sequential run
    group_offset: int   # fft1_sumsq_counter at `first` (ranges are cut on averaging-group borders)


def block_ranges(nblocks, world, avg1num=1, mix1_overlap=True):
    """Cut one stream's transforms into per-rank ranges aligned to fft1_sumsq averaging groups
    (fft1.c:4507-4520) so that every fft1_sumsq row is produced by exactly one rank."""
    groups = (nblocks + avg1num - 1) // avg1num
    out = []
    for gr in stream_assignment(groups, world):
        first = gr.start * avg1num
        last = min(gr.stop * avg1num, nblocks)
        cnt = max(0, last - first)
        out.append(BlockRange(first=first, count=cnt, warmup=1 if (mix1_overlap and first > 0 and cnt > 0) else 0,
                              group_offset=0))
    return out


def reduce_sumsq(rows, group=None):
    """All-reduce (sum) of averaged power rows, in place.  `rows` is a torch tensor on the
    rank's device; backend NCCL (GPU) or gloo (CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)
    return rows
