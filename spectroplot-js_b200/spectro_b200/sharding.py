"""Frame-range sharding of one long message across GPUs.

Unlike the reference's halo-free fan-out (lib/samples.js:253-258, lib/spectroplot.js:1206-1228),
frames keep their GLOBAL positions p_x = ~~(0.5 + stride * x), so the union of the shards is the
unsharded message bit for bit: images / gauges are disjoint column bands, histograms add, min/max
fold.  A shard needs samples [p(x_first), p(x_last) + n): its own range plus a window-length halo.
"""
from __future__ import annotations


def frame_pos(stride: float, x: int) -> int:
    return int(0.5 + stride * x)             # lib/worker.js:72 (int64 instead of int32)


def plan_shards(total_samples: int, n: int, total_width: int, world: int, align_samples: int = 16):
    """-> list of dicts {frame_first, width, sample_first, sample_count} (contiguous frame ranges).
    sample_first is rounded down to `align_samples` so any format's byte offset is 16-byte aligned."""
    stride = (total_samples - n) / (total_width - 1)
    out = []
    # shard boundaries on multiples of 8 frames: the fused kernels write whole 32-byte sectors (8 frames of a row), so
    # every frame keeps the kernel - and therefore the exact fp32 arithmetic - it has in the unsharded message
    cut = lambda g: total_width if g >= world else (g * total_width // world) // 8 * 8
    for g in range(world):
        x0 = cut(g)
        x1 = cut(g + 1)
        if x1 <= x0:
            out.append(dict(frame_first=x0, width=0, sample_first=0, sample_count=0))
            continue
        s0 = frame_pos(stride, x0)
        s1 = min(frame_pos(stride, x1 - 1) + n, total_samples)
        s0 -= s0 % align_samples
        out.append(dict(frame_first=x0, width=x1 - x0, sample_first=s0, sample_count=s1 - s0))
    return out


def shard_fields(shard: dict, total_samples: int, sample_width: int, total_width: int) -> dict:
    """The sp_request shard_* fields for one planned shard."""
    return dict(total_byte_length=total_samples * sample_width, total_width=total_width,
                frame_first=shard["frame_first"], buffer_first_sample=shard["sample_first"])


def merge_stats(parts):
    """Fold per-shard replies like lib/spectroplot.js:1229-1238: histograms add, min/max fold."""
    cB = sum(p["cB_hist"] for p in parts[1:]) + parts[0]["cB_hist"] if len(parts) > 1 else parts[0]["cB_hist"].copy()
    c = sum(p["c_hist"] for p in parts[1:]) + parts[0]["c_hist"] if len(parts) > 1 else parts[0]["c_hist"].copy()
    mn = min([0.0] + [p["dBfs_min"] for p in parts])
    mx = max([-200.0] + [p["dBfs_max"] for p in parts])
    return dict(cB_hist=cB, c_hist=c, dBfs_min=mn, dBfs_max=mx)


def allreduce_stats(dist, hist, minmax):
    """The one exchange step of the multi-GPU path: `hist` (int64 tensor, cB_hist | c_hist) is summed,
    `minmax` (float64 tensor [dBfs_min, dBfs_max]) is folded, in place, over all ranks.  Works on the
    NCCL backend (device tensors, bench.py) and on gloo (CPU tensors, tests)."""
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    minmax[1].neg_()
    dist.all_reduce(minmax, op=dist.ReduceOp.MIN)      # min(min) and -max(max) in one MIN reduction
    minmax[1].neg_()
    return hist, minmax


def stats_buffers(torch, hist_len: int, world: int, device):
    """Buffers of the one-collective merge: `local` = [cB_hist | c_hist | dBfs_min, dBfs_max as float64 bits] (int64),
    written by the engine (histogram pointer = local, min/max pointer = local[hist_len:]), and `gathered` =
    [world][hist_len + 2], filled by gather_stats().  -> (local, hist_view, minmax_view_f64, gathered)"""
    local = torch.zeros(hist_len + 2, dtype=torch.int64, device=device)
    gathered = torch.zeros(world, hist_len + 2, dtype=torch.int64, device=device)
    return local, local[:hist_len], local[hist_len:].view(torch.float64), gathered


def gather_stats(dist, local, gathered):
    """The one exchange step of the multi-GPU path as ONE collective: every rank's histograms and min/max
    (about 10 KB) are all-gathered; fold_gathered() finishes the merge where the reply is read."""
    dist.all_gather_into_tensor(gathered.view(-1), local)
    return gathered


def fold_gathered(torch, gathered, hist_len: int):
    """-> (hist int64 [hist_len] summed over ranks, dBfs_min, dBfs_max) like lib/spectroplot.js:1229-1238."""
    hist = gathered[:, :hist_len].sum(0)
    mm = gathered[:, hist_len:].contiguous().view(torch.float64)
    return hist, min(0.0, float(mm[:, 0].min())), max(-200.0, float(mm[:, 1].max()))
