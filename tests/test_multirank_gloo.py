"""world_size-2 test of the multi-GPU host logic on CPU (gloo): frame-range shard planning with
the window-length halo, the shard_* request fields, and the all-reduce merge used by bench.py.
Each rank renders ITS frames with the CPU oracle (frame by frame, from its own shard buffer only);
the merged result must equal the unsharded message exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

FMT, SW, N, WIDTH, S, SEED = "CS16", 4, 256, 57, 20011, 99


def _render_shard_with_oracle(O, shard_buf, sh, stride, w, wt, cm):
    """Frames [frame_first, frame_first+width) at their GLOBAL positions, read from the shard buffer."""
    cB = np.zeros(1000, np.int64); c = np.zeros(len(cm), np.int64)
    cols, mn, mx = [], 0.0, -200.0
    for x in range(sh["frame_first"], sh["frame_first"] + sh["width"]):
        p0 = int(0.5 + stride * x) - sh["sample_first"]
        fr = shard_buf[p0 * SW:(p0 + N) * SW]
        assert len(fr) == N * SW, "shard buffer (with halo) must cover the frame"
        r = O.render(fr + fr, FMT, N, 2, w, 1 / wt, 6, 30, cm)       # 2 identical frames, stride == N
        cols.append(r.image[:, 0])
        # histograms of a single frame: half of the 2-frame message
        cB += (r.cB_hist // 2).astype(np.int64); c += (r.c_hist // 2).astype(np.int64)
        mn = min(mn, r.dBfs_min); mx = max(mx, r.dBfs_max)
    return np.stack(cols, 1), cB, c, mn, mx


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "spectroplot-js_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from spectro_b200 import sharding
    from helpers import injective_cmap
    cm = injective_cmap(256)
    w, wt = O.window("hann", N)
    stride = (S - N) / (WIDTH - 1)
    sh = sharding.plan_shards(S, N, WIDTH, world)[rank]
    fields = sharding.shard_fields(sh, S, SW, WIDTH)
    assert fields["total_byte_length"] == S * SW and fields["frame_first"] == sh["frame_first"]
    shard_buf = O.synth(FMT, sh["sample_first"], sh["sample_count"], S, SEED).tobytes()   # generated per shard
    img, cB, c, mn, mx = _render_shard_with_oracle(O, shard_buf, sh, stride, w, wt, cm)
    hist = torch.from_numpy(np.concatenate([cB, c]))
    mm = torch.tensor([mn, mx], dtype=torch.float64)
    # the one-collective variant used by bench.py (all-gather + fold) must agree with the two all-reduces
    local, hview, mmview, gathered = sharding.stats_buffers(torch, len(hist), world, "cpu")
    hview.copy_(hist); mmview.copy_(mm)
    sharding.gather_stats(dist, local, gathered)
    g_hist, g_min, g_max = sharding.fold_gathered(torch, gathered, len(hist))
    sharding.allreduce_stats(dist, hist, mm)
    assert torch.equal(g_hist, hist) and g_min == float(mm[0]) and g_max == float(mm[1])
    tiles = [None] * world
    dist.all_gather_object(tiles, (sh["frame_first"], img))
    if rank == 0:
        full = np.concatenate([t for _, t in sorted(tiles, key=lambda t: t[0])], axis=1)
        whole = O.render(O.synth(FMT, 0, S, S, SEED).tobytes(), FMT, N, WIDTH, w, 1 / wt, 6, 30, cm)
        ok = (np.array_equal(full, whole.image) and np.array_equal(hist[:1000].numpy(), whole.cB_hist.astype(np.int64))
              and np.array_equal(hist[1000:].numpy(), whole.c_hist.astype(np.int64))
              and float(mm[0]) == whole.dBfs_min and float(mm[1]) == whole.dBfs_max)
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_shards_merge_to_the_unsharded_message(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
