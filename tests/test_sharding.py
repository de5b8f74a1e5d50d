"""Multi-GPU host logic on CPU: line sharding and the host-side gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from stroke_derenderer_b200.pipeline import gather_in_order, shard_lines
from stroke_derenderer_b200.synth import config_widths, n_tiles_for_width


def test_shard_lines_balanced_and_complete():
    widths = config_widths(4096)
    tiles = np.array([n_tiles_for_width(int(w)) for w in widths])
    assert int(tiles.sum()) == 51565
    assert int(np.array([n_tiles_for_width(int(w)) for w in config_widths(512)]).sum()) == 6438   # config 3
    for world in (1, 2, 4, 8):
        shards = shard_lines(widths, world)
        assert sorted(i for s in shards for i in s) == list(range(4096))
        loads = [int(tiles[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= 20, loads       # one line is at most 20 tiles
        assert shards == shard_lines(widths, world)       # deterministic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    widths = config_widths(64)
    mine = shard_lines(widths, world)[rank]
    # stand-in for the per-line GPU result: something that depends on the line only
    local = [{"line": i, "n_tiles": n_tiles_for_width(int(widths[i])), "sum": int(widths[i]) * 3} for i in mine]
    out = gather_in_order(local, mine, len(widths), world, rank)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_in_input_order_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    widths = config_widths(64)
    assert [o["line"] for o in out] == list(range(64))
    assert all(o["sum"] == int(widths[i]) * 3 for i, o in enumerate(out))


def test_gather_world1():
    out = gather_in_order(["a", "b"], [1, 0], 2, 1, 0)
    assert out == ["b", "a"]


# ---- shared-memory gather arena, two processes (gloo), fake per-rank results --------------------------------
def _fake_rank_results(widths, idx, lpc, writer, seed):
    """What a rank writes in one step, from a seeded generator (stands in for the GPU's D2H copies)."""
    from stroke_derenderer_b200 import _lib
    rng = np.random.default_rng(seed)
    truth = {}
    writer.begin_step()
    for c in range((len(idx) + lpc - 1) // lpc):
        ii = idx[c * lpc:(c + 1) * lpc]
        lines, plan = _lib.plan_lines([int(widths[i]) for i in ii])
        key = ("chunk", c)
        pl = writer.get((key, "planes"), int(plan.px_total)); pl[:] = rng.integers(0, 2, pl.size) * 255
        num = writer.get((key, "num"), 4 * len(ii)).view(np.int32); num[:] = rng.integers(1, 6, len(ii))
        rows = int((num - 1).sum())
        st = writer.get((key, "stats"), 20 * rows).view(np.int32); st[:] = rng.integers(0, 100, st.size)
        lgs = np.concatenate([[0], np.cumsum(rng.integers(0, 3, len(ii)))]).astype(np.int64)
        ng = int(lgs[-1])
        groups = rng.integers(0, 50, (ng, 6)).astype(np.int64)
        writer.put(c, "groups", groups); writer.put(c, "lgs", lgs)
        cr = writer.get((key, "crops"), ng * 224 * 224); cr[:] = rng.integers(0, 255, cr.size)
        writer.set(c, n_lines=len(ii), px_total=int(plan.px_total), n_rows=rows, n_groups=ng, crop_size=224, done=1)
        so = np.concatenate([[0], np.cumsum(num - 1)])
        for k, i in enumerate(ii):
            off, pitch = int(lines[k]["px_off"]), int(lines[k]["pitch"])
            truth[i] = (pl[off:off + 128 * pitch].reshape(128, pitch)[:, :int(widths[i])].copy(), int(num[k]),
                        st.reshape(-1, 5)[so[k]:so[k + 1]].copy(), groups[lgs[k]:lgs[k + 1]].copy(),
                        cr.reshape(-1, 224, 224)[lgs[k]:lgs[k + 1]].copy())
    writer.end_step()
    return truth


def _arena_caps(widths, shards, lpc):
    from stroke_derenderer_b200 import gather as G
    return [G.region_capacity(sum(n_tiles_for_width(int(widths[i])) for i in s), len(s),
                              sum(128 * ((int(widths[i]) + 127) // 128 * 128) for i in s), (len(s) + lpc - 1) // lpc) for s in shards]


def _arena_worker(rank, world, port, name, q):
    from stroke_derenderer_b200 import gather as G
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    widths, lpc = config_widths(24), 4
    shards = shard_lines(widths, world)
    caps = _arena_caps(widths, shards, lpc)
    if rank == 0:
        arena = G.ResultArena(name, caps, rank, create=True)
    dist.barrier()
    if rank != 0:
        arena = G.ResultArena(name, caps, rank, create=False)
    wr = G.RegionWriter(arena.region(rank), (len(shards[rank]) + lpc - 1) // lpc)
    ok = True
    for step in (1, 2):
        _fake_rank_results(widths, shards[rank], lpc, wr, seed=100 * step + rank)
        dist.barrier()                       # every rank's bytes are in the arena
        if rank == 0:
            got = G.GatheredResults(arena, shards, widths, lpc, step=step)
            for r in range(world):          # recompute what each rank wrote from its seed
                import numpy as _np
                scratch = _np.zeros(caps[r], _np.uint8)
                truth = _fake_rank_results(widths, shards[r], lpc, G.RegionWriter(scratch, (len(shards[r]) + lpc - 1) // lpc), seed=100 * step + r)
                for i, (m, n, st, gr, cr) in truth.items():
                    ok &= bool(np.array_equal(got.mask(i), m) and got.num(i) == n and np.array_equal(got.stats(i), st)
                               and np.array_equal(got.groups(i), gr) and np.array_equal(got.crops(i), cr))
            del got
        dist.barrier()                       # readers done before the next step overwrites
    if rank == 0:
        q.put(ok)
    arena.close()
    dist.destroy_process_group()


def test_result_arena_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    name = f"sd_test_arena_{os.getpid()}"
    procs = [ctx.Process(target=_arena_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=180) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not os.path.exists(f"/dev/shm/{name}")


def test_result_arena_rejects_incomplete_region():
    from stroke_derenderer_b200 import gather as G
    import pytest
    widths, lpc = config_widths(6), 4
    shards = shard_lines(widths, 1)
    arena = G.ResultArena(f"sd_test_arena_inc_{os.getpid()}", _arena_caps(widths, shards, lpc), 0, create=True)
    try:
        with pytest.raises(RuntimeError):
            G.GatheredResults(arena, shards, widths, lpc, step=1)      # nothing written yet
        wr = G.RegionWriter(arena.region(0), 2)
        with pytest.raises(MemoryError):
            wr.alloc(arena.region(0).nbytes + 1)
    finally:
        arena.close()
