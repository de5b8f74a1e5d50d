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
