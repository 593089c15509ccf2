"""CPU, world size 2 over gloo: the N > 1 host logic of the path -- sharding of the batch over ranks and the single
collective (broadcast of the table blob built by rank 0), as bench.py runs it over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libjpeg_b200 import sharding, synth, parse
    total = 11
    frames = [synth.encode(synth.source_image(64, 48, 1 + (i % 3)), 75, (2, 2), 4).tobytes() for i in range(total)]
    b, e = sharding.shard_range(total, rank, world)
    mine = frames[b:e]
    infos = [parse(f) for f in mine]
    blob = sharding.build_tables(mine[0])
    recv, same = sharding.broadcast_tables(blob, dist)
    # a rank whose frames use other tables must notice
    other = synth.encode(synth.source_image(64, 48, 9), 30, (2, 2), 4).tobytes()  # different quantisation
    _, same_other = sharding.broadcast_tables(sharding.build_tables(other) if rank == 1 else blob, dist)
    counts = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(counts)
    q.put((rank, b, e, int(counts.item()), same, same_other, recv.tobytes()[:16], [(i.width, i.height, i.n_intervals) for i in infos]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_share_tables(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, n0, same0, so0, head0, i0), (r1, b1, e1, n1, same1, so1, head1, i1) = res
    assert (b0, e0, b1, e1) == (0, 6, 6, 11) and n0 == n1 == 11          # disjoint, covering, balanced
    assert same0 and same1 and head0 == head1                            # identical tables: one broadcast serves all
    assert so0 and not so1                                               # rank 1's foreign tables are detected
    assert all(x == (64, 48, 3) for x in i0 + i1)


def test_shard_range_properties():
    from libjpeg_b200.sharding import shard_range
    for total in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
