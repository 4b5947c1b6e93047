"""CPU, world_size 2, gloo: the N > 1 path is "every rank decodes its own shard, no data-path
collective" — what has to hold is that the shards cover the batch exactly once, that the result
gather is complete, and that the timing reduction is a max over ranks.  The decode itself is done
by the oracle here (this test has no GPU); on the GPU box bench.py --gpus N runs the same host logic."""
import os
import socket
import zlib

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import datagen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import threebz_b200 as t
    from oracle import o3bz
    lo, hi = t.shard.rank_slice(n_total, rank, world)
    ms = [datagen.member(4096, 9000 + i, "zlib") for i in range(lo, hi)]
    sums = torch.zeros(n_total, dtype=torch.int64)
    for i, (plain, comp) in zip(range(lo, hi), ms):
        r = o3bz.decompress_vector(comp, "zlib", out_cap=len(plain))
        assert r["verdict"] == 0 and r["out"] == plain
        sums[i] = r["checksum"]
    dist.barrier()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)          # host-side result gathering
    times = t.shard.reduce_max([1.0 + rank, 5.0 - rank], dist)
    # size-aware partition through the C ABI: identical on every rank
    lens = [(i * 7919) % 5000 + 1 for i in range(101)]
    owner = t.shard.partition(lens, world)
    own = torch.tensor(owner, dtype=torch.int64)
    chk = own.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((sums.tolist(), times, owner, bool((chk == own).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, n_total = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    sums, times, owner, same = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [zlib.adler32(datagen.text(4096, 9000 + i)) for i in range(n_total)]
    assert sums == want                      # every member decoded exactly once, by exactly one rank
    assert times == [2.0, 5.0]               # max over ranks
    assert same and set(owner) == {0, 1}
    load = [sum(l for l, o in zip([(i * 7919) % 5000 + 1 for i in range(101)], owner) if o == d) for d in range(2)]
    assert abs(load[0] - load[1]) <= 5000


def test_rank_slice_covers_everything():
    import threebz_b200 as t
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = t.shard.rank_slice(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        t.shard.rank_slice(4, 2, 2)
