"""world_size-2 check of the multi-GPU host logic on CPU (gloo): each rank evaluates its contiguous shard and the
gathered verdicts equal the single-process result.  (The CUDA kernel is replaced by the table simulator here;
on the GPU box the same sharding feeds one engine per rank, see bench.py.)"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from helpers import Oracle, Sim
from pingoo_b200.shard import shard_batch, shard_range

N = 4001  # deliberately not divisible by the world size


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rules, payloads, _ = synth.make_ruleset(16, config_id=1)
    batch = synth.RequestStream(config_id=1, payloads=payloads).generate(0, N)
    mine = shard_batch(batch, rank, world)
    v = Sim(rules).evaluate(mine).astype(np.int64)
    # max-over-ranks timing plumbing used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    sizes = [shard_range(N, r, world)[1] - shard_range(N, r, world)[0] for r in range(world)]
    bufs = [torch.zeros(s, dtype=torch.int64) for s in sizes]
    pad = max(sizes)
    gathered = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)]
    mine_t = torch.zeros(pad, dtype=torch.int64)
    mine_t[: len(v)] = torch.from_numpy(v)
    dist.all_gather(gathered, mine_t)
    if rank == 0:
        full = np.concatenate([g[:s].numpy() for g, s in zip(gathered, sizes)]).astype(np.uint32)
        np.save(os.path.join(out_dir, "gathered.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    rules, payloads, _ = synth.make_ruleset(16, config_id=1)
    batch = synth.RequestStream(config_id=1, payloads=payloads).generate(0, N)
    want = Oracle(rules).evaluate(batch, threads=4)
    assert len(got) == N and np.array_equal(got, want)


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 1000, 12_500_001):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
