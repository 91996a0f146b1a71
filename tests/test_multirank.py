"""CPU, world size 2 over gloo: the N>1 host logic -- contiguous block sharding, no data-path collective,
per-rank CRC chains folded with bra_crc32c_combine, max-over-ranks timing reduction (what bench.py does
under torchrun). The per-rank "device work" is stood in by the oracle here; the GPU path itself is covered by
tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

import bra_workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, block, data, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import bra_pkg
    from oracle_lib import Oracle
    pkg = bra_pkg.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    nblk = (len(data) + block - 1) // block
    lo, hi = pkg.shard_range(nblk, rank, world)
    chain, lens = 0, []
    for b in range(lo, hi):
        blk = data[b * block:(b + 1) * block]
        hdr, payload, crc_raw = oracle.encode_block(blk)
        chain = oracle.crc32c(hdr, chain)     # reference chunks.c:248
        chain = oracle.crc32c(blk, chain)     # == combine(chain, crc_raw, len) of chunks.c:249
        lens.append(len(blk))
    gathered = [None] * world
    dist.all_gather_object(gathered, (chain, pkg.chain_bytes(lens), lo, hi))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # bench.py: max over ranks
    dist.barrier()
    if rank == 0:
        result_q.put((gathered, float(t.item())))
    dist.destroy_process_group()


def test_two_ranks_shard_and_fold(pkg, oracle, vocab):
    block = 4096
    data = wl.gen_text(9 * block + 123, vocab, 7).tobytes()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, block, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    # ranges tile the block list without overlap
    nblk = (len(data) + block - 1) // block
    assert gathered[0][2] == 0 and gathered[0][3] == gathered[1][2] and gathered[1][3] == nblk
    # folding the per-rank chains == the single-process chain over all blocks
    whole = 0
    for b in range(nblk):
        blk = data[b * block:(b + 1) * block]
        hdr, _, _ = oracle.encode_block(blk)
        whole = oracle.crc32c(blk, oracle.crc32c(hdr, whole))
    assert pkg.fold_crc_chains([(g[0], g[1]) for g in gathered]) == whole


def test_shard_range_properties(pkg):
    for nblk in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            ranges = [pkg.shard_range(nblk, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nblk
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
