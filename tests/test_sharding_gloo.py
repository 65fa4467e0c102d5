"""CPU, world_size 2 over gloo: the N>1 host path - shard on the chunk grid, code shards
independently (here with the CPU emulation of the kernels, there is no GPU), exchange only the
compressed sizes, lay the container out from their prefix sum.  The merged container must be the
single-process container byte for byte ("chunks partitioned by index across the GPUs")."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import emul
    from slimfastq_b200 import api, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.illumina(3000)
    a, b, phase = api.split_on_grid(data, world, 1 << 18)[rank]
    blob = emul.compress(data[a:b], 3, 1 << 18, phase=phase)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(blob)], dtype=torch.int64))       # the only exchange
    offsets = [0]
    for s in sizes:
        offsets.append(offsets[-1] + int(s))
    gathered = [None] * world
    dist.all_gather_object(gathered, blob)
    if rank == 0:
        merged = api.merge_containers(gathered)
        q.put((offsets, emul.decompress(merged) == data and merged == emul.compress(data, 3, 1 << 18), len(merged)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    offsets, ok, n = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok and len(offsets) == 3 and offsets[-1] > n - 200
