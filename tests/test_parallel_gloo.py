"""world_size-2 gloo test (CPU) of the image-parallel host logic: sharding + the single
all-gather of per-image records."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, equal, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from e3dge_b200 import parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_latent = 6
        width = par.record_length(n_latent)
        g = torch.Generator().manual_seed(0)
        every = torch.randn(n_images, width, generator=g)  # the records a single process would hold
        lo, hi = par.shard_range(n_images, rank, world)
        got = par.gather_records(every[lo:hi].clone(), equal_shards=equal)
        ok = torch.equal(got, every)
        w, wd, m = par.unpack_records(got, n_latent)
        ok = ok and w.shape == (n_images, 9, 256) and wd.shape == (n_images, n_latent, 512) \
            and m.shape == (n_images, par.N_METRICS)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images,equal", [(8, True), (8, False), (7, False), (1, False)])
def test_sharded_records_gather_to_the_single_process_result(n_images, equal):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, equal, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_shard_range_partitions_exactly():
    from e3dge_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
