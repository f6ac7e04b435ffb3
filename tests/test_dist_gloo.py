"""world_size-2 `gloo` test of the data-parallel host logic (parallel.py): parameter broadcast, gradient all-reduce
hook, batch sharding.  CPU only; the engine is replaced by a minimal stand-in carrying the same buffers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Eng:
    def __init__(self, rank):
        self.params = torch.full((10,), float(rank + 1))
        self.stats = torch.full((4,), float(10 * (rank + 1)))
        self.grads = torch.arange(10, dtype=torch.float32) * (rank + 1)
        self.world_size, self.grad_hook, self._graphs, self._weights_dirty = 1, None, {"x": 1}, False


class _Model:
    def __init__(self, rank):
        self.engine = _Eng(rank)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import parallel
    r, w = parallel.init_process_group_from_env("gloo")
    m = parallel.make_data_parallel(_Model(rank))
    e = m.engine
    ok = r == rank and w == world and e.world_size == 2 and not e._graphs and e._weights_dirty
    ok = ok and torch.equal(e.params, torch.full((10,), 1.0)) and torch.equal(e.stats, torch.full((4,), 10.0))
    e.grad_hook(e.grads)
    ok = ok and torch.equal(e.grads, torch.arange(10, dtype=torch.float32) * 3)       # sum over ranks
    # bucketed form used inside the captured step: two slices reduced asynchronously, then waited for
    g2 = torch.arange(10, dtype=torch.float32) * (rank + 1)
    works = [e.grad_hook(g2[6:], True), e.grad_hook(g2[:6], True)]
    for w in works:
        w.wait()
    ok = ok and torch.equal(g2, torch.arange(10, dtype=torch.float32) * 3)
    lo, hi = parallel.shard_batch(32, rank, world)
    ok = ok and (lo, hi) == (16 * rank, 16 * rank + 16)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_data_parallel_hooks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_batch_rejects_ragged():
    sys.path.insert(0, ROOT)
    import deeplab_b200  # noqa: F401
    from deeplab_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_batch(17, 0, 2)
    assert parallel.shard_batch(128, 7, 8) == (112, 128)
