"""N>1 path on CPU: world_size-2 gloo run of the bucketed gradient all-reduce (host logic of row (e))."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

README = dict(filters=(8, 16, 24, 32, 48), se_reduction=(4,) * 5,
              strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
              kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)), att_sub_samp=((1, 1, 1),) * 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import m1b200  # noqa: F401
    from m1b200.model import unets
    from m1b200.model.distribute import BucketedGradSync
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = unets.networks.M1((4, 16, 16), 4, 2, summary=False, build=False, dense_skip=True, probabilistic=True,
                          prob_latent_dims=(3, 2, 1, 0), **README)
    P = m.params
    flat = torch.arange(P.total, dtype=torch.float32) * (rank + 1)
    sync = BucketedGradSync(P, bucket_bytes=64 << 10)
    # completion order of a backward walk: reverse layout order, shared weights finish on their 2nd use
    names = [sp.name for sp in sorted(P.specs.values(), key=lambda sp: sp.offset)]
    uses = {n: (2 if n.startswith('p') else 1) for n in names}
    joins = []
    sync.pre_fire = lambda: joins.append(len(sync.fired))      # Engine.join_side in the product: once per bucket,
    sync.begin(flat, uses)                                      # BEFORE its all-reduce is issued
    fired_before_finish = 0
    for rep in (0, 1):
        for n in reversed(names):
            sync.param_done(n)
        if rep == 0:
            assert not sync.fired or all(uses[nm] == 1 for b in sync.fired
                                         for nm, bb in sync.bucket_of.items() if bb == b)
    fired_before_finish = len(sync.fired)
    sync.finish()
    expect = torch.arange(P.total, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(flat, expect) and joins == list(range(1, len(sync.buckets) + 1))
    q.put((rank, ok, len(sync.buckets), fired_before_finish, sorted(sync.fired) == list(range(len(sync.buckets)))))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, nb, fired, complete in res:
        assert ok, f"rank {rank}: all-reduced gradient differs from the sum over ranks"
        assert nb > 4 and complete
        assert fired == nb, "every bucket must fire from inside the backward walk (overlap), not at finish()"


def test_bucket_layout_covers_flat_buffer():
    import m1b200  # noqa: F401
    from m1b200.model import unets
    from m1b200.model.distribute import BucketedGradSync
    m = unets.networks.M1((4, 16, 16), 3, 2, summary=False, build=False, **README)
    s = BucketedGradSync(m.params, bucket_bytes=16 << 10)
    assert s.buckets[0][0] == 0 and s.buckets[-1][1] == m.params.total
    for (a, b), (c, d) in zip(s.buckets, s.buckets[1:]):
        assert b == c and b > a
    for sp in m.params.specs.values():
        a, b = s.buckets[s.bucket_of[sp.name]]
        assert a <= sp.offset and sp.offset + sp.size <= b
