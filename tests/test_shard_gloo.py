"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: index sharding, max/sum reduction of step statistics and the
ordered all-gather of per-item results -- the N>1 path of bench.py minus the kernels."""
import os, socket
import torch.multiprocessing as mp

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p

def _worker(rank, world, port, q):
    import torch.distributed as dist
    from harmony_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10001
    lo, hi = shard.shard_range(n, rank, world)
    local = bytes((i * 7 + 3) % 2 for i in range(lo, hi))        # stand-in for this rank's verify booleans
    full = shard.gather_results(local, n)
    stats = shard.reduce_step_stats(10.0 + rank, 20.0 - rank, float(hi - lo))
    q.put((rank, lo, hi, full == bytes((i * 7 + 3) % 2 for i in range(n)), stats))
    dist.destroy_process_group()

def test_two_rank_shard_reduce_gather():
    world = 2; port = _free_port()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 10001          # contiguous cover
    for r in res:
        assert r[3] is True
        assert r[4] == (11.0, 20.0, 10001.0)                                          # max, max, sum

def test_shard_range_edges():
    from harmony_b200 import shard
    for n in (0, 1, 7, 250, 10000):
        for w in (1, 2, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
