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

def _split_worker(rank, world, port, q, corrupt):
    """verify_triples_split over gloo with oracle-backed stand-ins for the device calls (the record carries the slice's item / bad
    counts; the real records are group elements, tests/test_gpu_parity.py::test_split_batch_partial_records_fold)."""
    import os, struct, sys
    import torch.distributed as dist
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from harmony_b200 import shard, workload as wl
    import oracle_lib
    orc = oracle_lib.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k = 13
    sks = [wl.sk_bytes(wl.seeded_sk("split", i)) for i in range(k)]
    msgs = [wl.seeded_bytes("split/m", i, 32) for i in range(k)]
    pks = [orc.get_public_key(s) for s in sks]; sigs = [orc.sign_hash(s, m) for s, m in zip(sks, msgs)]
    for i in corrupt: msgs[i] = bytes([msgs[i][0] ^ 1]) + msgs[i][1:]
    calls = {"partial": 0, "exact": 0}
    def exact(p, s, m, ml):
        calls["exact"] += 1
        n = len(s) // 96
        return bytes(1 if orc.verify_hash(s[96 * i:96 * i + 96], p[48 * i:48 * i + 48], m[ml * i:ml * i + ml]) else 0 for i in range(n))
    def partial(p, s, m, ml):
        calls["partial"] += 1
        r = exact(p, s, m, ml); calls["exact"] -= 1
        return struct.pack("<II", len(r), len(r) - sum(r)).ljust(872, b"\0")
    def fold(records):
        assert len(records) == world and all(len(r) == 872 for r in records)
        return all(struct.unpack("<II", r[:8])[1] == 0 for r in records)
    res, settled = shard.verify_triples_split(b"".join(pks), b"".join(sigs), b"".join(msgs), 32, partial=partial, fold=fold, exact=exact)
    q.put((rank, res, settled, calls["partial"], calls["exact"]))
    dist.destroy_process_group()

def test_two_rank_split_batch_protocol():
    for corrupt in ([], [2, 11]):
        world = 2; port = _free_port()
        ctx = mp.get_context("spawn"); q = ctx.Queue()
        ps = [ctx.Process(target=_split_worker, args=(r, world, port, q, corrupt)) for r in range(world)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=180) for _ in range(world))
        [p.join(timeout=60) for p in ps]
        want = bytes(0 if i in corrupt else 1 for i in range(13))
        for r in res:
            assert r[1] == want and r[2] == (not corrupt) and r[3] == 1 and r[4] == (1 if corrupt else 0)
