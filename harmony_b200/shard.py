"""harmony_b200/shard.py -- multi-GPU plumbing for the BLS path (one process per GPU, torch.distributed).

Rounds / triples are independent, so items shard by contiguous index range with NO data-path collective;
the only communication is bookkeeping: max-over-ranks of the device times, sum of the verified counts, and (for the
gathered-result API) an all-gather of the per-rank result bytes.  Works with backend "nccl" (GPU box) and "gloo"
(CPU tests, world_size 2)."""
import torch
import torch.distributed as dist

def shard_range(n: int, rank: int, world: int):
    """Contiguous item range [lo, hi) of `rank`: i in [g*n/G, (g+1)*n/G)."""
    return (rank * n) // world, ((rank + 1) * n) // world

def reduce_step_stats(dev_ms: float, e2e_ms: float, nsig: float, device=None):
    """(max dev_ms, max e2e_ms, sum nsig) over ranks; identity when not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dev_ms, e2e_ms, nsig
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=device)
    s = torch.tensor([nsig], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), float(s[0])

def gather_results(local: bytes, n_total: int, device=None) -> bytes:
    """All ranks obtain the n_total result bytes in item order (each rank verified its shard_range)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bytes(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    mine = torch.zeros(pad, dtype=torch.uint8, device=device)
    mine[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(mine.device)
    outs = [torch.zeros(pad, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, mine)
    return b"".join(bytes(o[:sz].cpu().numpy().tobytes()) for o, sz in zip(outs, sizes))

def all_gather_records(record: bytes, device=None):
    """All ranks obtain every rank's fixed-size partial record (hbls_rlc_partial), in rank order: the one data-path collective of
    the split-batch protocol (NCCL all-gather over NVLink / NVSwitch on the GPU box; 872 B per rank, latency-bound)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [bytes(record)]
    world = dist.get_world_size()
    mine = torch.frombuffer(bytearray(record), dtype=torch.uint8).to(device if device is not None else "cpu")
    outs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]

def verify_triples_split(pks48: bytes, sigs96: bytes, msgs: bytes, msg_len: int, device=None, partial=None, fold=None, exact=None):
    """ONE batch of k independent (pk, msg, sig) triples split over the ranks by contiguous index (BASELINE configs[3]):
    partial record of the local slice -> all-gather -> identical local fold (one final exponentiation).  If the fold proves the
    whole batch, every result is 1; otherwise each rank verifies its slice exactly and the result bytes are gathered.
    Returns (k result bytes in item order -- identical on every rank, True if the combined check settled it).
    partial / fold / exact default to the CUDA backend (bls.RlcPartial / bls.RlcFold / bls.VerifyBatch); the gloo CPU tests inject
    oracle-backed stand-ins to exercise the protocol without a GPU."""
    if partial is None or fold is None or exact is None:
        from harmony_b200 import bls
        partial = partial or bls.RlcPartial; fold = fold or bls.RlcFold; exact = exact or bls.VerifyBatch
    k = len(sigs96) // 96
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(k, rank, world)
    sl = lambda blob, w: blob[lo * w:hi * w]
    rec = partial(sl(pks48, 48), sl(sigs96, 96), sl(msgs, msg_len), msg_len)
    records = all_gather_records(rec, device=device)
    if fold(records):
        return b"\x01" * k, True
    local = exact(sl(pks48, 48), sl(sigs96, 96), sl(msgs, msg_len), msg_len)
    return gather_results(local, k, device=device), False
