#!/usr/bin/env python
"""bench.py -- headline benchmark: BLS aggregate-verify sigs/s on the 250-validator FBFT commit batch (BASELINE.json).

A "step" = one pass of the hot path over one batch of synthetic input: B independent commit rounds
(bitmap32 || aggSig96 || payload48 each, BASELINE configs[1]) verified against one device-resident 250-key
committee, i.e. B x { Mask.SetMask ; Sign.Deserialize ; aggSig.VerifyHash(mask.AggregatePublic, payload) }
(reference internal/chain/engine.go:619-642).  `value` counts constituent signatures (set bits) per second with the
inputs already in HBM; `e2e` is the same metric through the host-buffer C-ABI call (H2D/D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rounds B] [--impl reference]
Multi-GPU: torchrun, one rank per GPU; rounds shard by index with no data-path collective (weak scaling).
"""
import argparse, ctypes, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np
from harmony_b200 import workload as wl

N_COMMITTEE = 250
MSG_LEN = 48
MAC32_PER_MUL, MAC32_PER_SQR = 300, 234          # SURVEY.md 8d fixed conversion (12-limb CIOS)

def log(*a):
    print(*a, file=sys.stderr, flush=True)

# ------------------------------------------------------------------ synthetic workload (deterministic, byte-level only)
def make_committee_sks(n=N_COMMITTEE):
    return [wl.seeded_sk("c2", i) for i in range(n)]

def make_rounds(sks, B, seed, rank=0):
    """B rounds: distinct bitmap (k cycles 167/200/250 = quorum .. full), distinct 48-byte commit payload,
    and the secret scalar sum(sk_i over the bitmap) mod r whose SignHash IS the round's aggregate signature."""
    n = len(sks)
    rng = np.random.Generator(np.random.Philox(key=[seed, rank]))
    ks = np.array([wl.quorum_k(n), 200, n], dtype=np.int64)[np.arange(B) % 3]
    order = np.argsort(rng.random((B, n)), axis=1)
    member = np.zeros((B, n), dtype=bool)
    np.put_along_axis(member, order, np.arange(n)[None, :] < ks[:, None], axis=1)
    padded = np.zeros((B, ((n + 7) // 8) * 8), dtype=np.uint8); padded[:, :n] = member
    bitmaps = np.packbits(padded, axis=1, bitorder="little")                       # LSB-first (mask.go:110-112)
    limbs = np.array([[(k >> (32 * j)) & 0xffffffff for j in range(8)] for k in sks], dtype=np.int64)
    sums = member.astype(np.int64) @ limbs                                          # B x 8, each < 250 * 2^32
    agg_sk = b"".join(wl.sk_bytes(sum(int(sums[b, j]) << (32 * j) for j in range(8))) for b in range(B))
    payload = np.zeros((B, MSG_LEN), dtype=np.uint8)                                # LE64(blockNum) || hash32 || LE64(viewID)
    payload[:, 0:4] = rng.integers(0, 256, (B, 4), dtype=np.uint8)
    payload[:, 8:40] = rng.integers(0, 256, (B, 32), dtype=np.uint8)
    payload[:, 40:44] = rng.integers(0, 256, (B, 4), dtype=np.uint8)
    nsig = int(member.sum())
    return bitmaps.tobytes(), agg_sk, payload.tobytes(), nsig

# ------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, dev):
        self.dev = dev; self.samples = []; self.reasons = set(); self.stop = False; self.t = None; self.maxmhz = None
    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.maxmhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v: self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)
    def start(self): self.t = threading.Thread(target=self._run, daemon=True); self.t.start()
    def finish(self):
        self.stop = True
        if self.t: self.t.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons), "samples": len(s)}

# ------------------------------------------------------------------ CPU arm (oracle = restated reference path)
def oracle_lib():
    """The only place bench.py touches oracle/: cpu_baseline leg and --impl reference."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    return ol.load()

def cpu_rounds_per_s(orc, pks, bitmaps, sigs, msgs, blen, budget_s, threads):
    """Reference call pattern per round: NewMask+SetMask (250 G1 adds) ; Sign.Deserialize ; VerifyHash, on `threads` host threads
    (ctypes releases the GIL).  Runs ~budget_s seconds; returns (rounds/s, rounds done, all results correct)."""
    B = len(sigs) // 96
    h = orc.committee(pks)
    done = [0] * threads; good = [True] * threads
    t_end = time.perf_counter() + budget_s
    def work(t):
        j = t
        while time.perf_counter() < t_end:
            r = j % B
            rc = orc.committee_aggregate_verify(h, bitmaps[r * blen:(r + 1) * blen], sigs[r * 96:(r + 1) * 96], msgs[r * MSG_LEN:(r + 1) * MSG_LEN])
            good[t] &= (rc == 1); done[t] += 1; j += threads
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ths]; [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done), all(good)

# executed work of the batched pairing stage: (Fp mul, Fp sqr) per GROUP of G rounds, counted by running the device code
# compiled for the host (tests/emu: emu_rlc_stage_counts; pinned by tests/test_emu_logic.py::test_emu_rlc_stage_counts)
RLC_EXEC_FP_OPS = {4: {"scale": (7856, 2576), "pairing": (34661, 764)},
                   8: {"scale": (15046, 5094), "pairing": (54329, 764)}}
NCU_DRAM_BYTES_PER_LAUNCH = {("k_rlc_pairing_split<8>", 303104): 28.78e9 + 109.09e9}
def rlc_group_size(B, sm_count, tpb_split=512):
    """Mirror of the host's choice in hbls.cu launch_verify_tail: 8 when B/8 lane pairs still fill every SM, else 4."""
    env = os.environ.get("HBLS_RLC_G")
    if env in ("4", "8"): return int(env)
    return 8 if 2 * (B // 8) >= sm_count * tpb_split else 4
def rlc_exec_mac32(G):
    ops = RLC_EXEC_FP_OPS[G]
    return tuple((ops[k][0] * 300 + ops[k][1] * 234) / G for k in ("scale", "pairing"))

def stage_mac32_per_round(orc, pks, bitmaps, sigs, msgs, blen, sample=12):
    """ALGORITHMIC work per round and per pipeline stage from the oracle's Fp mul/sqr counter."""
    h = orc.committee(pks)
    fn = orc.L.ho_profile_aggregate_verify
    fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
    acc = np.zeros(12)
    for r in range(sample):
        out = (ctypes.c_uint64 * 12)()
        rc = fn(h, bitmaps[r * blen:(r + 1) * blen], blen, sigs[r * 96:(r + 1) * 96], msgs[r * MSG_LEN:(r + 1) * MSG_LEN], MSG_LEN, out)
        assert rc == 1
        acc += np.array(list(out), dtype=np.float64)
    acc /= sample
    return [float(acc[2 * s] * MAC32_PER_MUL + acc[2 * s + 1] * MAC32_PER_SQR) for s in range(6)]

# ------------------------------------------------------------------ main arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sks = make_committee_sks()
    orc = oracle_lib()
    pks = [orc.get_public_key(wl.sk_bytes(k)) for k in sks]
    S = 64
    bitmaps, agg_sk, msgs, nsig = make_rounds(sks, S, seed=2024)
    sigs = b"".join(orc.sign_hash(agg_sk[32 * j:32 * j + 32], msgs[MSG_LEN * j:MSG_LEN * j + MSG_LEN]) for j in range(S))
    blen = (N_COMMITTEE + 7) // 8
    per = max(1.0, min(8.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rounds_per_s(orc, pks, bitmaps, sigs, msgs, blen, min(per, 1.0), threads)
    rates = []; ok = True; rounds = 0
    for _ in range(args.steps):
        r, d, g = cpu_rounds_per_s(orc, pks, bitmaps, sigs, msgs, blen, per, threads)
        rates.append(r); ok &= g; rounds += d
    rps = float(np.mean(rates)); sig_per_round = nsig / S
    value = rps * sig_per_round
    line = {"impl": "reference", "metric": "BLS aggregate-verify sigs/sec (250-validator FBFT commit batch)", "value": value, "unit": "sigs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 (6x64-bit Montgomery limbs)", "data": "synthetic",
            "config": {"workload": "FBFT commit-phase: 250-validator committee, FastAggregateVerify per round (SetMask + Deserialize + VerifyHash)",
                       "committee": N_COMMITTEE, "msg_len": MSG_LEN, "signers_per_round": "167/200/250 cycling"},
            "cpu_baseline": {"value": value, "unit": "sigs/s", "cores": threads, "kind": "port",
                             "sample": f"{rounds} rounds over {args.steps} x {per:.1f}s windows on {threads} threads; restated CPU path (oracle/hbls_oracle.c), real libbls not buildable here"},
            "e2e": {"value": value, "unit": "sigs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "all_correct": bool(ok)}
    print(json.dumps(line), flush=True)

def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the BLS backend has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the single JSON line (no NCCL version banner)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"): os.environ["NCCL_DEBUG"] = "WARN"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from harmony_b200 import bls
    bls.Init(device=local)
    L = bls.lib()
    B = args.rounds
    blen = (N_COMMITTEE + 7) // 8
    sks = make_committee_sks()
    pks_blob = bls.GetPublicKeyBatch(b"".join(wl.sk_bytes(k) for k in sks))
    pks = [pks_blob[48 * i:48 * i + 48] for i in range(N_COMMITTEE)]
    com = bls.Committee(pks)
    t0 = time.time()
    bitmaps, agg_sk, msgs, nsig = make_rounds(sks, B, seed=2024, rank=rank)
    sigs, ok = bls.SignHashBatch(agg_sk, msgs, MSG_LEN)
    assert ok == b"\x01" * B
    log(f"[rank {rank}] inputs: {B} rounds, {nsig} constituent sigs, generated in {time.time() - t0:.1f}s")

    # device-resident copies (value) and pinned host copies (e2e)
    def dev(b): return torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    def pin(b): return torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()
    d_bm, d_sig, d_msg = dev(bitmaps), dev(sigs), dev(msgs)
    d_res = torch.zeros(B, dtype=torch.uint8, device="cuda")
    h_bm, h_sig, h_msg = pin(bitmaps), pin(sigs), pin(msgs)
    h_res = torch.zeros(B, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")               # > 126 MB L2
    stream = torch.cuda.Stream()

    def step_device():
        rc = L.hbls_aggregate_verify_batch_device(com.h, B, d_bm.data_ptr(), blen, d_sig.data_ptr(), d_msg.data_ptr(), MSG_LEN, d_res.data_ptr(), stream.cuda_stream)
        assert rc == 0, rc
    def step_host():
        rc = L.hbls_aggregate_verify_batch(com.h, B, h_bm.data_ptr(), blen, h_sig.data_ptr(), h_msg.data_ptr(), MSG_LEN, h_res.data_ptr())
        assert rc == 0, rc
    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_(); step_device()
    barrier()
    assert int(d_res.sum().item()) == B, "warm-up verification returned a false negative"

    sampler = ClockSampler(local); sampler.start()
    launches0 = bls.KernelLaunchCount()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms = np.zeros(6)
    barrier()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            flush.zero_()                                   # L2 flush between timed iterations (outside the event pair)
            evs[i][0].record(stream); step_device(); evs[i][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = bls.KernelLaunchCount() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    assert int(d_res.sum().item()) == B
    # per-kernel durations: separate passes with event records between the kernels (kept out of the timed region)
    bls.StageTimingEnable(True)
    n_stage = 3
    with torch.cuda.stream(stream):
        for i in range(n_stage):
            flush.zero_(); step_device()
            st = bls.StageTimingGet(); stage_ms += np.array(st if len(st) == 6 else [0] * 6)
    bls.StageTimingEnable(False)
    barrier()
    stage_ms /= n_stage

    # the exact per-round algorithm (batch mode 0: one 2-pair Miller loop + final exponentiation per round) on the same batch
    bls.SetBatchMode(0)
    ex_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
    with torch.cuda.stream(stream):
        flush.zero_(); step_device()
        for a, b in ex_evs:
            flush.zero_(); a.record(stream); step_device(); b.record(stream)
    barrier()
    exact_ms = float(np.median([a.elapsed_time(b) for a, b in ex_evs]))
    assert int(d_res.sum().item()) == B
    bls.SetBatchMode(1)

    # e2e: host buffers through the public C-ABI call, copies inside the timed region
    step_host(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert int(h_res.sum().item()) == B
    clocks = sampler.finish()
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        rc = L.hbls_aggregate_verify_batch(com.h, 1, h_bm.data_ptr(), blen, h_sig.data_ptr(), h_msg.data_ptr(), MSG_LEN, h_res.data_ptr())
        lat.append((time.perf_counter() - t0) * 1e3)
    single_round_ms = float(np.median(lat))
    # informational: the leader's commit-phase vote collection (R9, consensus/leader.go:227-290): 250 individual
    # signatures on ONE payload, one device call (single-bit bitmaps; H(m) computed once)
    vmsg = msgs[:MSG_LEN]
    vsigs, vok = bls.SignHashBatch(b"".join(wl.sk_bytes(k) for k in sks), vmsg * N_COMMITTEE, MSG_LEN)
    vbm = bytearray(N_COMMITTEE * blen)
    for i in range(N_COMMITTEE): vbm[i * blen + (i >> 3)] |= 1 << (i & 7)
    vlat = []
    for _ in range(5):
        t0 = time.perf_counter()
        vres = com.AggregateVerifyBatch(bytes(vbm), vsigs, vmsg * N_COMMITTEE, MSG_LEN)
        vlat.append((time.perf_counter() - t0) * 1e3)
    assert vres == b"\x01" * N_COMMITTEE
    votes250_ms = float(np.median(vlat))

    from harmony_b200 import shard
    dev_ms_max, e2e_ms_max, nsig_total = shard.reduce_step_stats(dev_ms, e2e_s * 1e3, float(nsig), device="cuda")
    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return

    value = nsig_total * args.steps / (dev_ms_max * 1e-3)
    e2e_value = nsig_total * args.steps / (e2e_ms_max * 1e-3)

    # roofline of the dominant kernel: integer pipe (IMAD.WIDE) -- HBM is idle by construction
    peak = bls.ProbeMac32PerS(8192)
    orc = oracle_lib()
    S = min(B, 64)
    macs = stage_mac32_per_round(orc, pks, bitmaps, sigs, msgs, blen, sample=min(12, S))
    names = list(bls.STAGE_NAMES)
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    if B >= sm_count * 256: names[0] = "k_mask_aggregate_serial"
    exact_macs = list(macs)                         # the oracle's per-round (exact) algorithm, stage by stage
    rlc = bls.GetBatchMode() == 1 and B >= 1024 and os.environ.get("HBLS_RLC", "1") != "0"
    if rlc:
        # batched form: slot 4 = coefficient scaling + group sums, slot 5 = (G+1)-pair Miller loop + ONE final exponentiation
        # per group of G rounds (G = 8 when the batch fills the chip that way, else 4).  Executed Fp-mul/sqr counts of that algorithm come from the device code compiled
        # for the host (tests/emu: emu_rlc_stage_counts; pinned by tests/test_emu_logic.py), x300 / x234 MAC32 each.
        G = rlc_group_size(B, sm_count)
        names[4], names[5] = "k_rlc_scale+k_rlc_group_sum", f"k_rlc_pairing_split<{G}>"
        macs = macs[:4] + list(rlc_exec_mac32(G))
    elif stage_ms[4] < 0.02 * stage_ms[5]:         # fused launch: Miller loops + final exponentiation in one kernel
        macs = macs[:4] + [0.0, macs[4] + macs[5]]
        names[5] = "k_pairing_verify_split" if os.environ.get("HBLS_SPLIT", "1") != "0" else "k_pairing_verify"
    dom = int(np.argmax(stage_ms))
    achieved = macs[dom] * B / (stage_ms[dom] * 1e-3)
    total_macs = sum(macs)
    step_s = dev_ms / args.steps * 1e-3
    bytes_per_round = blen + 96 + MSG_LEN + 1
    roofline = {"bound": "int32-imad", "kernel": names[dom], "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TMAC32/s",
                "frac": achieved / peak,
                # DRAM bytes of ONE launch of the dominant kernel from the committed ncu --set full capture of this workload
                # (profiles/r1_ncu_rlc_pairing_g8.txt): per-thread Fp12 temporaries spilling past L2, not input traffic
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((names[dom], B)),
                "traffic_source": "profiles/r1_ncu_rlc_pairing_g8.txt (dram__bytes_read.sum + dram__bytes_write.sum)" if (names[dom], B) in NCU_DRAM_BYTES_PER_LAUNCH else None,
                "peak_source": "hbls_probe_mac32_per_s: register-resident IMAD.WIDE.U32 probe measured live on this GPU",
                "algorithm": (f"random-linear-combination batch check, groups of {G} rounds (exact per-round pass only when a group fails)"
                              if rlc else "exact per-round FastAggregateVerify"),
                "work_counted": "Fp multiplications/squarings the kernel's algorithm performs (x300 / x234 MAC32), not instructions issued",
                "algorithmic_mac32_per_round": total_macs, "kernel_mac32_per_round": macs[dom],
                "pipeline_frac": total_macs * B / step_s / peak,
                "exact_algorithm_mac32_per_round": sum(exact_macs),
                "pipeline_frac_exact_equivalent": sum(exact_macs) * B / step_s / peak,
                "stage_ms_note": "per-kernel CUDA-event times from 3 extra passes after the timed region",
                "stage_ms": {n: float(m) for n, m in zip(names, stage_ms)},
                "stage_mac32_per_round": {n: m for n, m in zip(names, macs)},
                "hbm_algorithmic_gbs": bytes_per_round * B / step_s / 1e9}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_peak_gbs_measured"] = peaks.get("hbm_gbs")
    except Exception:
        roofline["hbm_peak_gbs_measured"] = None

    threads = os.cpu_count() or 1
    cpu = None
    if world == 1:
        rps, done, good = cpu_rounds_per_s(orc, pks, bitmaps[:S * blen], sigs[:S * 96], msgs[:S * MSG_LEN], blen, 12.0, threads)
        rps1, done1, good1 = cpu_rounds_per_s(orc, pks, bitmaps[:S * blen], sigs[:S * 96], msgs[:S * MSG_LEN], blen, 4.0, 1)
        spr = nsig / B
        cpu = {"value": rps * spr, "unit": "sigs/s", "cores": threads, "kind": "port",
               "sample": f"first {S} rounds of the same workload cycled for 12 s on {threads} threads ({done} rounds); restated CPU path oracle/hbls_oracle.c",
               "single_thread_value": rps1 * spr, "all_correct": bool(good and good1),
               "reference_published_ms": {"VerifyHash": 1.5, "Sign.Deserialize": 0.52, "src": "reference test/chain/vrf/main.go:109-112, test/chain/reward/main.go:239-245 (hardware unspecified)"}}

    line = {"metric": "BLS aggregate-verify sigs/sec (250-validator FBFT commit batch)", "value": value, "unit": "sigs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (12x32-bit Montgomery limbs, IMAD.WIDE)",
            "data": "synthetic",
            "config": {"workload": "FBFT commit-phase: 250-validator committee, single message per round, FastAggregateVerify per round (BASELINE configs[1]), " + f"{world}xB200",
                       "committee": N_COMMITTEE, "rounds_per_step_per_gpu": B, "signers_per_round": "167/200/250 cycling", "msg_len": MSG_LEN,
                       "sharding": "rounds by index, no data-path collective", "l2": "256 MiB flush between timed steps",
                       "timing": "CUDA events per step on the launching stream, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "sigs/s", "h2d_bytes_per_step": B * (blen + 96 + MSG_LEN), "d2h_bytes_per_step": B,
                    "ms_per_step": e2e_ms_max / args.steps, "api": "hbls_aggregate_verify_batch (pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_s_timed_region": t_wall,
            "exact_mode": {"ms_per_step": exact_ms, "value_this_rank": nsig / (exact_ms * 1e-3), "unit": "sigs/s",
                           "note": "hbls_set_batch_mode(0): every round verified on its own (no random linear combination), rank 0, 3 steps"},
            "single_round_latency_ms": single_round_ms, "leader_250_votes_same_msg_latency_ms": votes250_ms}
    if cpu: line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1: dist.destroy_process_group()

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=303104, help="rounds per step per GPU (default: 37 888 groups of 8 = one lane pair per group on 148 SMs x 512 threads)")
    ap.add_argument("--impl", default="hbls", choices=["hbls", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "hbls":
        log("note: warm-up < 3 steps requested")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)

if __name__ == "__main__":
    main()
