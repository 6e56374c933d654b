#!/usr/bin/env python
"""Benchmark of the B200-native hyper-greco proving path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps K --warmup W                 # our arm
    python bench.py --impl reference --gpus 1 --steps K --warmup W   # CPU arm: the oracle port on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

    torchrun ... bench.py --gpus N --shard                         # extra: ONE Lasso-node proof split over N GPUs (strong scaling)

A step = one batch of `--inflight` independent proofs per GPU (default 4; each in its own context on its own host thread, so that
host work and PCIe copies of one proof overlap device work of the others); a proof = gkr::prove_gkr of the BFV SK-encryption
circuit (sk_encryption_circuit.rs:455-457) on a synthetic witness of the n=32768, k=16, Goldilocks parameter set (BASELINE.json
metric config). `value` = proofs/s over all GPUs; the latency of one proof alone is reported next to it. ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.md row "Goldilocks/Ext2, n=32768, 59, 16 - GKR prove: 5.06 s (0.198 proofs/s)" (reference README.md:44, Apple M1):
# the only published number for this metric; vs_baseline = value / this
PUBLISHED_PROOFS_PER_SEC = 1.0 / 5.06
METRIC = "gkr_prove_proofs_per_sec"
UNIT = "proofs/s"
DEFAULT_CONFIG = "32768_16x59_65537"
LAST_WITNESS = None


def workload_desc(name, P, nv, m):
    return {
        "workload": f"gkr::prove_gkr of the BFV SK-encryption circuit n={P.N} k={P.K} goldilocks/ext2: Lasso node (num_vars={nv}, memories={m}, C=4, M=65536) "
                    f"+ {2 * P.K + 1} FFT layers of 2^{P.log2_size} + {P.K} product layers + the Vanilla relay/scale/sum layers; a step = one batch of independent proofs per GPU",
        "scope": "the reference's `GKR prove` span (sk_encryption_circuit.rs:455-457): every node's claim reduction incl. LassoNode::polynomialize; "
                 "circuit values resident on the device (witness gen = circuit.evaluate is outside the span, as in the reference)",
        "params": name,
        "parallelism": "independent proof instances: `proofs_in_flight_per_gpu` per GPU (one host thread + one context each), no data-path collective",
        "cache": "working set ~4 GB per proof >> 126 MB L2, no explicit flush between steps",
    }


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML every 100 ms; faster polling measurably slows the
    launch path through the driver lock)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz, self._stop, self._t = index, [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80,
                 "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_case(name, seed):
    import numpy as np
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params, witness
    P = params.PARAMS[name]
    args = witness.synth_witness(P, seed)
    inp = np.array(witness.lasso_inputs(P, args), dtype=np.uint64)
    global LAST_WITNESS
    LAST_WITNESS = witness.get_inputs(P, args)
    return P, inp, witness.lasso_lookup_bounds(P), witness.lasso_lookup_segments(P), witness.lasso_num_vars(P)


def oracle_case(hgo, bounds, segs):
    import numpy as np
    opp = hgo.Preprocessing(bounds)
    rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in segs])
    return opp, rows


def run_reference(args):
    """CPU arm: the reference's own implementation cannot be built here (Rust nightly + un-vendored git deps),
    so this times the oracle port (oracle/protocol.hpp + oracle/gkr.hpp, OpenMP over all host cores) on the same workload:
    the `GKR prove` span with the circuit already evaluated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import hgo
    hgo.build()
    P, inp, bounds, segs, nv = make_case(args.config, 0)
    opp, rows = oracle_case(hgo, bounds, segs)
    hgo.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    cores = hgo.num_threads()
    ins, ct0is = LAST_WITNESS
    prep = hgo.bfv_prepare(0, P, ins, ct0is)
    t0 = time.perf_counter()
    hgo.bfv_prove_prepared(prep)
    t1 = time.perf_counter() - t0
    budget = 240.0
    steps = max(1, min(args.steps, int(budget / max(t1, 1e-3)) - 1))
    warm = 0 if steps < args.steps else max(0, min(args.warmup - 1, 1))
    for _ in range(warm):
        hgo.bfv_prove_prepared(prep)
    t0 = time.perf_counter()
    for _ in range(steps):
        hgo.bfv_prove_prepared(prep)
    dt = time.perf_counter() - t0
    v = steps / dt
    sample = f"{steps} full proofs of the same workload (first call {t1:.1f}s used as warm-up" + (f"; --steps {args.steps} capped to fit ~4 min" if steps < args.steps else "") + ")"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm + 1,
            "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": v / PUBLISHED_PROOFS_PER_SEC, "dtype": "u64", "data": "synthetic",
            "config": workload_desc(args.config, P, nv, opp.num_memories),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class ProofSlot:
    """One proof in flight: its own hg_ctx (streams), prover (6 GB of device buffers), pinned host copy of the witness."""

    def __init__(self, api, np, torch, P, host_np, ct_np, device):
        self.api, self.np = api, np
        self.ctx = api.Context(device)
        self.prover = api.BfvSkEncryptProver(self.ctx, P)             # setup + configure (sk_encryption_circuit.rs:319-363)
        n_in = sum(v.size for v in host_np)
        self.pinned = torch.empty(n_in + ct_np.size, dtype=torch.int64).pin_memory()
        h_all = self.pinned.numpy().view(np.uint64)
        self.h_views, off = [], 0
        for v in host_np:
            h_all[off:off + v.size] = v
            self.h_views.append(h_all[off:off + v.size])
            off += v.size
        self.h_ct = h_all[n_in:]
        self.h_ct[:] = ct_np
        self.dev_inputs = [api.DeviceBuffer.from_numpy(self.ctx, v) for v in host_np]
        self.d_ct = api.DeviceBuffer.from_numpy(self.ctx, ct_np)
        self.prover.circuit.evaluate(self.dev_inputs)                 # witness gen (outside the `GKR prove` span, :439-453)
        tr0 = api.Keccak256Transcript()
        L = self.prover.ct0is_log2_size
        point = tr0.squeeze_challenges(L)                             # :445
        value = api.mle_eval_batch(self.ctx, self.d_ct, 1, L, point)[0]   # :446
        el = point.shape[1]
        self.out_claims = [(np.zeros((0, el), np.uint64), np.zeros(el, np.uint64)), (point, value)]

    def resident(self):
        """the `GKR prove` span: gkr::prove_gkr on device-resident circuit values"""
        tr = self.api.Keccak256Transcript()
        tr.squeeze_challenges(self.prover.ct0is_log2_size)
        self.prover.circuit.prove_gkr(self.out_claims, tr, self.api.MODE_PREFETCH)
        return tr

    def e2e(self):
        """BfvEncrypt::prove from HOST vectors (pinned): H2D of the witness vectors and of ct0is, circuit.evaluate, output claim,
        prove_gkr, proof bytes on the host"""
        return self.prover.prove_host(self.h_views, self.h_ct)[0]

    def close(self):
        self.prover.circuit.free()
        self.prover.lasso.free()
        self.ctx.close()


def run_slots(slots, fn_name, n):
    """every slot runs n proofs on its own host thread (the C calls release the GIL); returns when all are done"""
    if len(slots) == 1:
        f = getattr(slots[0], fn_name)
        for _ in range(n):
            f()
        return
    errs = []

    def work(s):
        try:
            f = getattr(s, fn_name)
            for _ in range(n):
                f()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ths = [threading.Thread(target=work, args=(s,)) for s in slots]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.lib()

    P, inp, bounds, segs, nv = make_case(args.config, seed=rank)
    ins, ct0is = LAST_WITNESS
    flat = [ins["s"], ins["e"], ins["k1"]] + list(ins["ais"]) + list(ins["r1is"]) + [ins["r2is"]]
    host_np = [np.array(v, dtype=np.uint64) for v in flat]
    n_in_elems = sum(v.size for v in host_np)
    ct_np = np.array(ct0is, dtype=np.uint64).reshape(-1)
    B = max(1, args.inflight)
    slots = [ProofSlot(api, np, torch, P, host_np, ct_np, local) for _ in range(B)]
    s0 = slots[0]
    ctx, prover, pp = s0.ctx, s0.prover, s0.prover.pp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_resident, step_e2e = s0.resident, s0.e2e

    # device spin-up (setup, untimed): the first ~20 proofs after process start run up to 12 % slower (first-touch of the work
    # buffers, lazily created kernels/attributes, challenge-chain cache); a prover service is measured in steady state
    spin_up = 0
    t_spin = time.perf_counter()
    while spin_up < 10 or time.perf_counter() - t_spin < 0.4:
        run_slots(slots, "resident", 1)
        spin_up += 1
    run_slots(slots, "resident", max(args.warmup, 3))
    tr = step_resident()
    proof_len = len(tr.into_proof())
    host_phases = prover.circuit.timing()
    l0 = ctx.launch_count
    step_resident()
    launches_per_proof = ctx.launch_count - l0

    # ---- single-proof latency (one proof at a time on this GPU), CUDA events on the library's stream
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    n_lat = max(10, args.steps // 2)
    for _ in range(n_lat):
        step_resident()
    e1.record(ext)
    barrier()
    latency_ms = e0.elapsed_time(e1) / n_lat

    # ---- timed region: exactly K steps; a step = one batch of B independent proofs, one per in-flight slot. Every slot ends each
    # proof with a stream synchronise, so the device is idle at both events; events on torch's current stream between two
    # device-wide synchronisations measure the device time of the whole region. Max over ranks.
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w0 = time.perf_counter()
    run_slots(slots, "resident", args.steps)
    torch.cuda.synchronize()
    e1.record()
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, wall * 1000.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, wall_ms = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms_total / 1000.0)

    # ---- end to end through the C ABI with HOST buffers (H2D of the inputs and D2H of the proof messages inside)
    run_slots(slots, "e2e", 3)
    barrier()
    w0 = time.perf_counter()
    run_slots(slots, "e2e", args.steps)
    barrier()
    e2e_wall = time.perf_counter() - w0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t[0])
    w0 = time.perf_counter()
    for _ in range(n_lat):
        step_e2e()
    e2e_latency_ms = 1000.0 * (time.perf_counter() - w0) / n_lat

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel class: CUDA events around every launch, separate pass right after the timed one
        ctx.profile(True)
        prof_steps = 3
        for _ in range(prof_steps):
            step_resident()
        prof = ctx.profile_read()
        ctx.profile(False)
        dom = max(prof.items(), key=lambda kv: kv[1][1])
        dn, (dl, dms, dby) = dom
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        achieved = (dby / 1e9) / (dms / 1e3) if dms > 0 else 0.0
        traffic = traffic_step = None
        try:   # ncu DRAM bytes of the same kernels (profiles/traffic.json, written by scripts/make_profiles.py from an ncu run of this command)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dn)
            traffic, traffic_step = tj["bytes_per_launch"], tj["bytes_per_step"]
        except Exception:
            pass
        total_ms = sum(v[1] for v in prof.values())
        roofline = {"bound": "hbm", "kernel": dn, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                    "traffic": traffic, "traffic_bytes_per_proof": traffic_step, "algorithmic_bytes_per_launch": dby / max(dl, 1),
                    "algorithmic_bytes_per_proof": dby / prof_steps, "peak_source": peak_src,
                    "launches_per_proof": dl / prof_steps, "avg_launch_us": 1000.0 * dms / max(dl, 1),
                    "share_of_proof_kernel_time": dms / total_ms if total_ms else None,
                    "per_class": {k: {"launches": v[0] / prof_steps, "ms": v[1] / prof_steps, "alg_GB": v[2] / prof_steps / 1e9,
                                      "GBps": (v[2] / 1e9) / (v[1] / 1e3) if v[1] > 0 else None} for k, v in prof.items()},
                    "whole_proof_alg_GB": sum(v[2] for v in prof.values()) / prof_steps / 1e9,
                    "whole_proof_frac": (sum(v[2] for v in prof.values()) / prof_steps / 1e9) / ((ms_total / (args.steps * B)) / 1e3) / peak}
        # ---- CPU baseline: the oracle port on this box's host cores, one full proof (N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import hgo
            hgo.build()
            hgo.set_num_threads(os.cpu_count() or 1)
            sess = hgo.bfv_prepare(0, P, ins, ct0is)
            t0 = time.perf_counter()
            oproof = sess.prove()
            dt = time.perf_counter() - t0
            same = oproof == step_e2e()
            cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": hgo.num_threads(), "kind": "port",
                   "sample": f"1 full proof of the same witness ({dt:.1f}s); GPU proof bytes == CPU proof bytes: {same}"}
        cfg = dict(workload_desc(args.config, P, nv, pp.num_memories), proofs_in_flight_per_gpu=B, spin_up_steps=spin_up,
                   witness="one synthetic witness per rank (seed = rank), proved by every slot of the rank")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "ms_per_proof": ms_total / (args.steps * B), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": value / PUBLISHED_PROOFS_PER_SEC if args.config == DEFAULT_CONFIG else None,
                "vs_baseline_note": "published: 5.06 s/proof on an Apple M1 (reference README.md:44, BASELINE.md); other hardware",
                "dtype": "u64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * ((n_in_elems + ct_np.size) * 8 + node_chal_bytes(nv) + 4096)),
                        "d2h_bytes_per_step": int(B * proof_len * 2), "single_proof_latency_ms": e2e_latency_ms},
                "gpu_launches": int(launches_per_proof * args.steps * B), "gpu_launches_per_proof": int(launches_per_proof),
                "single_proof_latency_ms": latency_ms, "wall_ms_per_step": wall_ms / args.steps, "proof_bytes": proof_len,
                "host_phases_us": host_phases, "roofline": roofline, "cpu_baseline": cpu}
    barrier()
    for s in slots:
        s.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def run_shard(args):
    """one LassoNode::prove_claim_reduction split over the ranks by grand-product terms. Every rank holds the
    node input on its device; per step each rank proves its shard, the message buffers are gathered to rank 0 over NCCL and
    summed in the field, rank 0 serialises. Time = wall clock of K steps between barriers, max over ranks (the step
    includes a host-side merge, so device events alone would miss part of it)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, inp, bounds, segs, nv = make_case(args.config, seed=0)      # the SAME witness on every rank
    ctx = api.Context(local)
    pp = api.LassoPreprocessing.preprocess(bounds)
    node = api.LassoNode(ctx, pp, nv, segs)
    buf = api.DeviceBuffer.from_numpy(ctx, inp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        tr = api.Keccak256Transcript()
        part = node.prove_shard(buf, tr, rank, world, n_inputs=inp.size)
        merged = api.gather_and_merge(api.GOLDILOCKS, part, None) if world > 1 else part
        if rank == 0:
            node.emit_shard(merged)
            return tr.into_proof()
        return None

    for _ in range(max(args.warmup, 3) + 20):   # + device spin-up, as in the main arm
        proof = step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    t = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    same = None
    if rank == 0:
        tr = api.Keccak256Transcript()
        node.prove_claim_reduction(buf, tr, api.MODE_PREFETCH, n_inputs=inp.size)
        same = tr.into_proof() == proof
        print(json.dumps({"metric": "lasso_node_sharded_proofs_per_sec", "value": args.steps / float(t[0]), "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1000.0 * float(t[0]) / args.steps,
                          "higher_is_better": True, "scaling": "strong", "dtype": "u64", "data": "synthetic",
                          "config": {"workload": f"ONE LassoNode::prove_claim_reduction (num_vars={nv}, memories={pp.num_memories}) split over {world} GPUs by "
                                                 "grand-product terms; message buffers gathered to rank 0 and summed in the field", "params": args.config},
                          "sharded_proof_equals_single_gpu_proof": same}))
    barrier()
    node.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def node_chal_bytes(nv, lm=16):
    gp = lambda k: sum(1 + ((1 + j) if j else 0) for j in range(k))
    return 16 * (2 * nv + 2 + gp(nv) + gp(lm))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=4, help="independent proofs in flight per GPU (a step = one batch of that many proofs)")
    ap.add_argument("--shard", action="store_true", help="extra measurement: one Lasso-node proof split over the N GPUs (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.shard:
        run_shard(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
