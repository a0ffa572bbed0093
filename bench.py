#!/usr/bin/env python3
"""bench.py — genome Mbp/s of the anchoring hot path (SML build + MemHash match find) on B200.

A "step" is one pass of the hot path over one batch of synthetic genomes: 2-bit pack, spaced-seed
extraction, LSD radix sort of the seed union (SML build), then multi-MUM finding with ungapped
extension.  Default workload = BASELINE.json configs[1]: 8 synthetic 5 Mbp genomes, weight-15
palindromic spaced seed (progressiveMauve default), 1 GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|small]

value  : Mbp/s with the ASCII genomes already resident in HBM when the timed region starts.
e2e    : Mbp/s through the public C-ABI call with pinned HOST buffers (H2D of the genomes and D2H of the
         MatchList inside the timed region).
N > 1  : launched by torchrun, one rank per GPU, SHARDED path (mems_find_matches_sharded): every rank extracts
         a block of the genomes, one NCCL all-to-all moves each seed range to its owner, a second one moves hits
         to the owner of their diagonal.  Weak scaling: the genome count stays 8 and every genome is N times
         longer (N = 1 is exactly the single-GPU workload), so each GPU carries a constant 40 Mbp.
--impl reference : times the UNMODIFIED reference (oracle/_ref, single-threaded MemorySML + MemHash) on
         the host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (n_genomes, length, seed weight, mode, description)
    "c1": (2, 5_000_000, 15, "memhash", "2 x 5 Mbp pairwise MUMs, w15 (BASELINE configs[0])"),
    "c2": (8, 5_000_000, 15, "memhash", "8 x 5 Mbp multi-MUM, w15 palindromic spaced seed (BASELINE configs[1])"),
    "c3": (1, 100_000_000, 19, "repeat", "1 x 100 Mbp RepeatHash, 200 families x 20 copies, w19 (BASELINE configs[2])"),
    "small": (4, 200_000, 15, "memhash", "4 x 0.2 Mbp (debug)"),
}
# bounded CPU samples of each workload (about 10-30 s of single-thread reference work)
CPU_SAMPLE = {"c1": (2, 2_000_000), "c2": (8, 500_000), "c3": (1, 10_000_000), "small": (4, 200_000)}
REF_STEP_SAMPLE = {"c1": (2, 500_000), "c2": (8, 150_000), "c3": (1, 3_000_000), "small": (4, 100_000)}


def make_genomes(name, n_genomes, length, seed):
    from libmems_b200 import synth
    if name in synth.BASELINE_WORKLOADS:  # the inputs the full-size parity tests pin (tests/golden/full_*.json)
        return synth.baseline_genomes(name, n_genomes, length, seed)
    return synth.genome_family(n_genomes, length, seed=seed)


class ClockSampler:
    """SM clock and throttle reasons while the timed regions run (B200_PROFILING.md's clocks line), sampled
    in-process through NVML every 20 ms: an `nvidia-smi -lms` child takes the driver lock while it starts up and
    stalled the first timed region by milliseconds."""

    def __init__(self, device):
        self.device, self.rows, self.stop_flag, self.t, self.nvml = device, [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()  # first query outside the timed region
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001 - NVML missing or refused: report it instead of failing the bench
            self.nvml, self.err = None, str(e)

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if self.device < len(ids) and ids[self.device].strip().isdigit():
                return int(ids[self.device])
        return self.device

    def _sample(self):
        n = self.nvml
        self.rows.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)),
                          int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))))

    def _loop(self):
        while not self.stop_flag.wait(0.02):
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                return

    def stop(self):
        if not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        n = self.nvml
        names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        try:
            n.nvmlShutdown()
        except Exception:  # noqa: BLE001
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "source": "NVML, 20 ms period, all three timed regions"}


def run_reference(args, name):
    """The reference's own CPU implementation (oracle/_ref) on a bounded sample, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from checkers import Reference
    n_genomes, length, weight, mode, desc = WORKLOADS[name]
    sg, sl = REF_STEP_SAMPLE[name]
    line = {"impl": "reference", "metric": "genome Mbp/s (SML build + MemHash match find)", "unit": "Mbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "genomes": n_genomes,
                       # our arm scales weakly: every genome is `gpus` times longer; the reference is timed on a bounded
                       # sample of that workload (cpu_baseline.sample), its Mbp/s does not depend on the length
                       "genome_length": length * (args.gpus if mode != "repeat" else 1), "seed_weight": weight,
                       "seed_pattern": None}}
    if not Reference.available():
        line["unavailable"] = "oracle/_ref/libmems_ref.so not present (needs /root/reference at build time)"
        print(json.dumps(line))
        return
    R = Reference()
    seed = R.get_seed(weight)
    line["config"]["seed_pattern"] = hex(seed)
    gs = make_genomes(name, sg, sl, seed=2)
    mbp = sum(len(g) for g in gs) / 1e6
    m = 1 if mode == "repeat" else 0
    for _ in range(args.warmup):
        R.find_matches(m, gs, seed)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, info = R.find_matches(m, gs, seed)
    dt = time.perf_counter() - t0
    v = mbp * args.steps / dt
    sample = "%d x %.2f Mbp per step (same generator and seed pattern as the full workload)" % (sg, sl / 1e6)
    line.update({"value": v, "ms_per_step": 1e3 * dt / args.steps,
                 "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": 1, "kind": "reference", "sample": sample,
                                  "sml_s": info["sml_s"], "find_s": info["find_s"]},
                 "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def cpu_baseline(name):
    from checkers import Reference
    n_genomes, length, weight, mode, desc = WORKLOADS[name]
    if not Reference.available():
        return {"value": None, "unit": "Mbp/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    R = Reference()
    sg, sl = CPU_SAMPLE[name]
    seed = R.get_seed(weight)
    gs = make_genomes(name, sg, sl, seed=2)
    t0 = time.perf_counter()
    _, info = R.find_matches(1 if mode == "repeat" else 0, gs, seed)
    dt = time.perf_counter() - t0
    mbp = sum(len(g) for g in gs) / 1e6
    return {"value": mbp / dt, "unit": "Mbp/s", "cores": 1, "kind": "reference",
            "sample": "%d x %.2f Mbp of the same synthetic family (one pass)" % (sg, sl / 1e6),
            "sml_build_mbp_s": mbp / info["sml_s"], "sml_s": info["sml_s"], "find_s": info["find_s"],
            "host_cores_available": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    name = args.workload
    if args.impl == "reference":
        return run_reference(args, name)

    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) are sent to stderr, the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import libmems_b200 as mems

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_genomes, length, weight, mode, desc = WORKLOADS[name]
    seed = mems.get_seed(weight)
    match_mode = mems.MODE_REPEAT if mode == "repeat" else mems.MODE_MEMHASH
    stream = torch.cuda.Stream()
    ctx = mems.Context(local_rank, stream=stream.cuda_stream)
    if world == 1:
        gs = make_genomes(name, n_genomes, length, seed=2)
        mbp_total = sum(len(g) for g in gs) / 1e6
        host = [torch.from_numpy(g).pin_memory() for g in gs]
        dev = [h.cuda(non_blocking=False) for h in host]

        def step(bufs):
            smls = ctx.create_smls([(b.data_ptr(), b.numel()) for b in bufs], seed)
            flat, info = ctx.find_matches(smls, mode=match_mode)
            for s in smls:
                s.close()
            return flat, info
    else:
        # sharded: same 8-genome family, every genome `world` times longer; this rank holds only its block
        if mode == "repeat":
            raise SystemExit("RepeatHash runs on one GPU (replicas only); use --gpus 1")
        length = length * world
        gs = make_genomes(name, n_genomes, length, seed=2)  # identical on every rank (seeded generator)
        lens = [len(g) for g in gs]
        mbp_total = sum(lens) / 1e6
        first, count = mems.shard_sequence_range(n_genomes, rank, world)
        host = [torch.from_numpy(g).pin_memory() if first <= i < first + count else None for i, g in enumerate(gs)]
        dev = [h.cuda(non_blocking=False) if h is not None else None for h in host]
        del gs
        uid = [mems.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=torch.device("cuda", local_rank))
        comm = ctx.create_comm(uid[0], rank, world)

        def step(bufs):
            seqs = [(b.data_ptr(), b.numel()) if b is not None else None for b in bufs]
            return ctx.find_matches_sharded(comm, seqs, lens, seed, mode=match_mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(bufs, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        info = None
        for _ in range(steps):
            flat, info = step(bufs)
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, flat, info

    for _ in range(args.warmup):
        step(dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one poller per node: concurrent nvidia-smi loops contend for the driver and stall the ranks
        sampler.start()
    # timed region 1: inputs resident in HBM -> `value`
    ctx.profile_reset()
    ms_dev, flat, info = timed(dev, args.steps)
    launches = ctx.launch_count()
    # timed region 2: the same steps with every launch bracketed by CUDA events on the launching stream
    # (per-kernel durations for the roofline; the event records add host overhead, so it is not `value`)
    ctx.profile_enable(True)
    ms_prof, _, _ = timed(dev, args.steps)
    prof = ctx.profile()
    ctx.profile_enable(False)
    # timed region 3: pinned host buffers in, MatchList out -> `e2e`
    for _ in range(2):
        step(host)
    ms_e2e, flat_h, info_h = timed(host, args.steps)
    clocks = sampler.stop()
    rank_kernel_ms = None
    if world > 1:
        # per-kernel device time of every rank (ms per step): shows which rank / kernel sets the max-over-ranks time
        mine = {k: v["ms"] / args.steps for k, v in prof.items()}
        rank_kernel_ms = [None] * world
        dist.all_gather_object(rank_kernel_ms, mine)
    n_matches_total, n_hits_total, d2h_total = info["n_matches"], info["n_hits"], len(flat_h) * 8 + 64
    if world > 1:
        t = torch.tensor([n_matches_total, n_hits_total, d2h_total], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        n_matches_total, n_hits_total, d2h_total = (int(x) for x in t.tolist())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        total_ms = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                       "share": v["ms"] / total_ms,
                       "gbs": (v["bytes"] / v["ms"] / 1e6) if v["ms"] > 0 and v["bytes"] > 0 else None}
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        # dominant kernel = largest share of device time among the launches of the timed region
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        hbm_kernels = {k: v for k, v in prof.items() if v["bytes"] > 0}
        roof_name, roof = max(hbm_kernels.items(), key=lambda kv: kv[1]["ms"])
        achieved = roof["bytes"] / roof["ms"] / 1e6
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/), if it is the
        # same kernel and workload
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if roof_name == "radix_pass" and name == "c2" and world == 1:
                traffic = tj["onesweep_kernel<u32>"]["dram_bytes_per_launch"]
        except (OSError, KeyError):
            pass
        # the SML-build stage as a whole (pack + extract + all radix passes), algorithmic bytes per DESIGN.md §4
        sml = [prof[k] for k in ("pack", "extract", "radix_pass") if k in prof]
        sml_ms, sml_bytes = sum(v["ms"] for v in sml), sum(v["bytes"] for v in sml)
        roofline = {"bound": "hbm", "kernel": roof_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic,
                    "sml_build": {"ms_per_step": sml_ms / args.steps, "achieved": sml_bytes / sml_ms / 1e6 if sml_ms else None,
                                  "frac": sml_bytes / sml_ms / 1e6 / hbm_peak if sml_ms else None,
                                  "mbp_per_s": mbp_total / world / (sml_ms / args.steps / 1e3) if sml_ms else None},
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "launches": roof["launches"], "avg_launch_ms": roof["ms"] / max(roof["launches"], 1),
                    "dominant_kernel_by_time": top[0], "kernel_ms_per_step": total_ms / args.steps,
                    "profiled_ms_per_step": ms_prof / args.steps}
        line = {
            "metric": "genome Mbp/s (SML build + MemHash match find)",
            "value": mbp_total * args.steps / (ms_dev / 1e3), "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32" if 2 * weight + 1 <= 32 else "u64", "data": "synthetic",
            "config": {"workload": desc, "genomes": n_genomes, "genome_length": length, "seed_weight": weight,
                       "seed_pattern": hex(seed), "multi_gpu": ("sharded: genome blocks per rank, seed-range all-to-all + diagonal all-to-all over NCCL; "
                                     "weak scaling by genome length") if world > 1 else "n/a",
                       "l2": "per-step working set (%.0f MB of seed records per GPU) exceeds the 126 MB L2" %
                             (mbp_total / world * (8 if 2 * weight + 1 <= 32 else 12))},
            "e2e": {"value": mbp_total * args.steps / (ms_e2e / 1e3), "unit": "Mbp/s",
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(mbp_total * 1e6),
                    "d2h_bytes_per_step": int(d2h_total)},
            "gpu_launches": launches, "matches_per_step": n_matches_total, "hits_per_step": n_hits_total,
            "matches_per_s": n_matches_total * args.steps / (ms_dev / 1e3),
            "roofline": roofline, "kernels": kernels, "clocks": clocks,
        }
        if rank_kernel_ms:
            names = sorted({k for d in rank_kernel_ms for k in d})
            line["kernel_ms_max_over_ranks"] = {k: max(d.get(k, 0.0) for d in rank_kernel_ms) for k in names}
            line["kernel_ms_total_per_rank"] = [sum(d.values()) for d in rank_kernel_ms]
            try:
                # the seed-record exchange against NVLink 5 (900 GB/s per direction and GPU): bytes this rank sends to
                # its peers, (world - 1) / world of its records, over the duration of the exchange incl. its barrier
                ex = next((v for k, v in prof.items() if k.endswith("all_to_all_records")), None)
                if ex and ex["ms"] > 0:
                    sent = ex["bytes"] * (world - 1) / world
                    gbs = sent / ex["ms"] / 1e6
                    line["exchange"] = {"kernel": next(k for k in prof if k.endswith("all_to_all_records")),
                                        "sent_bytes_per_step": sent / args.steps, "ms_per_step": ex["ms"] / args.steps,
                                        "achieved": gbs, "peak": 900.0, "unit": "GB/s", "frac": gbs / 900.0,
                                        "note": "rank 0; the time includes waiting for the slowest rank at the barrier"}
            except Exception as e:  # noqa: BLE001 - never let a report field break the bench line
                line["exchange"] = {"error": str(e)}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(name)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
