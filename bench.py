#!/usr/bin/env python3
"""bench.py — genome Mbp/s of the anchoring hot path (SML build + MemHash match find) on B200.

A "step" is one pass of the hot path over one batch of synthetic genomes: 2-bit pack, spaced-seed
extraction, LSD radix sort of the seed union (SML build), then multi-MUM finding with ungapped
extension.  Default workload = BASELINE.json configs[1]: 8 synthetic 5 Mbp genomes, weight-15
palindromic spaced seed (progressiveMauve default), 1 GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|small]
                  [--no-extras] [--no-cpu-baseline]

value  : Mbp/s with the ASCII genomes already resident in HBM when the timed region starts.
e2e    : Mbp/s through the public C-ABI call with pinned HOST buffers (H2D of the genomes and D2H of the
         MatchList inside the timed region); e2e.h2d_ms / d2h_ms / kernel_ms come from one more pass of the same
         call with every copy and launch bracketed by CUDA events.
N > 1  : launched by torchrun, one rank per GPU, SHARDED path (mems_find_matches_sharded): every rank extracts
         a block of the genomes, one all-to-all moves each seed range to its owner, a second one moves hits
         to the owner of their diagonal.  Weak scaling: the genome count stays 8 and every genome is N times
         longer (N = 1 is exactly the single-GPU workload), so each GPU carries a constant 40 Mbp.  Before anything
         is timed the union of the ranks' MatchLists is compared with a single-GPU run of the same input on rank 0
         (`parity_check`).  `extra` carries the other BASELINE configs that fit the launch: configs 1 and 3 and the
         64-bit-key sort at N = 1, one seed weight of config 4 (50 x 5 Mbp) at N = 2/4, config 5 (16 x 100 Mbp) at N = 8.
--impl reference : times the UNMODIFIED reference (oracle/_ref: MemorySML + MemHash, single-threaded code) on the
         host cores over a bounded sample of the same workload — one independent instance per host core, since the
         reference path itself has no threads — and says in `config` what the sample is.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "genome Mbp/s (SML build + MemHash match find)"
WORKLOADS = {
    # name: (n_genomes, length, seed weight, mode, description)
    "c1": (2, 5_000_000, 15, "memhash", "2 x 5 Mbp pairwise MUMs, w15 (BASELINE configs[0])"),
    "c2": (8, 5_000_000, 15, "memhash", "8 x 5 Mbp multi-MUM, w15 palindromic spaced seed (BASELINE configs[1])"),
    "c3": (1, 100_000_000, 19, "repeat", "1 x 100 Mbp RepeatHash, 200 families x 20 copies, w19 (BASELINE configs[2])"),
    "small": (4, 200_000, 15, "memhash", "4 x 0.2 Mbp (debug)"),
}
# bounded CPU samples of each workload: cpu_baseline leg (one pass) and reference arm (per step and worker)
CPU_SAMPLE = {"c1": (2, 2_000_000), "c2": (8, 500_000), "c3": (1, 10_000_000), "small": (4, 200_000)}
REF_STEP_SAMPLE = {"c1": (2, 500_000), "c2": (8, 150_000), "c3": (1, 3_000_000), "small": (4, 100_000)}


def make_genomes(name, n_genomes, length, seed):
    from libmems_b200 import synth
    if name in synth.BASELINE_WORKLOADS:  # the inputs the full-size parity tests pin (tests/golden/full_*.json)
        return synth.baseline_genomes(name, n_genomes, length, seed)
    return synth.genome_family(n_genomes, length, seed=seed)


class ClockSampler:
    """SM clock and throttle reasons while the timed regions run (B200_PROFILING.md's clocks line), read through NVML by
    the thread that drives the GPU, between the SML build call (which returns while the GPU is still sorting) and the
    match-finding call: the GPU is under load, and no CUDA call of ours is in flight.  A polling thread (or an
    `nvidia-smi -lms` child) contends with the CUDA calls of the steps for the driver's locks: at a 20 ms period it
    stalled single steps by up to a second."""

    def __init__(self, device):
        self.device, self.rows, self.nvml, self.last = device, [], None, 0.0
        self.period = float(os.environ.get("MEMS_BENCH_CLOCK_PERIOD", "0.02"))

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()  # first query outside the timed region
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

    def tick(self):
        """Called by the step function while the GPU works: one sample if the last one is older than the period."""
        if self.nvml is None:
            return
        now = time.perf_counter()
        if now - self.last >= self.period:
            self.last = now
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                self.nvml = None

    def stop(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        n = self.nvml or __import__("pynvml")
        names = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        try:
            n.nvmlShutdown()
        except Exception:  # noqa: BLE001
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "source": "NVML, read by the driving thread while the GPU works on a step (at most every %g ms), "
                                              "warm-up and all timed regions of the headline workload" % (1e3 * self.period)}


# ------------------------------------------------------------------------------------------------ reference arm
_REF = {}


def _ref_worker_init():
    from checkers import Reference
    _REF["R"] = Reference()


def _ref_worker_step(job):
    name, sg, sl, gen_seed, seed, mode = job
    gs = _REF.get("gs_%d" % gen_seed)
    if gs is None:
        gs = _REF["gs_%d" % gen_seed] = make_genomes(name, sg, sl, seed=gen_seed)
    _, info = _REF["R"].find_matches(mode, gs, seed)
    return info["sml_s"], info["find_s"]


def run_reference(args, name):
    """The reference's own CPU implementation (oracle/_ref) on a bounded sample, rank 0 only: one independent
    MemorySML + MemHash instance per host core (the reference path is single-threaded code)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from checkers import Reference
    n_genomes, length, weight, mode, desc = WORKLOADS[name]
    sg, sl = REF_STEP_SAMPLE[name]
    cores = max(1, min(os.cpu_count() or 1, 64)) if not args.ref_cores else args.ref_cores
    sample = ("%d independent instances (one per host core), each %d x %.2f Mbp per step of the same synthetic family "
              "and seed pattern as the full workload" % (cores, sg, sl / 1e6))
    line = {"impl": "reference", "metric": METRIC, "unit": "Mbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            # what this arm really runs: a BOUNDED SAMPLE of the GPU arm's workload (the reference needs ~190 s for one
            # full 8 x 5 Mbp step and gets slower per Mbp as the genomes grow, so the sample flatters it)
            "config": {"workload": "bounded sample of: " + desc, "genomes": sg, "genome_length": sl, "seed_weight": weight,
                       "seed_pattern": None, "instances": cores,
                       "full_workload": {"genomes": n_genomes, "genome_length": length * (args.gpus if mode != "repeat" else 1)}}}
    if not Reference.available():
        line["unavailable"] = "oracle/_ref/libmems_ref.so not present (needs /root/reference at build time)"
        print(json.dumps(line))
        return
    seed = Reference().get_seed(weight)
    line["config"]["seed_pattern"] = hex(seed)
    m = 1 if mode == "repeat" else 0
    jobs = [(name, sg, sl, 2 + i, seed, m) for i in range(cores)]
    mbp = cores * sg * sl / 1e6  # (indels change a genome's length by < 0.1 %)
    with mp.get_context("fork").Pool(cores, initializer=_ref_worker_init) as pool:
        for _ in range(max(args.warmup, 1)):  # the first pass also generates each worker's genomes
            pool.map(_ref_worker_step, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            times = pool.map(_ref_worker_step, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    v = mbp * args.steps / dt
    line.update({"value": v, "ms_per_step": 1e3 * dt / args.steps,
                 "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": cores, "kind": "reference", "sample": sample,
                                  "sml_s": float(np.mean([t[0] for t in times])), "find_s": float(np.mean([t[1] for t in times])),
                                  "one_instance_mbp_s": sg * sl / 1e6 / float(np.mean([t[0] + t[1] for t in times]))},
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
            "sample": "%d x %.2f Mbp of the same synthetic family (one pass, one instance)" % (sg, sl / 1e6),
            "sml_build_mbp_s": mbp / info["sml_s"], "sml_s": info["sml_s"], "find_s": info["find_s"],
            "host_cores_available": os.cpu_count()}


# ------------------------------------------------------------------------------------------------ helpers
def records_digest(flats):
    """(n, n distinct, sha256) of the canonically sorted union of flat [SeqCount, Length, starts...] record arrays."""
    import libmems_b200 as mems
    from libmems_b200 import synth
    recs = sorted(m for f in flats for m in mems.flat_to_matches(f))
    return len(recs), len(set(recs)), synth.matchlist_digest(recs)


def traffic_from_profiles(kernel_tag):
    """DRAM bytes per launch from this round's committed ncu capture — only while the kernel's source is unchanged."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj[kernel_tag]
        src = open(os.path.join(ROOT, ent["source"]), "rb").read()
        if hashlib.sha256(src).hexdigest() != ent["source_sha256"]:
            return None
        return ent["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--ref-cores", type=int, default=0)
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
    from libmems_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # The host thread drives ~60 launches and a handful of waits per step.  MEMS_BENCH_PRIO=nice / fifo ask the scheduler
    # to favour this process (an experiment: it made no difference on the pool's boxes); the default leaves it alone.
    prio = os.environ.get("MEMS_BENCH_PRIO", "none")
    host_priority = "unchanged"
    try:
        if prio == "fifo":
            os.sched_setscheduler(0, os.SCHED_FIFO, os.sched_param(10))
            host_priority = "SCHED_FIFO 10"
        elif prio == "nice":
            os.setpriority(os.PRIO_PROCESS, 0, -20)
            host_priority = "nice -20"
    except (OSError, AttributeError) as e:
        host_priority = "unchanged (%s)" % type(e).__name__
    # run this rank on the CPUs next to its GPU: the page-locked buffers of the copies are then allocated on the GPU's
    # NUMA node (eight ranks whose buffers sit behind the other socket share one inter-socket link)
    numa = "unchanged"
    try:
        if world == 1 and not os.environ.get("MEMS_BENCH_BIND"):
            raise RuntimeError("single GPU: not bound")
        import pynvml
        pynvml.nvmlInit()
        vis = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
        phys = int(vis[local_rank]) if local_rank < len(vis) else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        numa = "bound to the GPU's CPU set (%d CPUs)" % len(os.sched_getaffinity(0))
    except Exception as e:  # noqa: BLE001
        numa = "unchanged (%s)" % type(e).__name__
    n_genomes, length, weight, mode, desc = WORKLOADS[name]
    seed = mems.get_seed(weight)
    match_mode = mems.MODE_REPEAT if mode == "repeat" else mems.MODE_MEMHASH
    stream = torch.cuda.Stream()
    ctx = mems.Context(local_rank, stream=stream.cuda_stream)
    comm = None
    if world > 1:
        uid = [mems.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=torch.device("cuda", local_rank))
        comm = ctx.create_comm(uid[0], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, bufs, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        info, pend, walls = None, None, []
        for _ in range(steps):
            t0 = time.perf_counter()
            # The records of a step leave the device behind its last kernel (mems_b200.h, mems_matches_wait): the previous
            # step's MatchList is given back only after this step's call has been issued, so its D2H copy overlaps this
            # step's kernels.  Every copy lies inside the timed region: the last one is waited for before the end mark.
            prev = pend
            pend, info = step(bufs)
            prev = None  # noqa: F841 - waits for its records, then returns its page-locked buffer to the library's pool
            walls.append(1e3 * (time.perf_counter() - t0))
        pend.wait()
        b.record(stream)
        flat = pend.records()
        pend = None
        print("[bench r%d] host ms per step: %s" % (rank, " ".join("%.2f" % w for w in walls)), file=sys.stderr)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, flat, info

    sampler = ClockSampler(local_rank)

    def single_step(m, sd):
        def step(bufs):
            smls = ctx.create_smls([(b.data_ptr(), b.numel()) for b in bufs], sd)
            sampler.tick()  # the GPU is sorting; no CUDA call of this process is in flight
            pend, info = ctx.find_matches(smls, mode=m, wait=False)
            for s in smls:
                s.close()
            return pend, info
        return step

    parity = None
    if world == 1:
        gs = make_genomes(name, n_genomes, length, seed=2)
        mbp_total = sum(len(g) for g in gs) / 1e6
        host = [torch.from_numpy(g).pin_memory() for g in gs]
        dev = [h.cuda(non_blocking=False) for h in host]
        step = single_step(match_mode, seed)
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

        def step(bufs):
            seqs = [(b.data_ptr(), b.numel()) if b is not None else None for b in bufs]
            res = ctx.find_matches_sharded(comm, seqs, lens, seed, mode=match_mode, wait=False)
            sampler.tick()  # (one collective call per step: the sample falls between two steps)
            return res

        # ---- parity of the sharded MatchList, before anything is timed: the union of the ranks' shares against a
        # single-GPU run of the same input on rank 0 (both as canonically sorted record lists)
        pend, info = step(dev)
        flat = pend.records()
        shares = [None] * world
        dist.all_gather_object(shares, np.asarray(flat).copy())
        if rank == 0:
            n_sh, n_sh_distinct, dig_sh = records_digest(shares)
            c1 = mems.Context(local_rank)
            sm = c1.create_smls(gs, seed)
            f1, i1 = c1.find_matches(sm, mode=match_mode, order=mems.ORDER_CANONICAL)
            n_1, _, dig_1 = records_digest([f1])
            for s in sm:
                s.close()
            del f1, sm
            c1.close()
            parity = {"n_matches": n_sh, "distinct": n_sh_distinct, "digest": dig_sh, "single_gpu_n_matches": n_1,
                      "single_gpu_digest": dig_1, "equal": bool(dig_sh == dig_1 and n_sh == n_sh_distinct == n_1),
                      "how": "sha256 of the canonically sorted union of all ranks' records vs mems_find_matches on rank 0 over the same genomes"}
            print("[bench] sharded parity:", parity, file=sys.stderr)
        del shares, flat, pend
        dist.barrier()
        del gs

    if rank == 0:  # one reader per node
        sampler.start()  # before the warm-up: NVML's start-up cost must not land in a timed region
    # warm-up with the same buffer lifetime pattern as the timed regions (one MatchList alive outside, one per step):
    # the first cudaHostAlloc of a result buffer costs milliseconds
    flat, _ = step(dev)
    timed(step, dev, args.warmup)
    # timed region 1: inputs resident in HBM -> `value`
    ctx.profile_reset()
    ms_dev, flat, info = timed(step, dev, args.steps)
    launches = ctx.launch_count()
    # timed region 2: the same steps with every launch bracketed by CUDA events on the launching stream
    # (per-kernel durations for the roofline; the event records add host overhead, so it is not `value`)
    ctx.profile_enable(True)
    ms_prof, _, _ = timed(step, dev, args.steps)
    prof = ctx.profile()
    ctx.profile_enable(False)
    # timed region 3: pinned host buffers in, MatchList out -> `e2e`
    timed(step, host, 2)
    ms_e2e, flat_h, info_h = timed(step, host, args.steps)
    # region 4 (short): the e2e call again with copies and launches bracketed by events -> where e2e's time goes
    ctx.profile_reset()
    ctx.profile_enable(True)
    n_split = max(2, min(args.steps, 5))
    ms_e2e_prof, _, _ = timed(step, host, n_split)
    prof_e2e = ctx.profile()
    ctx.profile_enable(False)
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

    # ---- the other BASELINE configs that fit this launch (short runs; each gets its own sub-line)
    extra = {}
    if not args.no_extras and name == "c2":
        flat = flat_h = None
        host = dev = None
        try:
            extra = run_extras(mems, synth, ctx, comm, rank, world, timed, single_step)
        except Exception as e:  # noqa: BLE001 - an extra must never cost the headline line
            extra = {"error": repr(e)}

    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        is_copy = lambda k: k.startswith("copy_")  # noqa: E731
        kern = {k: v for k, v in prof.items() if not is_copy(k)}
        total_ms = sum(v["ms"] for v in kern.values()) or 1.0
        kernels = {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                       "share": v["ms"] / total_ms,
                       "gbs": (v["bytes"] / v["ms"] / 1e6) if v["ms"] > 0 and v["bytes"] > 0 else None}
                   for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
        # dominant kernel = largest share of device time among the launches of the timed region
        top = max(kern.items(), key=lambda kv: kv[1]["ms"])
        hbm_kernels = {k: v for k, v in kern.items() if v["bytes"] > 0 and not k.startswith(("peer_", "nccl_"))}
        roof_name, roof = max(hbm_kernels.items(), key=lambda kv: kv[1]["ms"])
        achieved = roof["bytes"] / roof["ms"] / 1e6
        key_tag = "u32" if 2 * weight + 1 <= 32 else "u64"
        traffic = None
        if roof_name == "radix_pass" and name == "c2" and world == 1:
            traffic = traffic_from_profiles("onesweep_kernel<%s>" % key_tag)
        # the SML-build stage as a whole (pack + planes + extract + all radix passes), algorithmic bytes per DESIGN.md §4
        sml = [prof[k] for k in ("pack", "planes", "extract", "radix_pass") if k in prof]
        sml_ms, sml_bytes = sum(v["ms"] for v in sml), sum(v["bytes"] for v in sml)
        walk_ms = sum(v["ms"] for k, v in kern.items() if "walk" in k) / args.steps
        roofline = {"bound": "hbm", "kernel": roof_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic,
                    "sml_build": {"ms_per_step": sml_ms / args.steps, "achieved": sml_bytes / sml_ms / 1e6 if sml_ms else None,
                                  "frac": sml_bytes / sml_ms / 1e6 / hbm_peak if sml_ms else None,
                                  "mbp_per_s": mbp_total / world / (sml_ms / args.steps / 1e3) if sml_ms else None},
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                    "launches": roof["launches"], "avg_launch_ms": roof["ms"] / max(roof["launches"], 1),
                    "dominant_kernel_by_time": top[0], "kernel_ms_per_step": total_ms / args.steps,
                    "extension_ms_per_step": walk_ms, "profiled_ms_per_step": ms_prof / args.steps}
        e2e_kernel_ms = sum(v["ms"] for k, v in prof_e2e.items() if not is_copy(k)) / n_split
        line = {
            "metric": METRIC,
            "value": mbp_total * args.steps / (ms_dev / 1e3), "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": key_tag, "data": "synthetic",
            "config": {"workload": desc, "genomes": n_genomes, "genome_length": length, "seed_weight": weight,
                       "seed_pattern": hex(seed), "multi_gpu": ("sharded: genome blocks per rank, seed-range all-to-all + diagonal all-to-all "
                                     "(peer-to-peer over NVLink); weak scaling by genome length") if world > 1 else "n/a",
                       "host_affinity": numa, "host_priority": host_priority,
                       "l2": "per-step working set (%.0f MB of seed records per GPU) exceeds the 126 MB L2" %
                             (mbp_total / world * (8 if 2 * weight + 1 <= 32 else 12))},
            "e2e": {"value": mbp_total * args.steps / (ms_e2e / 1e3), "unit": "Mbp/s",
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(mbp_total * 1e6),
                    "d2h_bytes_per_step": int(d2h_total),
                    # rank 0, from the event-bracketed pass: time of the H2D copies of the genomes, of the D2H copy of the
                    # MatchList, and of all kernels; the rest of ms_per_step is launch gaps and host round trips
                    "h2d_ms": prof_e2e.get("copy_in_sequences", {}).get("ms", 0.0) / n_split,
                    "d2h_ms": prof_e2e.get("copy_out_matches", {}).get("ms", 0.0) / n_split,
                    "kernel_ms": e2e_kernel_ms, "bracketed_ms_per_step": ms_e2e_prof / n_split,
                    "delivery": "every step's MatchList is copied to page-locked host memory inside the timed region; the copy "
                                "of step i runs on the library's copy stream while step i+1's kernels run (the records are waited "
                                "for when the caller asks for them, mems_matches_wait); the bracketed pass copies in line"},
            "gpu_launches": launches, "matches_per_step": n_matches_total, "hits_per_step": n_hits_total,
            "matches_per_s": n_matches_total * args.steps / (ms_dev / 1e3),
            "roofline": roofline, "kernels": kernels, "clocks": clocks,
        }
        if parity is not None:
            line["parity_check"] = parity
        if extra:
            line["extra"] = extra
        if rank_kernel_ms:
            names = sorted({k for d in rank_kernel_ms for k in d})
            line["kernel_ms_max_over_ranks"] = {k: max(d.get(k, 0.0) for d in rank_kernel_ms) for k in names}
            line["kernel_ms_total_per_rank"] = [sum(v for k, v in d.items() if not is_copy(k)) for d in rank_kernel_ms]
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
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_extras(mems, synth, ctx, comm, rank, world, timed, single_step):
    """Short runs of the other BASELINE configs (device-resident inputs, a few warm-up and timed steps each)."""
    import torch
    out = {}
    W, K = 3, 5
    hbm_peak = float(measured_peaks().get("hbm_gbs", 6650.0))
    if world == 1:
        for cfg in ("c1", "c3"):
            g, n, w, mode, desc = WORKLOADS[cfg]
            sd = mems.get_seed(w)
            gs = make_genomes(cfg, g, n, seed=2)
            mbp = sum(len(x) for x in gs) / 1e6
            dev = [torch.from_numpy(x).cuda() for x in gs]
            del gs
            step = single_step(mems.MODE_REPEAT if mode == "repeat" else mems.MODE_MEMHASH, sd)
            timed(step, dev, W)
            ms, flat, info = timed(step, dev, K)
            ctx.profile_reset()
            ctx.profile_enable(True)
            timed(step, dev, 2)
            prof = ctx.profile()
            ctx.profile_enable(False)
            kern_ms = sum(v["ms"] for k, v in prof.items() if not k.startswith("copy_")) / 2
            sub = {"workload": desc, "value": mbp * K / (ms / 1e3), "unit": "Mbp/s", "ms_per_step": ms / K, "steps": K, "warmup": W,
                   "matches_per_step": info["n_matches"], "hits_per_step": info["n_hits"], "kernel_ms_per_step": kern_ms,
                   "host_replay_ms_per_step": info["host_replay_ms"]}
            if mode == "repeat":
                sub["note"] = "RepeatHash is returned in the reference's table order: the table replay runs on host threads"
            rp = prof.get("radix_pass")
            if rp and rp["ms"] > 0:
                gbs = rp["bytes"] / rp["ms"] / 1e6
                sub["radix_pass"] = {"key": "u64" if 2 * w + 1 > 32 else "u32", "avg_launch_ms": rp["ms"] / rp["launches"],
                                     "achieved": gbs, "unit": "GB/s", "peak": hbm_peak, "frac": gbs / hbm_peak}
            out[cfg] = sub
            del dev, flat
        return out
    # ---- sharded extras.  Every rank generates only its own block: genome 0 is the shared base, genome g its mutation
    # under seed + g (tools/run_sharded.py does the same)
    import torch.distributed as dist

    def sharded_case(n_genomes, length, weight, steps, warm):
        sd = mems.get_seed(weight)
        first, count = mems.shard_sequence_range(n_genomes, rank, world)
        base = synth.random_genome(length, np.random.default_rng(12345))
        mine = {g: (base if g == 0 else synth.mutate(base, np.random.default_rng(12345 + g))) for g in range(first, first + count)}
        del base
        lens_t = torch.zeros(n_genomes, dtype=torch.int64, device="cuda")
        for g, a in mine.items():
            lens_t[g] = len(a)
        dist.all_reduce(lens_t)
        lens = [int(x) for x in lens_t.tolist()]
        dev = {g: torch.from_numpy(a).cuda() for g, a in mine.items()}
        del mine
        bufs = [dev.get(g) for g in range(n_genomes)]

        def step(bufs):
            seqs = [(b.data_ptr(), b.numel()) if b is not None else None for b in bufs]
            return ctx.find_matches_sharded(comm, seqs, lens, sd, wait=False)

        timed(step, bufs, warm)
        ms, flat, info = timed(step, bufs, steps)
        ctx.profile_reset()
        ctx.profile_enable(True)
        timed(step, bufs, 1)
        prof = ctx.profile()
        ctx.profile_enable(False)
        kern_ms = sum(v["ms"] for k, v in prof.items() if not k.startswith("copy_"))
        t = torch.tensor([info["n_matches"], info["n_hits"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        mx = torch.tensor([kern_ms, float(info["max_run"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mbp = sum(lens) / 1e6
        return {"value": mbp * steps / (ms / 1e3), "unit": "Mbp/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warm,
                "total_mbp": mbp, "matches_per_step": int(t[0]), "hits_per_step": int(t[1]),
                "kernel_ms_per_step_max_over_ranks": float(mx[0]), "max_run": int(mx[1]), "seed_pattern": hex(sd),
                "rank0_kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}}

    if world == 8:
        sub = sharded_case(16, 100_000_000, 19, 3, 2)
        sub["workload"] = "16 x 100 Mbp, w19, SML build + multi-MUM finding sharded over 8 GPUs (BASELINE configs[4])"
        out["c5"] = sub
    else:
        sub = sharded_case(50, 5_000_000, 15, 3, 2)
        sub["workload"] = "50 x 5 Mbp, requested weight 15, sharded by seed range over %d GPUs (one point of BASELINE configs[3])" % world
        out["c4_w15"] = sub
    return out


if __name__ == "__main__":
    main()
