#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-base operator pipeline.

Workload (BASELINE.json configs[1]): hg38-shaped genome (24 chromosomes,
3,088,269,832 bases, fp64) with synthetic coverage (reads of length U{50..150},
mean depth 5); one step = depth accumulation of all intervals followed by
`smooth --window=101`.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

Prints ONE JSON line (see the contract in the task description).  The timed
region is bracketed by a barrier + torch.cuda.synchronize() on both sides, timed
with CUDA events on the stream the kernels run on, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
        ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
        ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
        ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
        ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]

WINDOW = 101
DEPTH = 5
READ_MIN, READ_MAX = 50, 150
FP64_NOMINAL = 148 * 64 * 1.965e9      # FP64 lanes * SMs * max SM clock (instructions/s)
FP64_MEASURED = 1.745e13               # scripts/micro/fp64_peak.cu on this pool's B200 (gpurun, round 1)
SMOOTH_TRAFFIC_BYTES_PER_BASE = 15.99   # (24.71 GB read + 24.67 GB written) / 3,088,269,832 bases: ncu --set full, profiles/r1_bench_kernels_ncu_full.csv
METRIC = "Gbp/s, hg38 depth accumulation + smooth --window=101 (fp64)"


def scaled_genome(scale):
    return [(n, max(1000, l // scale)) for n, l in HG38]


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- synthetic data
def synth_intervals(torch, device, sorted_chroms, seed=20261018):
    """cfg2 coverage: per chromosome round(len*5/100) reads, start uniform,
    length U{50..150}.  Returns (chrom_sorted_index, start, end) int32 device tensors."""
    segs, starts, ends = [], [], []
    for si, (name, length) in enumerate(sorted_chroms):
        m = int(round(length * DEPTH / 100.0))
        if m == 0:
            continue
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + si)
        span = max(1, length - READ_MAX)
        s = torch.randint(0, span, (m,), generator=gen, device=device, dtype=torch.int64)
        ln = torch.randint(READ_MIN, READ_MAX + 1, (m,), generator=gen, device=device, dtype=torch.int64)
        e = torch.clamp(s + ln, max=length)
        segs.append(torch.full((m,), si, dtype=torch.int32, device=device))
        starts.append(s.to(torch.int32)); ends.append(e.to(torch.int32))
        del s, ln, e
    return torch.cat(segs), torch.cat(starts), torch.cat(ends)


# ----------------------------------------------------------------------------- CPU baseline
def _cpu_sample_worker(args):
    """one bounded sample of the workload on one core: accumulate (the array form of the
    reference loop genodsp.c:1325-1329, oracle port -- the reference itself only ingests text)
    then `smooth --window=101` by the reference's own op_smooth_apply (oracle/_ref) when built."""
    n, seed = args
    import numpy as np
    from checkers import Oracle, RefGenome, have_ref
    rng = np.random.default_rng(seed)
    m = int(round(n * DEPTH / 100.0))
    s = rng.integers(0, n - READ_MAX, m).astype(np.uint32)
    e = (s + rng.integers(READ_MIN, READ_MAX + 1, m)).astype(np.uint32)
    orc = Oracle()
    if have_ref():
        g = RefGenome([("chrS", n)])
        v = g.vec["chrS"]
        t0 = time.perf_counter()
        orc.accumulate(v, s, e)
        t1 = time.perf_counter()
        g.apply("smooth", "--window=%d" % WINDOW)
        t2 = time.perf_counter()
        g.close()
        kind = "reference"
    else:
        v = np.zeros(n)
        t0 = time.perf_counter()
        orc.accumulate(v, s, e)
        t1 = time.perf_counter()
        orc.smooth(v, WINDOW)
        t2 = time.perf_counter()
        kind = "port"
    return (t1 - t0, t2 - t1, kind)


def cpu_baseline_single(n=160_000_000):        # 10-20 s of single-core work
    acc, smo, kind = _cpu_sample_worker((n, 1))
    tot = acc + smo
    return {"value": n / tot / 1e9, "unit": "Gbp/s", "cores": 1, "kind": kind,
            "sample": "one %d-base chromosome, depth %d (%d intervals): accumulate %.2fs (oracle array loop) + "
                      "smooth W=%d %.2fs (%s op_smooth_apply)" % (n, DEPTH, int(round(n * DEPTH / 100.0)), acc, WINDOW,
                                                                  smo, "reference" if kind == "reference" else "oracle"),
            "stage_gbps": {"accumulate": n / acc / 1e9, "smooth": n / smo / 1e9}}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on all host cores (one process per core, each a
    bounded sample of the workload: the reference has no threads, chromosomes are its only
    parallelism)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 8_000_000
    ctx = mp.get_context("spawn")
    times = []
    kind = "port"
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_sample_worker, [(n, 100 + step * cores + i) for i in range(cores)])
            dt = time.perf_counter() - t0
            kind = res[0][2]
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = cores * n / (ms / 1e3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hg38 depth accumulation + smooth --window=101 (bounded sample: %d x %d-base "
                                   "chromosomes per step, depth %d)" % (cores, n, DEPTH)},
            "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": cores, "kind": kind,
                             "sample": "%d processes x one %d-base chromosome per step" % (cores, n)},
            "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from genodsp_b200 import slab
    from genodsp_b200.genome import Genome

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)

    chroms = scaled_genome(args.scale)
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    sorted_chroms = [chroms[i] for i in order]
    total_bases = sum(l for _, l in chroms)
    h = (WINDOW - 1) // 2

    if world == 1:
        g = Genome(chroms, device=local)
        plan = []
    else:
        segs_s, buffer_cells = slab.partition([l for _, l in sorted_chroms], world, rank, h)
        segs = [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s]
        g = Genome(chroms, device=local, segs=segs, buffer_cells=buffer_cells)
        plan = slab.halo_plan([l for _, l in sorted_chroms], world, rank, h)

    # intervals: generated for the whole genome with a per-chromosome seed (identical on every
    # rank), then each rank keeps those that overlap a piece it owns, indexed by layout segment
    cs, st, en = synth_intervals(torch, device, sorted_chroms)
    if world == 1:
        seg_t, start_t, end_t = cs, st, en
    else:
        keep_seg, keep_s, keep_e = [], [], []
        for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
            si = order.index(ci)
            p1 = pos0 + (hi - lo)
            msk = (cs == si) & (en.to(torch.int64) > pos0) & (st.to(torch.int64) < p1)
            keep_seg.append(torch.full((int(msk.sum()),), k, dtype=torch.int32, device=device))
            keep_s.append(st[msk]); keep_e.append(en[msk])
        seg_t, start_t, end_t = torch.cat(keep_seg), torch.cat(keep_s), torch.cat(keep_e)
        del cs, st, en
    n_intervals = int(seg_t.shape[0])
    my_bases = g.cells

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step_resident(timers=None):
        if timers is not None:
            timers[0].record()
        g.accumulate(seg_t, start_t, end_t, host=False)
        if timers is not None:
            timers[1].record()
        if plan:
            slab.exchange_halos(g.sig, plan, dist)
        if timers is not None:
            timers[2].record()
        g.smooth(WINDOW)
        if timers is not None:
            timers[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_ev = [[ev() for _ in range(4)] for _ in range(args.steps)]
    t_begin, t_end = ev(), ev()
    barrier()
    t_begin.record()
    launches0 = g.launches
    for k in range(args.steps):
        step_resident(stage_ev[k])
    timed_launches = g.launches - launches0
    t_end.record()
    barrier()
    total_ms = t_begin.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    acc_ms = sum(e[0].elapsed_time(e[1]) for e in stage_ev) / args.steps
    xch_ms = sum(e[1].elapsed_time(e[2]) for e in stage_ev) / args.steps
    smo_ms = sum(e[2].elapsed_time(e[3]) for e in stage_ev) / args.steps

    # ---- end to end through the C-ABI with HOST buffers: pinned interval arrays in, fp64 signal out
    e2e_steps = max(1, min(args.steps, 3))
    seg_h = seg_t.cpu().pin_memory(); start_h = start_t.cpu().pin_memory(); end_h = end_t.cpu().pin_memory()
    out_h = torch.empty(g.buffer_cells, dtype=torch.float64, pin_memory=True)

    def step_e2e():
        g.accumulate_pinned(seg_h, start_h, end_h)
        if plan:
            slab.exchange_halos(g.sig, plan, dist)
        g.smooth_to_host(WINDOW, out_h)        # FIR of piece k+1 overlaps the D2H of piece k

    step_e2e()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps

    vals = torch.tensor([total_ms, acc_ms, xch_ms, smo_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, acc_ms, xch_ms, smo_ms, e2e_ms = [float(x) for x in vals.tolist()]
    n_iv = torch.tensor([n_intervals], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(n_iv)
    n_iv_total = int(n_iv.item())

    if rank == 0:
        ms_per_step = total_ms / args.steps
        peak, peak_src = measured_hbm_peak()
        per_gpu_bases = total_bases / world
        smooth_gbs = 16.0 * per_gpu_bases / (smo_ms / 1e3) / 1e9
        fp64_rate = 2.0 * WINDOW * per_gpu_bases / (smo_ms / 1e3)
        acc_bytes = 16.0 * per_gpu_bases + 28.0 * n_iv_total / world
        line = {
            "metric": METRIC, "value": total_bases / (ms_per_step / 1e3) / 1e9, "unit": "Gbp/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "hg38-shaped 24 chromosomes, %d bases%s; %d intervals (reads U{50..150}, depth %d); "
                                   "step = depth accumulation (reads binned per 8192-cell tile, shared-memory difference + scan per tile) + smooth "
                                   "--window=%d" % (total_bases, "" if args.scale == 1 else " (lengths / %d)" % args.scale,
                                                    n_iv_total, DEPTH, WINDOW),
                       "l2": "inputs larger than L2 (%.1f GB signal per GPU)" % (8.0 * per_gpu_bases / 1e9),
                       "parallelism": "slab x%d, halo %d cells" % (world, h) if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": total_bases / (e2e_ms / 1e3) / 1e9, "unit": "Gbp/s",
                    "h2d_bytes_per_step": 12 * n_iv_total, "d2h_bytes_per_step": 8 * total_bases,
                    "ms_per_step": e2e_ms,
                    "what": "gdsp_accumulate_host(pinned seg/start/end) + gdsp_smooth per chromosome piece, each piece's fp64 result copied to pinned host memory while the next piece is computed"},
            "gpu_launches": timed_launches,
            "stages": {"accumulate": {"ms": acc_ms, "gbp_s": total_bases / (acc_ms / 1e3) / 1e9,
                                      "achieved_gbs": acc_bytes / (acc_ms / 1e3) / 1e9,
                                      "frac_hbm": acc_bytes / (acc_ms / 1e3) / 1e9 / peak},
                       "halo_exchange": {"ms": xch_ms},
                       "smooth": {"ms": smo_ms, "gbp_s": total_bases / (smo_ms / 1e3) / 1e9,
                                  "achieved_gbs": smooth_gbs, "frac_hbm": smooth_gbs / peak,
                                  "fp64_instr_per_base": 2 * WINDOW,
                                  "fp64_lane_ops_per_s": 2 * WINDOW * per_gpu_bases / (smo_ms / 1e3)}},
            "roofline": {"kernel": "k_smooth_ct", "bound": "hbm", "achieved": smooth_gbs, "peak": peak, "unit": "GB/s",
                         "frac": smooth_gbs / peak, "traffic": SMOOTH_TRAFFIC_BYTES_PER_BASE * per_gpu_bases if SMOOTH_TRAFFIC_BYTES_PER_BASE else None,
                         "peak_source": peak_src,
                         "binding_limit": {"kind": "fp64_issue", "achieved": fp64_rate, "unit": "FP64 instr/s",
                                           "peak_nominal": FP64_NOMINAL, "frac_nominal": fp64_rate / FP64_NOMINAL,
                                           "peak_measured": FP64_MEASURED, "frac_measured": fp64_rate / FP64_MEASURED,
                                           "peak_measured_source": "scripts/micro/fp64_peak.cu (DMUL+DADD, constant-bank operand)"},
                         "note": "16 B/bp algorithmic; the exact-order FIR needs 2*W=202 separately rounded FP64 "
                                 "instructions per base, so FP64 issue (64 lanes/SM), not HBM, is the binding limit for W=101"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="divide every chromosome length (1 = full hg38)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
