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
FP64_PER_BASE = 3 * ((WINDOW - 1) // 2) + 2   # k_smooth_sym: one DMUL + two DADD per symmetric tap pair, DMUL + DADD for the centre
SMOOTH_TRAFFIC_BYTES_PER_BASE = 16.27   # k_smooth_sym<50>: (25.92 GB read + 24.33 GB written) / 3,088,269,832 bases (strip ramps re-read 1.3 %)
SMOOTH_TRAFFIC_SOURCE = "profiles/r2_bench_launches.csv"
ACC_MOVED_BYTES_PER_BASE = 9.0          # 27.8 GB per accumulate launch set (k_bin_scatter_fixed 2.50 GB, k_bin_final 25.29 GB) / 3,088,269,832 bases
ACC_MOVED_SOURCE = "ncu constant, profiles/r2_bench_launches.csv (DRAM read + written by k_bin_scatter_fixed, k_tile_prefix, k_bin_final)"
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



# ----------------------------------------------------------------------------- NUMA
def bind_to_gpu_numa_node(torch, local):
    """Pin this process (and therefore the first-touch placement of the pinned host buffers of the
    end-to-end leg) to the CPUs of the NUMA node the GPU hangs off: at N > 1 eight ranks writing
    pinned memory of ONE node stop at ~93 GB/s aggregate.  Best effort."""
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        info.update({"pci": bdf, "node": node})
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return info
        os.sched_setaffinity(0, use)
        info.update({"bound": True, "cpus": len(use)})
    except Exception as e:
        info["error"] = repr(e)[:120]
    return info

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


def cli_cfg1():
    """BASELINE configs[0] through the drop-in CLI, text in and text out (genodsp_b200/bin/genodsp), beside the
    reference binary when it is built: wall clock of the whole process, CUDA initialisation included."""
    import hashlib
    import tempfile
    import numpy as np
    ours = os.path.join(ROOT, "genodsp_b200", "bin", "genodsp")
    ref = os.path.join(ROOT, "oracle", "_ref", "genodsp")
    if not os.path.exists(ours):
        return None
    rng = np.random.default_rng(1)
    n, m = 10_000_000, 1_000_000
    cmd = ["--chromosomes=g.chroms", "--novalue", "=", "sum", "--window=101", "=", "localmax", "--neighborhood=11"]
    out = {"workload": "one 10 Mbp chromosome, 1 M intervals, --novalue = sum --window=101 = localmax --neighborhood=11 (text to text)"}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        with open(os.path.join(d, "g.chroms"), "w") as f:
            f.write("chr1 %d\n" % n)
        st = np.sort(rng.integers(0, n - 150, m)); en = st + rng.integers(50, 151, m)
        with open(os.path.join(d, "reads.iv"), "w") as f:
            f.write("".join("chr1\t%d\t%d\n" % (a, b) for a, b in zip(st.tolist(), en.tolist())))
        digests = {}
        for name, binary in (("ours", ours), ("reference", ref)):
            if not os.path.exists(binary):
                continue
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                with open(os.path.join(d, "reads.iv"), "rb") as fin, open(os.path.join(d, "out.txt"), "wb") as fout:
                    p = subprocess.run([binary] + cmd, stdin=fin, stdout=fout, stderr=subprocess.PIPE, cwd=d,
                                       env=dict(os.environ, GENODSP_TIMING="1"))
                dt = time.perf_counter() - t0
                if p.returncode != 0:
                    return {"error": p.stderr.decode(errors="replace")[-300:]}
                best = dt if best is None else min(best, dt)
            digests[name] = hashlib.md5(open(os.path.join(d, "out.txt"), "rb").read()).hexdigest()
            out[name + "_s"] = round(best, 3)
            if name == "ours":
                out["ours_phases_s"] = {" ".join(l.split()[1:-2]): float(l.split()[-2])
                                        for l in p.stderr.decode(errors="replace").splitlines() if l.startswith("[timing]")}
        if len(digests) == 2:
            out["identical_output"] = digests["ours"] == digests["reference"]
    out["note"] = ("the CUDA driver's initialisation (cuInit + context: 1.0-2.3 s on this pool, scripts/micro/cuda_init_time.cu) is most of our "
                   "wall clock on a 10 Mbp toy; hg38/16 text-to-text is in profiles/r2_cli_bench.log")
    return out


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
ALIGN_CUTS = 8192    # slab cuts inside a chromosome are multiples of this in chromosome coordinates (every kernel's tile divides it)
HALO = 4096          # readable cells either side of a slab cut: open/close 1001 reach 1003, clump's carry tile is 4096


def second_track(g, np):
    """cfg5's second signal: sorted disjoint intervals covering ~50 % of every owned piece, values k/1024"""
    rng = np.random.default_rng(99)
    bs, bstart, bend, bval = [], [], [], []
    for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(g.segs):
        n = hi - lo
        m = max(1, n // 2000)
        cuts = np.unique(rng.integers(0, n, 2 * m)) + pos0
        a, b = cuts[0::2], cuts[1::2]
        m2 = min(a.size, b.size)
        bs.append(np.full(m2, k, np.uint32)); bstart.append(a[:m2].astype(np.uint32)); bend.append(b[:m2].astype(np.uint32))
        bval.append(rng.integers(1, 2048, m2) / 1024.0)
    return g.interval_table(np.concatenate(bs), np.concatenate(bstart), np.concatenate(bend), np.concatenate(bval))


class Pipelines:
    """The BASELINE.json pipelines on one rank's Genome (whole genome at world 1, a slab otherwise).
    Every method is one operator; slab-aware where the operator needs more than its own cells."""

    def __init__(self, torch, dist, g, plan, world, rank, intervals, lengths):
        from genodsp_b200 import capi, slab
        import numpy as np
        self.t, self.dist, self.g, self.plan, self.world, self.rank = torch, dist, g, plan, world, rank
        self.slab, self.capi, self.np = slab, capi, np
        self.seg_t, self.start_t, self.end_t = intervals
        # N > 1: the library's own NCCL binding (gdsp_comm_*: halos by ncclSend/ncclRecv, counts by ncclAllReduce);
        # GDSP_BENCH_TORCH_COMM=1 falls back to torch.distributed collectives for comparison
        if world > 1:
            self.comm = slab.DistComm(dist, g.device) if os.environ.get("GDSP_BENCH_TORCH_COMM") else slab.GdspComm(g, dist)
        else:
            self.comm = slab.VirtualComm([g])
        self.tableB = second_track(g, np)
        cap = max(1024, g.cells // 4)
        self.rbufs = (torch.empty(cap, dtype=torch.int32, device=g.device), torch.empty(cap, dtype=torch.int32, device=g.device),
                      torch.empty(cap, dtype=torch.float64, device=g.device))
        self.nruns = 0
        self.vars = {}
        self.known = None
        self.lengths = lengths
        self.overlap = slab.OverlappedExchange(g, plan, dist, (WINDOW - 1) // 2,
                                               use_gdsp_comm=not os.environ.get("GDSP_BENCH_TORCH_COMM")) if world > 1 else None
        self.taps = None

    def close(self):
        self.tableB.close()
        if self.overlap is not None:
            self.overlap.close()
        if hasattr(self.comm, "close"):
            self.comm.close()

    def xch(self, radius):
        if self.world > 1:
            if hasattr(self.comm, "exchange"):
                self.comm.exchange(self.g.sig, self.plan, radius)            # ncclSend/ncclRecv from C, stream-ordered
            else:
                self.slab.exchange_halos(self.g.sig, self.plan, self.dist, radius)

    # ---- operators
    def depth(self):
        self.g.accumulate(self.seg_t, self.start_t, self.end_t, host=False)

    def smooth(self):
        g = self.g
        if self.overlap is None:
            g.smooth(WINDOW)
            return
        import ctypes as C
        from genodsp_b200.genome import hann_taps
        if self.taps is None:
            self.taps = hann_taps(WINDOW)
        tp = self.taps.ctypes.data_as(C.POINTER(C.c_double))
        self.overlap.run(lambda lay: self.capi.check(g.lib.gdsp_smooth(g.ctx, lay, g._p(g.sig), g._p(g.tmp), WINDOW, tp)))
        g._swap()

    def localmax(self):
        self.xch(5); self.g.localmax(11)

    def sum100(self):
        self.xch(100); self.g.sum(100, denom=100.0)

    def percentile99(self):
        if self.world == 1:
            self.vars.update(self.g.percentile(99.0, destructive=True))
        else:
            (v,), n, ((below, equal),), nan = self.slab.slab_percentiles([self.g], self.comm, [99000], ranked=True)
            self.vars["percentile99"] = v
            self.known = (below, equal, nan, n, sum(self.lengths))

    def binarize_after_percentile(self):
        if self.world == 1:
            self.g.binarize(self.vars["percentile99"])           # consumes the pending sorted state: count + fill
        else:
            self.slab.slab_sorted_binarize([self.g], self.comm, self.vars["percentile99"], known=self.known)

    def binarize6(self):
        self.g.binarize(6.0)

    def open1001(self):
        self.xch(1003); self.g.open_(1001, 0.5)

    def close1001(self):
        self.xch(1003); self.g.close_(1001, 0.5)

    def clump(self):
        if self.world == 1:
            self.g.clump(0.5, 1000)
        else:
            self.xch(4096)
            self.slab.slab_clump_carries([self.g], self.comm, average=0.5, length=1000)

    def runs(self):
        self.nruns = self.g.runs_device(self.rbufs)[0]

    def chain5(self):
        c, tb = self.capi, self.tableB
        self.g.pointwise([(c.PW_IVL_ADD, 0.0, 0, 0, 0, tb), (c.PW_IVL_MUL, 0.0, 0, 0, 0, tb), (c.PW_IVL_SET, 0.0, 0, 0, 0, tb),
                          (c.PW_NONZERO_TO_ONE, 0.0), (c.PW_IVL_SET_OUTSIDE, 0.0, 0, 0, 0, tb), type(self.g).op_binarize(0.5)])

    def pipelines(self):
        """name -> [(operator name, callable, algorithmic bytes per base)]"""
        return {
            "pipe5": [("depth", self.depth, 16), ("smooth101", self.smooth, 16), ("localmax11", self.localmax, 16),
                      ("percentile99", self.percentile99, 8), ("binarize", self.binarize_after_percentile, 16), ("runs", self.runs, 8)],
            "cfg3": [("depth", self.depth, 16), ("sum100", self.sum100, 16), ("percentile99", self.percentile99, 8),
                     ("binarize", self.binarize_after_percentile, 16)],
            "cfg4": [("depth", self.depth, 16), ("binarize6", self.binarize6, 16), ("open1001", self.open1001, 16),
                     ("close1001", self.close1001, 16), ("clump", self.clump, 16), ("runs", self.runs, 8)],
            "cfg5": [("depth", self.depth, 16), ("chain6_one_launch", self.chain5, 16)],
        }


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
    lengths = [l for _, l in sorted_chroms]
    total_bases = sum(l for _, l in chroms)
    h = (WINDOW - 1) // 2

    if world == 1:
        g = Genome(chroms, device=local)
        plan = []
    else:
        segs_s, buffer_cells = slab.partition(lengths, world, rank, HALO, ALIGN_CUTS)
        segs = [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s]
        g = Genome(chroms, device=local, segs=segs, buffer_cells=buffer_cells)
        plan = slab.halo_plan(lengths, world, rank, HALO, ALIGN_CUTS)

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
    P = Pipelines(torch, dist, g, plan, world, rank, (seg_t, start_t, end_t), lengths)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    # --overlap: the headline step as a two-stream pipeline: accumulation of chromosome group k+1 (side stream) behind
    # the FIR of group k (compute stream).  Default: the two operators back to back on one stream.
    dsp = slab.DepthSmoothPipeline(g, plan, dist, WINDOW, seg_t, start_t, end_t, args.groups) if args.overlap else None

    def step_resident(timers=None):
        if dsp is not None:
            dsp.run(timed=timers is not None)
            if timers is not None:
                timers.append(dsp.stage_events)
            return
        if timers is not None:
            timers[0].record()
        P.depth()
        if timers is not None:
            timers[1].record()
        P.smooth()                       # at N > 1 the halo exchange runs on a side stream behind the interior FIR
        if timers is not None:
            timers[2].record()

    # the pipelined step must leave the bits of accumulate() followed by smooth(): checked on this rank's cells
    overlap_equal = None
    if dsp is not None:
        P.depth(); P.smooth()
        want = g.sig.clone()
        dsp.run()
        torch.cuda.synchronize()
        overlap_equal = all(bool(torch.equal(g.sig[lo:hi].view(torch.int64), want[lo:hi].view(torch.int64))) for (lo, hi, *_r) in g.segs)
        del want
        if world > 1:
            flag = torch.tensor([1 if overlap_equal else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            overlap_equal = bool(flag.item())

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
    stage_ev = [[ev() for _ in range(3)] for _ in range(args.steps)]
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
    if dsp is not None:
        # per-launch durations summed per stage (the stages overlap in time: their sum exceeds the step)
        acc_ms = sum(sum(a.elapsed_time(b) for a, b in e[3][0]) for e in stage_ev) / args.steps
        smo_ms = sum(sum(a.elapsed_time(b) for a, b in e[3][1]) for e in stage_ev) / args.steps
    else:
        acc_ms = sum(e[0].elapsed_time(e[1]) for e in stage_ev) / args.steps
        smo_ms = sum(e[1].elapsed_time(e[2]) for e in stage_ev) / args.steps

    # ---- end to end through the C-ABI with HOST buffers: pinned interval arrays in, fp64 signal out
    e2e_steps = max(1, min(args.steps, 3))
    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    seg_h = seg_t.cpu().pin_memory(); start_h = start_t.cpu().pin_memory(); end_h = end_t.cpu().pin_memory()
    out_h = torch.empty(g.buffer_cells, dtype=torch.float64, pin_memory=True)

    def step_e2e():
        g.accumulate_pinned(seg_h, start_h, end_h)
        if plan:
            P.xch(h)
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
    del out_h, seg_h, start_h, end_h
    os.sched_setaffinity(0, affinity0)

    # ---- every pipeline of BASELINE.json under the same clock: per-operator CUDA-event
    # times, max over ranks, best of `--stage-reps` passes after one warm-up pass
    stage_out = {}
    finals = {}
    if not args.no_stages:
        for name, ops in P.pipelines().items():
            best, best_ops = None, None
            for rep in range(args.stage_reps + 1):
                barrier()
                marks = [ev() for _ in range(len(ops) + 1)]
                marks[0].record()
                for k, (_, fn, _) in enumerate(ops):
                    fn(); marks[k + 1].record()
                barrier()
                ms = torch.tensor([marks[0].elapsed_time(marks[-1])] + [marks[k].elapsed_time(marks[k + 1]) for k in range(len(ops))],
                                  dtype=torch.float64, device=device)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ms = ms.tolist()
                if rep > 0 and (best is None or ms[0] < best):
                    best, best_ops = ms[0], ms[1:]
            stage_out[name] = (best, best_ops, [o[0] for o in ops], [o[2] for o in ops])
            finals[name] = dict(P.vars)
            # ---- parity of the slab run: rank 0 reruns the pipeline on the WHOLE genome on its own GPU (the N = 1
            # path) and every rank's pieces are compared with it bit for bit
            if world > 1 and not args.no_parity:
                if rank == 0:
                    if "whole" not in finals:
                        gw = Genome(chroms, device=local)
                        wi = synth_intervals(torch, device, sorted_chroms)
                        finals["whole"] = Pipelines(torch, dist, gw, [], 1, 0, wi, lengths)
                    W = finals["whole"]
                    for _, fn, _ in W.pipelines()[name]:
                        fn()
                    _ = W.g.sig                                   # materialise a pending sorted state, if any
                    whole_g = W.g
                    same_vars = all(W.vars.get(k) == v for k, v in P.vars.items())
                else:
                    whole_g, same_vars = None, True
                compared, differ, cut = slab.compare_with_whole(g, whole_g, dist, rank, world)
                finals.setdefault("parity", {})[name] = {"cells_compared": compared, "cells_differ": differ,
                                                         "cut_chromosome_pieces": cut, "variables_equal": bool(same_vars)}
        if "whole" in finals:
            finals["whole"].close(); finals["whole"].g.close(); del finals["whole"]

    vals = torch.tensor([total_ms, acc_ms, smo_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, acc_ms, smo_ms, e2e_ms = [float(x) for x in vals.tolist()]
    n_iv = torch.tensor([n_intervals], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(n_iv)
    n_iv_total = int(n_iv.item())

    ok = True
    if rank == 0:
        ms_per_step = total_ms / args.steps
        peak, peak_src = measured_hbm_peak()
        per_gpu_bases = total_bases / world
        smooth_gbs = 16.0 * per_gpu_bases / (smo_ms / 1e3) / 1e9
        fp64_rate = FP64_PER_BASE * per_gpu_bases / (smo_ms / 1e3)
        acc_bytes = 16.0 * per_gpu_bases + 28.0 * n_iv_total / world
        acc_moved = ACC_MOVED_BYTES_PER_BASE * per_gpu_bases
        pipes = {}
        for name, (best, best_ops, op_names, op_bytes) in stage_out.items():
            ops = {}
            for nm, ms, bpb in zip(op_names, best_ops, op_bytes):
                gbs = bpb * per_gpu_bases / (ms / 1e3) / 1e9
                ops[nm] = {"ms": round(ms, 3), "gbp_s": round(total_bases / (ms / 1e3) / 1e9, 2), "alg_bytes_per_bp": bpb,
                           "achieved_gbs_per_gpu": round(gbs, 1), "frac_hbm": round(gbs / peak, 3)}
            pipes[name] = {"ms": round(best, 3), "gbp_s": round(total_bases / (best / 1e3) / 1e9, 2), "ops": ops}
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
                       "parallelism": ("slab x%d, halo %d cells exchanged by ncclSend/ncclRecv (gdsp_comm_exchange_halos) on a side stream behind the interior FIR"
                                       % (world, h)) if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": total_bases / (e2e_ms / 1e3) / 1e9, "unit": "Gbp/s",
                    "h2d_bytes_per_step": 12 * n_iv_total, "d2h_bytes_per_step": 8 * total_bases,
                    "ms_per_step": e2e_ms, "numa": numa,
                    "what": "gdsp_accumulate_host(pinned seg/start/end) + gdsp_smooth per chromosome piece, each piece's fp64 result copied to pinned host memory while the next piece is computed"},
            "gpu_launches": timed_launches,
            "stages": {"accumulate": {"ms": acc_ms, "gbp_s": total_bases / (acc_ms / 1e3) / 1e9,
                                      "achieved_gbs": acc_bytes / (acc_ms / 1e3) / 1e9,
                                      "frac_hbm": acc_bytes / (acc_ms / 1e3) / 1e9 / peak,
                                      "moved_bytes_per_base": ACC_MOVED_BYTES_PER_BASE, "moved_source": ACC_MOVED_SOURCE,
                                      "frac_hbm_on_moved_bytes": acc_moved / (acc_ms / 1e3) / 1e9 / peak},
                       "smooth": {"ms": smo_ms, "gbp_s": total_bases / (smo_ms / 1e3) / 1e9,
                                  "achieved_gbs": smooth_gbs, "frac_hbm": smooth_gbs / peak,
                                  "fp64_instr_per_base": FP64_PER_BASE,
                                  "fp64_lane_ops_per_s": FP64_PER_BASE * per_gpu_bases / (smo_ms / 1e3),
                                  "includes": "halo exchange (side stream) + interior + edge launches" if world > 1 else "one launch"},
                       "pipelines": pipes},
            "roofline": {"kernel": "k_smooth_sym<50>", "bound": "fp64_issue", "achieved": smooth_gbs, "peak": peak, "unit": "GB/s",
                         "frac": smooth_gbs / peak,
                         "traffic": SMOOTH_TRAFFIC_BYTES_PER_BASE * per_gpu_bases if SMOOTH_TRAFFIC_BYTES_PER_BASE else None,
                         "traffic_source": "ncu constant (dram__bytes_read.sum + dram__bytes_write.sum per base, %s), not measured in this run" % SMOOTH_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "secondary": {"bound": "hbm", "frac": smooth_gbs / peak},
                         "binding_limit": {"kind": "fp64_issue", "achieved": fp64_rate, "unit": "FP64 instr/s",
                                           "peak_nominal": FP64_NOMINAL, "frac_nominal": fp64_rate / FP64_NOMINAL,
                                           "peak_measured": FP64_MEASURED, "frac_measured": fp64_rate / FP64_MEASURED,
                                           "peak_measured_source": "scripts/micro/fp64_peak.cu (DMUL+DADD, constant-bank operand)"},
                         "note": "`achieved`/`frac` are the contract's HBM figures (16 B/bp algorithmic); the exact-order FIR needs "
                                 "separately rounded FP64 multiplies and adds -- 2*W = 202 per base done directly, 3*(W-1)/2+2 = 152 with the "
                                 "product of each symmetric tap pair computed once (k_smooth_sym, round 2) -- so FP64 issue (64 lanes/SM), "
                                 "not HBM, is the binding limit for W=101: see binding_limit"},
        }
        line["config"]["step"] = ("two-stream pipeline over %d chromosome groups: accumulate(group k+1) on a high-priority side stream "
                                  "behind smooth(group k); stage times are sums of per-launch durations and overlap" % args.groups) \
            if dsp is not None else "accumulate, then smooth, on one stream"
        if overlap_equal is not None:
            line["pipelined_step_bit_equal_to_sequential"] = bool(overlap_equal)
            ok = ok and bool(overlap_equal)
        if "parity" in finals:
            par = finals["parity"]
            ok = ok and all(v["cells_differ"] == 0 and v["variables_equal"] for v in par.values())
            line["parity_check"] = {"against": "the same pipelines run on the whole genome on rank 0's GPU (the N = 1 path)",
                                    "cut_chromosomes": max(v["cut_chromosome_pieces"] for v in par.values()) - 0,
                                    "bit_equal": ok, "pipelines": par}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single()
            try:
                cli = cli_cfg1()
                if cli is not None:
                    line["stages"]["cli"] = {"cfg1": cli}
            except Exception as e:                            # the CLI leg never takes the bench line down
                line["stages"]["cli"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if dsp is not None:
        dsp.close()
    P.close()
    g.close()
    if world > 1:
        okt = torch.tensor([1 if ok else 0], device=device)
        dist.broadcast(okt, 0)
        ok = bool(okt.item())
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="divide every chromosome length (1 = full hg38)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-pipeline stage timings (pipe5, cfg3, cfg4, cfg5)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bit-comparison of the slab results with the whole-genome run on rank 0")
    ap.add_argument("--stage-reps", type=int, default=2)
    ap.add_argument("--overlap", action="store_true",
                    help="headline step as a two-stream pipeline (accumulate of chromosome group k+1 behind smooth of group k). Off by "
                         "default: measured on B200 it buys nothing -- the FIR stretches by what the accumulation takes, 45.49 vs 45.35 ms "
                         "(profiles/r2_overlap_experiment.md)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node while the pinned buffers of the e2e leg are allocated and used")
    ap.add_argument("--groups", type=int, default=6, help="chromosome groups of the two-stream pipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
