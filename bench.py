#!/usr/bin/env python3
"""bench.py — motif x bp scored per second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (libtfbs_cuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # host-CPU reference arm

Headline workload (config C4 of BASELINE.json, weak scaling): every GPU scans its own 100 Mb chunk
(+ Lmax-1 halo) of one N x 100 Mb position-addressable synthetic genome against an 800-PWM
JASPAR-sized library (lengths 6..30), both strands, thresholded hits (expected rate 1e-4 per
position per strand).  One unit = one motif evaluated at one start position on both strands.

A step = encode (ASCII -> 2-bit + mask) + hits-mode scan (filter scan + exact verify pass) of the chunk.
  value : inputs (ASCII bytes) already resident in HBM, hits left in HBM; CUDA events on the
          library's stream; L2 flushed (untimed) before every step.
  e2e   : the same step through the public API with HOST buffers (tfbs_scan_hits_host_shard): pinned
          ASCII -> H2D -> encode -> scan -> hits + count D2H into pinned memory, chunk-pipelined on
          three streams, halo windows dropped and positions made global on the device, all inside the
          timed region.
Further records on the same JSON line: `scaling_strong` (the SAME 100 Mb genome cut
into N shards), `c5` (the per-GPU shard of the 3.1 Gb x 800 + 4 shape tracks config), `product_api`
(one process driving all N GPUs through sharding.MultiGpuScanner, wall clock), `e2e_wall` (host
wall clock of the public call: warm, cold threshold vector, sorted), `cpu_baseline_legs`.
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

METRIC = "motif_x_bp_scored_per_sec"
UNIT = "motif*bp/s"
SEED = 39
C4_TOTAL = 100_000_000          # config C4 genome
C5_PER_GPU = 387_500_000        # config C5: 3.1 Gb / 8 GPUs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=100.0, help="Mb scanned per GPU (headline, weak scaling)")
    ap.add_argument("--motifs", type=int, default=800)
    ap.add_argument("--rate", type=float, default=1e-4, help="expected hit rate per position per strand")
    ap.add_argument("--chunks", type=int, default=6, help="chunks of the pipelined end-to-end pass")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip every additional record")
    ap.add_argument("--c5", choices=["auto", "on", "off"], default="auto",
                    help="config C5 record (387.5 Mb per GPU + 4 shape tracks); auto = when --gpus is 8")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU budget of the cpu_baseline sample")
    return ap.parse_args()


try:
    _ALL_CPUS = os.sched_getaffinity(0)     # before any NUMA binding of the GPU arm
except AttributeError:
    _ALL_CPUS = None


def host_threads() -> int:
    """Cores this process may run on.  NOT the OpenMP default: torch.distributed.run exports
    OMP_NUM_THREADS=1, which made the round-1 CPU arm single-threaded at N >= 2."""
    if _ALL_CPUS is not None:
        return max(len(_ALL_CPUS), 1)
    return max(os.cpu_count() or 1, 1)


def workload(args):
    from discover_tfbs_b200 import synth

    logodds, lens = synth.pwm_library(SEED, args.motifs, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=args.rate, gc=0.41, seed=SEED, sample=100_000)
    return logodds, lens, thr


def config_dict(args, n_gpus, per_gpu):
    return {
        "workload": "C4: 800-PWM JASPAR-sized library (L 6-30) x 100 Mb synthetic genome per GPU, both strands, "
                    "thresholded hits" if (args.motifs == 800 and per_gpu == C4_TOTAL) else
                    f"{args.motifs} PWMs (L 6-30) x {per_gpu} bp per GPU, both strands, thresholded hits",
        "motifs": args.motifs, "motif_len": "6-30", "genome_bp_per_gpu": per_gpu,
        "genome_bp_total": per_gpu * n_gpus, "hit_rate_target": args.rate, "mode": "hits",
        "sharding": f"chunk+halo x{n_gpus}, no collective", "l2": "flushed before every timed step",
        "seed": SEED,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None  # lines: (arrival time, text)
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def mark_end(self):
        self.t_end = time.monotonic()

    def stop(self):
        """Summary of the samples that arrived inside [mark_begin, mark_end] (the timed regions);
        nvidia-smi needs ~1 s to start on an 8-GPU box, so it is launched before the warm-up and,
        if the timed window is shorter than the sampling period, the warm-up samples are used."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [ln for t, ln in self.lines if self.t_begin is not None and self.t_begin <= t <= (self.t_end or 1e30) + 0.1]
        window = "timed region"
        if len(inside) < 2:
            inside, window = [ln for _, ln in self.lines], "warm-up + timed region"
        sm, smax, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(args, per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at the headline
    size, from the committed `ncu --set full` capture of this round (profiles/r02/ncu_hits_traffic.json,
    written by profiles/ncu_extract.py); None when the run is not that configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02", "ncu_hits_traffic.json")) as fh:
            t = json.load(fh)
        if (t["motifs"] == args.motifs and t["genome_bp"] == per_gpu and abs(t["rate"] - args.rate) < 1e-12):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


# ---- CPU arms ---------------------------------------------------------------------------------------
def cpu_pwm_rate(logodds, lens, thr, seconds, threads, min_bp=10_000_000, gc=0.41):
    """Time the oracle port of Biopython's scan on a bounded sample of the same workload: at least
    10 Mb x the whole library, more if the budget allows."""
    from discover_tfbs_b200 import synth
    from oracle import pssm_oracle

    probe = synth.genome(SEED, 0, 65_536, gc=gc)
    t0 = time.perf_counter()
    pssm_oracle.scan_count(probe, logodds, lens, thr, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    n = int(min(max(65_536 * seconds / dt, min_bp), 50e6))
    sample = synth.genome(SEED, 0, n, gc=gc)
    t0 = time.perf_counter()
    hits, _, used = pssm_oracle.scan_count(sample, logodds, lens, thr, threads=threads)
    dt = time.perf_counter() - t0
    units = int(sum(max(n - int(L) + 1, 0) for L in lens))
    return {"value": units / dt, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{n} bp x {len(lens)} PWMs both strands (oracle C port of Biopython _pwm.c semantics, "
                      f"OpenMP, {used} threads from sched_getaffinity), {dt:.1f} s, {hits} hits; linear in bp"}


def cpu_legs(threads):
    """CPU figures for the shape and feature-table legs: the DNAshapeR
    restatement timed here single-threaded and with OpenMP, and the REAL reference (utils.py via
    oracle/refshim.py) as timed in the build container by profiles/reference_cpu_legs.py (the reference
    cannot travel to the GPU box; its committed JSON is quoted with its provenance)."""
    from discover_tfbs_b200 import dnashape, synth

    out = {}
    n = 20_000_000
    seq = synth.genome(SEED, 0, n, gc=0.41)
    tables, kinds = dnashape.synthetic_tables(("MGW", "ProT", "Roll", "HelT"))
    starts = np.arange(0, n, 1 << 16, dtype=np.int64)
    lens = np.minimum(starts + (1 << 16), n) - starts
    from oracle import clib

    tab = np.ascontiguousarray(tables, dtype=np.float32)
    kd = np.ascontiguousarray(kinds, dtype=np.int32)
    for label, th in (("single_thread", 1), ("openmp", threads)):
        m = (n if th > 1 else n // 8) >> 16 << 16      # whole 64 Ki-base records only
        k = m >> 16
        buf = np.ascontiguousarray(seq[:m])
        res = np.empty((4, m), dtype=np.float32)
        call = lambda: clib.lib().oracle_shape_tracks_set(clib.ptr(buf), clib.ptr(starts), clib.ptr(lens), k, m,  # noqa: E731
                                                          clib.ptr(tab), clib.ptr(kd), 4, th, clib.ptr(res))
        call()                                         # untimed: first touch of the output pages
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        out[f"shape_tracks_{label}"] = {"value": m / dt, "unit": "bp/s (4 tracks)", "cores": th, "kind": "port",
                                        "sample": f"{m} bp in 64 Ki-base records, MGW ProT Roll HelT, DNAshapeR "
                                                  f"restatement (oracle C), {dt:.2f} s"}
    try:
        with open(os.path.join(ROOT, "profiles", "r02", "reference_cpu_legs.json")) as fh:
            out["reference_python"] = json.load(fh)
    except Exception as exc:
        out["reference_python"] = {"error": repr(exc)}
    return out


def run_reference(args):
    """Host-CPU reference arm.  Biopython (the PWM scorer the north star names) and DNAshapeR are
    not installable here and the reference has no PWM code of its own, so the arm times the oracle
    port of their algorithm with all host threads on a bounded sample (>= 10 Mb x 800) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from discover_tfbs_b200 import synth
    from oracle import pssm_oracle

    threads = host_threads()
    logodds, lens, thr = workload(args)
    per_gpu = int(args.genome_mb * 1e6)
    n = 10_000_000
    sample = synth.genome(SEED, 0, n)
    units = int(sum(max(n - int(L) + 1, 0) for L in lens))
    for _ in range(args.warmup):
        pssm_oracle.scan_count(sample, logodds, lens, thr, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits, _, used = pssm_oracle.scan_count(sample, logodds, lens, thr, threads=threads)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = units / dt
    desc = f"{n} bp x {len(lens)} PWMs both strands per step, {used} threads (sched_getaffinity), {hits} hits"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 accumulate -> f32", "data": "synthetic",
        "config": config_dict(args, args.gpus, per_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = oracle C port (Biopython PSSM semantics); the reference repo has no "
                "PWM scanner of its own and Biopython/DNAshapeR are absent from the image; throughput is "
                "per host (one CPU arm whatever --gpus says)",
    }
    emit(line)


# ---- GPU arm ----------------------------------------------------------------------------------------
class HitsJob:
    """One rank's shard of a genome scan: device-resident ASCII, encoded handle, pinned host buffers."""

    def __init__(self, ctx, lib, thr, lens, shard, total, rate, chunks, cap_factor=4.0, with_shape=False,
                 shape_sites=0):
        from discover_tfbs_b200 import _lib, dnashape, sharding

        self.ctx, self.lib, self.thr, self.lens, self.shard, self.chunks = ctx, lib, thr, lens, shard, chunks
        self.n = shard.scanned
        self.units = sharding.units_of_shard(shard, lens, total)
        self.dev_ascii = ctx.synth_ascii(SEED, shard.start, self.n, gc=0.41)   # position-addressable shard
        self.enc = ctx.encode_device(self.dev_ascii.ptr, self.n)
        self.cap = int(max(cap_factor * 2 * rate * self.units, 1 << 20))
        self.n_hits = ctx.scan_hits(self.enc, lib, thr, cap=self.cap, to_host=False)     # also sizes the scratch
        self.pinned_in = _lib.PinnedBuffer(self.n)
        self.pinned_in.array[:] = self.dev_ascii.download()
        self.pinned_out = _lib.PinnedBuffer(self.cap * 16)
        self.hits_out = self.pinned_out.view(_lib.HIT_DTYPE)
        self.enc_ms, self.scan_ms, self.verify_ms, self.shape_ms, self.direct_ms = [], [], [], [], []
        self.shape_tab, self.shape_sites, self.sites_bytes = None, int(shape_sites), 0
        if with_shape:
            st, sk = dnashape.synthetic_tables(("MGW", "ProT", "Roll", "HelT"))
            self.shape_tab = ctx.upload_shapes(st, sk)
            ctx.shape_tracks(self.enc, self.shape_tab, to_host=False)
        self.got = None

    def step_resident(self):
        self.enc.reencode_device(self.dev_ascii.ptr)
        self.enc_ms.append(self.ctx.last_kernel_ms())
        self.n_hits = self.ctx.scan_hits(self.enc, self.lib, self.thr, cap=self.cap, to_host=False)
        s, v = self.ctx.last_hits_ms()
        d = self.ctx.last_direct_ms()
        self.scan_ms.append(s - d)      # the table filter alone
        self.direct_ms.append(d)
        self.verify_ms.append(v)
        if self.shape_tab is not None:
            self.ctx.shape_tracks(self.enc, self.shape_tab, to_host=False)
            self.shape_ms.append(self.ctx.last_kernel_ms())

    def step_e2e(self, sort=False):
        # public end-to-end call: pinned host ASCII -> (H2D | encode + scan + verify | hits D2H) pipelined over
        # `chunks` chunks on three streams -> hits (global positions, halo dropped) in pinned host memory
        sh = self.shard
        self.got = self.ctx.scan_hits_host_shard(self.pinned_in.array, self.enc, self.lib, self.thr,
                                                 self.hits_out[: self.cap], owned=sh.owned, shard_start=sh.start,
                                                 chunks=self.chunks, sort=sort)
        if self.shape_tab is not None:
            # config C5: the 4 tracks of the shard stay in HBM (6.2 GB would not cross PCIe); what goes home is
            # the shape feature rows of the first `shape_sites` returned hits (reference slicing, L + 1 values)
            self.ctx.shape_tracks(self.enc, self.shape_tab, to_host=False)
            k = min(self.shape_sites, int(self.got.size))
            if k:
                h = self.got[:k]
                g = np.where(h["pos"] < 0, ~h["pos"], h["pos"]) - sh.start
                L = self.lens[h["motif"]].astype(np.int64)
                strand = np.where(h["pos"] < 0, -1, 1).astype(np.int8)
                rows = self.ctx.shape_sites(self.enc, self.shape_tab, np.zeros(k, np.int64), np.full(k, self.n, np.int64),
                                            g, g + L, strand, width=31, mode=0, trailing_empty=False)
                self.sites_bytes = int(rows.nbytes)
        return self.got

    def close(self):
        for h in (self.shape_tab, self.enc, self.dev_ascii, self.pinned_in, self.pinned_out):
            if h is not None:
                h.close()


def timed(ctx, fn, steps, warmup, barrier):
    """W untimed + K timed steps; L2 flushed (untimed) before each; device time from CUDA events on the
    library stream (copies of the pipelined path are ordered into it), host wall clock beside it."""
    for _ in range(warmup):
        ctx.flush_l2()
        fn()
    barrier()
    ms, wall = [], []
    launches0 = ctx.total_kernel_launches()   # the L2 flush is not counted by the library
    for _ in range(steps):
        ctx.flush_l2()
        t0 = time.perf_counter()
        ctx.timer_start()
        fn()
        ms.append(ctx.timer_stop())
        wall.append((time.perf_counter() - t0) * 1e3)
    launches = ctx.total_kernel_launches() - launches0
    barrier()
    return float(np.sum(ms)) / 1e3, float(np.mean(wall)), launches


def extras(ctx, args, peak):
    """Per-kernel HBM rooflines for the streaming kernels (encode, dense scan on config C2, shape tracks)
    and the per-record best-window kernel, each timed alone with CUDA events after warm-up, L2 flushed."""
    from discover_tfbs_b200 import dnashape, synth

    out = []
    # streaming kernels are timed on the per-GPU shard of config C5 (3.1 Gb / 8 GPUs = 387.5 Mb)
    n = int(C5_PER_GPU if args.genome_mb >= 100.0 else args.genome_mb * 1e6)
    dev = ctx.synth_ascii(SEED, 0, n, gc=0.41)
    enc = ctx.encode_device(dev.ptr, n)
    ms = []
    for i in range(6):
        ctx.flush_l2()
        enc.reencode_device(dev.ptr)
        ms.append(ctx.last_kernel_ms())
    t = float(np.mean(ms[3:]))

    def entry(kernel, workload, nbytes, t, **extra):
        return {"kernel": kernel, "workload": workload, "bound": "hbm", "achieved": nbytes / t / 1e6, "peak": peak,
                "unit": "GB/s", "frac": nbytes / t / 1e6 / peak, "ms": t, "algorithmic_bytes": nbytes, **extra}

    out.append(entry("encode_kernel", f"{n} bp ASCII -> 2-bit + mask", 1.375 * n, t, bytes_per_bp=1.375))
    # shape tracks, T = 4 (MGW, ProT, Roll, HelT): 0.375 + 16 B/bp
    shapes4 = ("MGW", "ProT", "Roll", "HelT")
    tables, kinds = dnashape.synthetic_tables(shapes4)
    tab = ctx.upload_shapes(tables, kinds)
    ms = []
    for i in range(6):
        ctx.flush_l2()
        ctx.shape_tracks(enc, tab, to_host=False)
        ms.append(ctx.last_kernel_ms())
    t = float(np.mean(ms[3:]))
    out.append(entry("shape_tracks_c16_kernel(+edge fix)", f"{n} bp x 4 tracks (MGW ProT Roll HelT) fp32",
                     16.375 * n, t, bytes_per_bp=16.375))
    tab.close()
    enc.close()
    dev.close()
    # dense scan on config C2: 6 PWMs (L 8,10,12,14,17,20) over 5 000 fg x 200 bp + 50 000 bkg x 204 bp
    fg, bkg = synth.peak_set(SEED, 5000, 200, 10)
    bf, sf, lf = synth.join_records(fg)
    bb, sb, lb = synth.join_records(bkg)
    buf = np.concatenate([bf, np.array([10], np.uint8), bb])
    starts, rlens = np.concatenate([sf, sb + bf.size + 1]), np.concatenate([lf, lb])
    mats = [synth.consensus_pwm(SEED + i, L)[0] for i, L in enumerate((8, 10, 12, 14, 17, 20))]
    lib = ctx.upload_pwms(mats)
    # tile the record set so the outputs exceed L2 several times over (6 x 2 x 4 B x n)
    reps = 8
    big = np.concatenate([buf, np.array([10], np.uint8)] * reps)
    enc = ctx.encode(big)
    units = int(sum(big.size - L + 1 for L in lib.lens))
    ms = []
    for i in range(6):
        ctx.flush_l2()
        ctx.scan_dense(enc, lib, to_host=False)
        ms.append(ctx.last_kernel_ms())
    t = float(np.mean(ms[3:]))
    nbytes = 8.0 * units + 0.375 * big.size
    out.append(entry("scan_dense_kernel", f"C2 x{reps}: 6 PWMs (L 8-20) x {big.size} bp, fwd+rev fp32 scores "
                                          "(int32 fixed-point accumulation)",
                     nbytes, t, bytes_per_unit=8.0, units_per_s=units / t * 1e3))
    enc.close()
    # per-record best windows on config C2 (the PWM feature columns): reads the records, writes S x M x 2 floats;
    # it re-scores every window in double (bit-exact), so it is compute-bound — the HBM fraction is reported as is
    enc = ctx.encode(buf)
    ms = []
    for i in range(6):
        ctx.flush_l2()
        best = ctx.scan_best(enc, lib, starts, rlens)
        ms.append(ctx.last_kernel_ms())
    t = float(np.mean(ms[3:]))
    units2 = int(sum(np.maximum(rlens - L + 1, 0).sum() for L in lib.lens))
    out.append(entry("scan_best_kernel", f"C2: 6 PWMs x 55 000 records ({buf.size} bp), best window per record and strand, "
                                         f"f64 re-scoring; D2H {best.nbytes} B instead of {8 * units2} B of dense scores",
                     0.375 * buf.size + best.nbytes + 16.0 * starts.size, t, units_per_s=units2 / t * 1e3,
                     d2h_bytes=int(best.nbytes)))
    enc.close()
    lib.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from discover_tfbs_b200 import _lib, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    host_group = None
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side group: ranks that wait while rank 0 drives every GPU must not park a spinning NCCL
        # kernel on their device (it would time-slice with rank 0's work there)
        host_group = dist.new_group(backend="gloo")
    n_gpus = world

    numa_before = _lib.bind_thread_near(local)   # pinned buffers of this rank live next to its GPU
    ctx = _lib.Context(local)  # raises if the library or the GPU is missing: no CPU fallback
    logodds, lens, thr = workload(args)
    halo = int(lens.max()) - 1
    lib = ctx.upload_pwms(logodds, lens)

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        ctx.sync()

    def reduce_max(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_sum(*vals):
        if world == 1:
            return [int(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [int(x) for x in t]

    def measure(job):
        """-> dict with whole-job resident and end-to-end numbers (max time over ranks, summed units)."""
        t_res, _, launches = timed(ctx, job.step_resident, args.steps, args.warmup, barrier)
        n_hits = job.n_hits
        t_e2e, wall_e2e, _ = timed(ctx, job.step_e2e, args.steps, args.warmup, barrier)
        d2h = int(job.got.size) * 16 + 8 * args.chunks + job.sites_bytes
        h2d = int(job.n) + 4 * len(lens)
        t_res, t_e2e, wall_e2e = reduce_max(t_res, t_e2e, wall_e2e)
        units, hits, launches, d2h_all, h2d_all = reduce_sum(job.units, n_hits, launches, d2h, h2d)
        return {"value": units * args.steps / t_res, "ms_per_step": t_res / args.steps * 1e3,
                "units_per_step": units, "hits_per_step": hits, "gpu_launches": launches,
                "e2e": {"value": units * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_all,
                        "d2h_bytes_per_step": d2h_all, "ms_per_step": t_e2e / args.steps * 1e3,
                        "host_wall_ms_per_step": wall_e2e}}

    # ---- headline: C4, weak scaling (100 Mb per GPU) ---------------------------------------------------
    per_gpu = int(args.genome_mb * 1e6)
    total = per_gpu * n_gpus
    shard = sharding.plan_shards(total, n_gpus, halo)[rank]
    sampler = ClockSampler(local)
    sampler.start()
    job = HitsJob(ctx, lib, thr, lens, shard, total, args.rate, args.chunks)
    barrier()
    sampler.mark_begin()
    head = measure(job)
    sampler.mark_end()
    clocks = sampler.stop()
    k_scan, k_verify, k_enc = float(np.mean(job.scan_ms[-args.steps:])), float(np.mean(job.verify_ms[-args.steps:])), \
        float(np.mean(job.enc_ms[-args.steps:]))
    k_direct = float(np.mean(job.direct_ms[-args.steps:]))
    n_scan, n_hits_rank = job.n, job.n_hits
    candidates = ctx.last_scan_candidates()
    blocks = lib.hits_blocks()
    direct = lib.hits_direct()

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        mat_bytes = float(sum(2 * 4 * int(L) * 4 for L in lens))
        algo_bytes = 0.375 * n_scan + mat_bytes + 16.0 * n_hits_rank
        # filter scan: one LDS.32 per (lane, base, column group).  Motifs are sorted by descending length and cut
        # into blocks (tfbs_hits_block_plan in csrc/scan_hits.cu): a block of longest motif L uses A 4-column +
        # B 3-column groups (minimal A + B under 192 KB of tables) and holds 32 motifs (16-bit filter fields) or
        # 64 motifs (8-bit fields, plain or sticky; chosen per block from the thresholds by the cost model)
        lookups = float(n_scan) * 32.0 * float(sum(a + b for a, b, _, _ in blocks))
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        smem_peak = 148 * 128 * sm_mhz * 1e6  # bytes/s of shared-memory bandwidth at the sampled clock
        try:
            smem_measured = ctx.measure_smem_gbs()   # conflict-free LDS.32 probe on this GPU, now
        except Exception:
            smem_measured = None
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT,
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 fixed-point filter + f64 re-score (hits)", "data": "synthetic",
            "config": config_dict(args, n_gpus, per_gpu),
            "units_per_step": head["units_per_step"], "hits_per_step": head["hits_per_step"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks,
            "roofline": {
                "kernel": "scan_hits_kernel", "bound": "hbm", "achieved": algo_bytes / k_scan / 1e6, "peak": peak,
                "unit": "GB/s", "frac": algo_bytes / k_scan / 1e6 / peak,
                "traffic": ncu_traffic(args, per_gpu), "peak_source": peak_src,
                "kernel_ms": k_scan, "algorithmic_bytes": algo_bytes,
                "note": "hits mode at 800 motifs is bounded by shared-memory LUT bandwidth, not HBM "
                        "(SURVEY.md §7/§8d): the HBM fraction on the algorithmic bytes is reported as is",
                "smem": {"lookups_per_s": lookups / k_scan * 1e3, "peak_lookups_per_s": 148 * 32 * sm_mhz * 1e6,
                         "lut_bytes_per_launch": lookups * 4.0, "achieved_gbs": lookups * 4.0 / k_scan / 1e6,
                         "peak_gbs": smem_peak / 1e9, "frac": lookups * 4.0 / k_scan / 1e6 / (smem_peak / 1e9),
                         "peak_def": "nominal: 148 SM x 128 B/clk x sampled SM clock",
                         "peak_measured_gbs": smem_measured,
                         "frac_of_measured": (lookups * 4.0 / k_scan / 1e6 / smem_measured) if smem_measured else None,
                         "peak_measured_def": "tfbs_measure_smem_gbs: conflict-free LDS.32 streamed by 148 x 32 warps, "
                                              "best of 3, measured in this run",
                         "loads_per_base": int(sum(a + b for a, b, _, _ in blocks)),
                         "direct_path": {"motifs": direct[0], "max_len": direct[1], "keys": direct[2],
                                         "note": "motifs of <= max_len columns looked up in an enumerated 12-mer hash "
                                                 "(scan_direct_kernel); they cost no table loads"},
                         "blocks": {"wide_32_motifs": sum(1 for b in blocks if b[2] == 0),
                                    "narrow_64_motifs": sum(1 for b in blocks if b[2] == 1),
                                    "sticky_narrow_64_motifs": sum(1 for b in blocks if b[2] == 2)}},
            },
            "kernel_ms": {"encode": k_enc, "scan_direct": k_direct, "scan_hits_filter": k_scan, "verify_candidates": k_verify,
                          "candidates": candidates, "hits": n_hits_rank},
        }

    # ---- host wall clock of the public call on one GPU: warm, cold threshold vector, sorted ---------------
    if not args.no_extras:
        wall = {}
        t0 = time.perf_counter()
        job.step_e2e()
        wall["warm_ms"] = (time.perf_counter() - t0) * 1e3
        thr_cold = (thr + np.float32(1e-3)).astype(np.float32)       # a vector the table cache has not seen
        saved = job.thr
        job.thr = thr_cold
        t0 = time.perf_counter()
        job.step_e2e()
        wall["cold_threshold_vector_ms"] = (time.perf_counter() - t0) * 1e3
        job.thr = saved
        job.step_e2e(sort=True)
        t0 = time.perf_counter()
        job.step_e2e(sort=True)
        wall["warm_sorted_ms"] = (time.perf_counter() - t0) * 1e3
        wall["note"] = ("time.perf_counter() around tfbs_scan_hits_host_shard on rank 0: cold = first call with a new "
                        "threshold vector (block plan + filter tables rebuilt on the host); sorted = device radix sort "
                        "before one D2H copy")
        if line is not None:
            line["e2e_wall"] = wall
    job.close()

    # ---- strong scaling: the SAME 100 Mb genome cut into N shards -----------------------------------------
    if not args.no_extras and per_gpu == C4_TOTAL:
        if n_gpus == 1:
            strong = {"genome_bp_total": C4_TOTAL, "genome_bp_per_gpu": C4_TOTAL, "value": head["value"],
                      "ms_per_step": head["ms_per_step"], "e2e": head["e2e"], "note": "N = 1: identical to the headline"}
        else:
            sh2 = sharding.plan_shards(C4_TOTAL, n_gpus, halo)[rank]
            job2 = HitsJob(ctx, lib, thr, lens, sh2, C4_TOTAL, args.rate, args.chunks)
            barrier()
            r = measure(job2)
            strong = {"genome_bp_total": C4_TOTAL, "genome_bp_per_gpu": sh2.owned, "value": r["value"],
                      "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "units_per_step": r["units_per_step"]}
            job2.close()
        if line is not None:
            line["scaling_strong"] = strong

    # ---- config C5: 3.1 Gb / 8 per GPU + 4 shape tracks ---------------------------------------------------
    if not args.no_extras and (args.c5 == "on" or (args.c5 == "auto" and n_gpus == 8)):
        total5 = C5_PER_GPU * n_gpus
        sh5 = sharding.plan_shards(total5, n_gpus, halo)[rank]
        job5 = HitsJob(ctx, lib, thr, lens, sh5, total5, args.rate, args.chunks, cap_factor=1.5, with_shape=True,
                       shape_sites=262_144)
        barrier()
        r = measure(job5)
        c5 = {"workload": f"C5: {total5} bp synthetic genome x {args.motifs} PWMs + 4 DNA-shape tracks "
                          f"(MGW ProT Roll HelT), {C5_PER_GPU} bp per GPU x {n_gpus}"
                          + ("" if n_gpus == 8 else " (the per-GPU shard of the 8-GPU config)"),
              "value": r["value"], "ms_per_step": r["ms_per_step"], "e2e": r["e2e"],
              "units_per_step": r["units_per_step"], "hits_per_step": r["hits_per_step"],
              "kernel_ms": {"encode": float(np.mean(job5.enc_ms[-args.steps:])),
                            "scan_direct": float(np.mean(job5.direct_ms[-args.steps:])),
                            "scan_hits_filter": float(np.mean(job5.scan_ms[-args.steps:])),
                            "verify_candidates": float(np.mean(job5.verify_ms[-args.steps:])),
                            "shape_tracks_4": float(np.mean(job5.shape_ms[-args.steps:]))},
              "e2e_shape": "4 tracks of the shard computed in HBM (6.2 GB per GPU, not copied); shape feature rows "
                           "(31 values x 4 tracks, reference slicing) of the first 262 144 returned hits copied back"}
        job5.close()
        if line is not None:
            line["c5"] = c5

    # ---- the product's multi-GPU API: ONE process drives all N GPUs (rank 0; the other ranks idle) ----------
    if not args.no_extras:
        barrier()
        if rank == 0:
            try:
                line["product_api"] = product_api(ctx, args, logodds, lens, thr, n_gpus)
            except Exception as exc:
                line["product_api"] = {"error": repr(exc)}
        if host_group is not None:
            dist.barrier(group=host_group)   # the other ranks sleep on the host, their GPUs stay idle
        barrier()

    if numa_before is not None:
        os.sched_setaffinity(0, numa_before)     # the CPU arms below use every core again
    if rank == 0:
        line["numa"] = {"bound_to_gpu_local_cpus": numa_before is not None,
                        "gpu_local_cpus": sorted(_lib.device_local_cpus(local) or [])[:64]}
        if not args.no_extras and n_gpus == 1:
            try:
                line["kernel_rooflines"] = extras(ctx, args, measured_peak_gbs()[0])
            except Exception as exc:  # extras must never hide the headline
                line["kernel_rooflines_error"] = repr(exc)
        if n_gpus == 1:   # the CPU baseline is timed at N = 1 only (one host, whatever N)
            try:
                line["cpu_baseline"] = cpu_pwm_rate(logodds, lens, thr, args.cpu_seconds, host_threads())
            except Exception as exc:
                line["cpu_baseline"] = {"error": repr(exc)}
            if not args.no_extras:
                try:
                    line["cpu_baseline_legs"] = cpu_legs(host_threads())
                except Exception as exc:
                    line["cpu_baseline_legs"] = {"error": repr(exc)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    lib.close()
    ctx.close()


def product_api(ctx, args, logodds, lens, thr, n_gpus):
    """sharding.MultiGpuScanner over `n_gpus` devices from ONE process (one host thread per GPU): a pinned
    n_gpus x 100 Mb host genome, per-device chunk-pipelined host scans, halo dropped on the device; host
    wall clock around scan() (threads start -> every shard's hits in pinned host memory)."""
    from discover_tfbs_b200 import _lib, sharding

    per = int(args.genome_mb * 1e6)
    n = per * n_gpus
    pinned = _lib.PinnedBuffer(n)
    step = 1 << 26
    for lo in range(0, n, step):   # the position-addressable generator runs on the device; the genome lives on the host
        d = ctx.synth_ascii(SEED, lo, min(step, n - lo), gc=0.41)
        pinned.array[lo: min(n, lo + step)] = d.download()
        d.close()
    units = int(sum(max(n - int(L) + 1, 0) for L in lens))
    cap = int(max(4 * 2 * args.rate * units / n_gpus, 1 << 20))
    out = {"genome_bp_total": n, "devices": n_gpus, "threads": n_gpus}
    with sharding.MultiGpuScanner(n, logodds, lens, devices=list(range(n_gpus)), cap_per_device=cap) as sc:
        for label, sort in (("unsorted", False), ("sorted", True)):
            for _ in range(max(args.warmup, 1)):
                sc.scan(pinned.array, thr, sort=sort, chunks=args.chunks, merge=False)
            ws = []
            for _ in range(args.steps):
                parts = sc.scan(pinned.array, thr, sort=sort, chunks=args.chunks, merge=False)
                ws.append(sc.last_ms["scan_wall_ms"])
            hits = int(sum(p.size for p in parts))
            w = float(np.mean(ws))
            out[label] = {"value": units / w * 1e3, "unit": UNIT, "host_wall_ms_per_step": w, "hits": hits,
                          "d2h_bytes_per_step": hits * 16, "h2d_bytes_per_step": n}
        t0 = time.perf_counter()
        merged = sharding.concat_sorted_shards(parts, len(lens))
        out["ordered_concatenation_ms"] = (time.perf_counter() - t0) * 1e3
        out["merged_hits"] = int(merged.size)
    pinned.close()
    out["note"] = ("per-shard results are views into pinned host memory in shard order; joining the sorted shards into "
                   "one (motif, pos, strand) array is an ordered concatenation (no comparison sort), timed separately")
    return out


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line goes to the real stdout; everything else any library prints (NCCL banners,
    warnings) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    import faulthandler

    faulthandler.enable()   # a crash inside the library prints the Python stack to stderr
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C-level stdout of this process (and NCCL's) now lands on stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
