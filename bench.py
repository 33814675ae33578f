#!/usr/bin/env python3
"""bench.py — motif x bp scored per second on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (libtfbs_cuda.so)
    python bench.py --impl reference --gpus N --steps K ...   # host-CPU reference arm

Workload (config C4 of BASELINE.json, weak scaling): every GPU scans its own 100 Mb chunk
(+ Lmax-1 halo) of one N x 100 Mb position-addressable synthetic genome against an 800-PWM
JASPAR-sized library (lengths 6..30), both strands, thresholded hits (expected rate 1e-4 per
position per strand).  One unit = one motif evaluated at one start position on both strands.

A step = encode (ASCII -> 2-bit + mask) + hits-mode scan of the chunk.
  value : inputs (ASCII bytes) already resident in HBM, hits left in HBM; CUDA events on the
          library's stream; L2 flushed (untimed) before every step.
  e2e   : the same step through the public API with HOST buffers (tfbs_scan_hits_host): pinned
          ASCII -> H2D -> encode -> scan -> hits + count D2H into pinned memory, chunk-pipelined on
          three streams, all inside the timed region.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=100.0, help="Mb scanned per GPU")
    ap.add_argument("--motifs", type=int, default=800)
    ap.add_argument("--rate", type=float, default=1e-4, help="expected hit rate per position per strand")
    ap.add_argument("--chunks", type=int, default=8, help="chunks of the pipelined end-to-end pass")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel roofline section")
    ap.add_argument("--with-shape", action="store_true",
                    help="config C5 flavour: every step also computes 4 DNA-shape tracks of the chunk")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU budget of the cpu_baseline sample")
    return ap.parse_args()


def workload(args):
    from discover_tfbs_b200 import synth

    logodds, lens = synth.pwm_library(SEED, args.motifs, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=args.rate, gc=0.41, seed=SEED, sample=100_000)
    return logodds, lens, thr


def config_dict(args, n_gpus, per_gpu):
    return {
        "workload": ("C4: 800-PWM JASPAR-sized library (L 6-30) x 100 Mb synthetic genome per GPU, "
                     "both strands, thresholded hits" if not (args.with_shape or per_gpu != 100_000_000) else
                     f"C5 flavour: {args.motifs} PWMs x {per_gpu} bp per GPU, hits"
                     + (" + 4 DNA-shape tracks" if args.with_shape else "")),
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


def cpu_sample_rate(logodds, lens, thr, seconds, gc=0.41):
    """Time the oracle port (all host threads) on a bounded sample of the same workload."""
    from discover_tfbs_b200 import synth
    from oracle import pssm_oracle

    probe = synth.genome(SEED, 0, 4096, gc=gc)
    t0 = time.perf_counter()
    _, _, threads = pssm_oracle.scan_count(probe, logodds, lens, thr, threads=0)
    dt = max(time.perf_counter() - t0, 1e-4)
    units_probe = int(sum(max(probe.size - int(L) + 1, 0) for L in lens))
    n = int(min(max(4096 * seconds / dt, 8192), 50e6))
    sample = synth.genome(SEED, 0, n, gc=gc)
    t0 = time.perf_counter()
    hits, csum, threads = pssm_oracle.scan_count(sample, logodds, lens, thr, threads=0)
    dt = time.perf_counter() - t0
    units = int(sum(max(n - int(L) + 1, 0) for L in lens))
    return {"value": units / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} bp x {len(lens)} PWMs both strands (oracle C port of Biopython _pwm.c "
                      f"semantics, OpenMP), {dt:.1f} s, {hits} hits; probe {units_probe} units"}


def run_reference(args):
    """Host-CPU reference arm.  Biopython (the PWM scorer the north star names) and DNAshapeR are
    not installable here and the reference has no PWM code of its own, so the arm times the oracle
    port of their algorithm with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from discover_tfbs_b200 import synth
    from oracle import pssm_oracle

    logodds, lens, thr = workload(args)
    per_gpu = int(args.genome_mb * 1e6)
    probe = synth.genome(SEED, 0, 4096)
    t0 = time.perf_counter()
    pssm_oracle.scan_count(probe, logodds, lens, thr)
    dt = max(time.perf_counter() - t0, 1e-4)
    total_steps = max(args.steps + args.warmup, 1)
    n = int(min(max(4096 * (90.0 / total_steps) / dt, 8192), 20e6))  # whole run ~90 s
    sample = synth.genome(SEED, 0, n)
    units = int(sum(max(n - int(L) + 1, 0) for L in lens))
    for _ in range(args.warmup):
        pssm_oracle.scan_count(sample, logodds, lens, thr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits, _, threads = pssm_oracle.scan_count(sample, logodds, lens, thr)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = units / dt
    desc = f"{n} bp x {len(lens)} PWMs both strands per step, {threads} threads, {hits} hits"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 accumulate -> f32", "data": "synthetic",
        "config": config_dict(args, args.gpus, per_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = oracle C port (Biopython PSSM semantics); the reference repo has no "
                "PWM scanner of its own and Biopython/DNAshapeR are absent from the image",
    }
    emit(line)


def extras(ctx, args, peak):
    """Per-kernel HBM rooflines for the streaming kernels (encode, dense scan on config C2,
    shape tracks), each timed alone with CUDA events after warm-up, L2 flushed."""
    from discover_tfbs_b200 import dnashape, synth

    out = []
    # streaming kernels are timed on the per-GPU shard of config C5 (3.1 Gb / 8 GPUs = 387.5 Mb)
    n = int(387.5e6 if args.genome_mb >= 100.0 else args.genome_mb * 1e6)
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
    buf = np.concatenate([synth.join_records(fg)[0], np.array([10], np.uint8), synth.join_records(bkg)[0]])
    rng = np.random.default_rng(SEED)
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
    out.append(entry("scan_dense_kernel", f"C2 x{reps}: 6 PWMs (L 8-20) x {big.size} bp, fwd+rev fp32 scores",
                     nbytes, t, bytes_per_unit=8.0, units_per_s=units / t * 1e3))
    enc.close()
    lib.close()
    del rng
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from discover_tfbs_b200 import _lib, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    ctx = _lib.Context(local)  # raises if the library or the GPU is missing: no CPU fallback
    logodds, lens, thr = workload(args)
    per_gpu = int(args.genome_mb * 1e6)
    total = per_gpu * n_gpus
    halo = int(lens.max()) - 1
    shard = sharding.plan_shards(total, n_gpus, halo)[rank]
    units = sharding.units_of_shard(shard, lens, total)
    n_scan = shard.scanned

    lib = ctx.upload_pwms(logodds, lens)
    dev_ascii = ctx.synth_ascii(SEED, shard.start, n_scan, gc=0.41)   # position-addressable shard
    enc = ctx.encode_device(dev_ascii.ptr, n_scan)
    # capacity: 4x the expected number of hits
    cap = int(max(4 * 2 * args.rate * units, 1 << 20))
    n_hits = ctx.scan_hits(enc, lib, thr, cap=cap, to_host=False)     # also sizes the scratch
    pinned_in = _lib.PinnedBuffer(n_scan)
    pinned_in.array[:] = dev_ascii.download()
    pinned_out = _lib.PinnedBuffer(cap * 16)
    hits_out = pinned_out.view(_lib.HIT_DTYPE)

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        ctx.sync()

    shape_tab = None
    if args.with_shape:
        from discover_tfbs_b200 import dnashape
        st, sk = dnashape.synthetic_tables(("MGW", "ProT", "Roll", "HelT"))
        shape_tab = ctx.upload_shapes(st, sk)
        ctx.shape_tracks(enc, shape_tab, to_host=False)

    def step_resident():
        enc.reencode_device(dev_ascii.ptr)
        return ctx.scan_hits(enc, lib, thr, cap=cap, to_host=False)

    def step_e2e():
        # public end-to-end call: pinned host ASCII -> (H2D | encode + scan | hits D2H) pipelined
        # over args.chunks chunks on three streams -> hits in pinned host memory
        return ctx.scan_hits_host(pinned_in.array, enc, lib, thr, hits_out[:cap], chunks=args.chunks)

    # ---- resident (value) ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        ctx.flush_l2()
        step_resident()
    barrier()
    sampler.mark_begin()
    launches0 = ctx.total_kernel_launches()
    step_ms, scan_ms, enc_ms, shape_ms = [], [], [], []
    for _ in range(args.steps):
        ctx.flush_l2()
        launches_flush = ctx.total_kernel_launches()
        ctx.timer_start()
        enc.reencode_device(dev_ascii.ptr)
        enc_ms.append(ctx.last_kernel_ms())
        n_hits = ctx.scan_hits(enc, lib, thr, cap=cap, to_host=False)
        scan_ms.append(ctx.last_kernel_ms())
        if shape_tab is not None:
            ctx.shape_tracks(enc, shape_tab, to_host=False)
            shape_ms.append(ctx.last_kernel_ms())
        step_ms.append(ctx.timer_stop())
        del launches_flush
    launches = ctx.total_kernel_launches() - launches0
    barrier()
    t_resident = float(np.sum(step_ms)) / 1e3

    # ---- end to end --------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        ctx.flush_l2()
        ctx.timer_start()
        got = step_e2e()
        e2e_ms.append(ctx.timer_stop())
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    t_e2e = float(np.sum(e2e_ms)) / 1e3
    d2h = int(got.size) * 16 + 8
    h2d = int(n_scan) + 8 * len(lens)

    # ---- reduce over ranks: max time, summed units ---------------------------------------------------
    if world > 1:
        t = torch.tensor([t_resident, t_e2e], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_resident, t_e2e = float(t[0]), float(t[1])
        u = torch.tensor([units, n_hits, launches], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        units_all, hits_all, launches_all = int(u[0]), int(u[1]), int(u[2])
    else:
        units_all, hits_all, launches_all = units, n_hits, launches

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k_ms = float(np.mean(scan_ms))
        mat_bytes = float(sum(2 * 4 * int(L) * 4 for L in lens))
        algo_bytes = 0.375 * n_scan + mat_bytes + 16.0 * n_hits
        # hits kernel: one LDS.32 per (lane, base, column group).  Motifs are sorted by descending
        # length and cut into blocks (tfbs_hits_block_plan in csrc/scan_hits.cu): a block of longest
        # motif L uses A 4-column + B 3-column groups (minimal A + B under 192 KB of tables) and holds
        # 32 motifs (16-bit filter fields) or 64 motifs (8-bit fields, plain or sticky; chosen per block
        # from the thresholds by the cost model in tfbs_build_hits_tables)
        blocks = lib.hits_blocks()
        lookups = float(n_scan) * 32.0 * float(sum(a + b for a, b, _, _ in blocks))
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        smem_peak = 148 * 128 * sm_mhz * 1e6  # bytes/s of shared-memory bandwidth at the sampled clock
        line = {
            "metric": METRIC, "value": units_all * args.steps / t_resident, "unit": UNIT,
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_resident / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 filter + f64 re-score (hits), f32 (dense)", "data": "synthetic",
            "config": config_dict(args, n_gpus, per_gpu),
            "units_per_step": units_all, "hits_per_step": hits_all,
            "e2e": {"value": units_all * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": {
                "kernel": "scan_hits_kernel", "bound": "hbm", "achieved": algo_bytes / k_ms / 1e6, "peak": peak,
                "unit": "GB/s", "frac": algo_bytes / k_ms / 1e6 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this size, from the
                # committed ncu --set full capture profiles/r01/ncu_full_hits_narrow_100Mb_raw.csv (part of
                # the 297 MB of hits is still in L2 when the kernel ends)
                "traffic": 62.62e6 + 106.75e6 if (args.motifs == 800 and per_gpu == 100_000_000
                                                  and abs(args.rate - 1e-4) < 1e-12) else None,
                "peak_source": peak_src,
                "kernel_ms": k_ms, "algorithmic_bytes": algo_bytes,
                "note": "hits mode at 800 motifs is bounded by shared-memory LUT bandwidth, not HBM "
                        "(SURVEY.md §7/§8d): the HBM fraction on the algorithmic bytes is reported as is",
                "smem": {"lookups_per_s": lookups / k_ms * 1e3, "peak_lookups_per_s": 148 * 32 * sm_mhz * 1e6,
                         "lut_bytes_per_launch": lookups * 4.0, "achieved_gbs": lookups * 4.0 / k_ms / 1e6,
                         "peak_gbs": smem_peak / 1e9, "frac": lookups * 4.0 / k_ms / 1e6 / (smem_peak / 1e9),
                         "peak_def": "148 SM x 128 B/clk x sampled SM clock",
                         "blocks": {"wide_32_motifs": sum(1 for b in blocks if b[2] == 0),
                                    "narrow_64_motifs": sum(1 for b in blocks if b[2] == 1),
                                    "sticky_narrow_64_motifs": sum(1 for b in blocks if b[2] == 2)}},
            },
            "kernel_ms": {"encode": float(np.mean(enc_ms)), "scan_hits": k_ms,
                          **({"shape_tracks_4": float(np.mean(shape_ms))} if shape_ms else {})},
        }
        if not args.no_extras:
            try:
                line["kernel_rooflines"] = extras(ctx, args, peak)
            except Exception as exc:  # extras must never hide the headline
                line["kernel_rooflines_error"] = repr(exc)
        try:
            line["cpu_baseline"] = cpu_sample_rate(logodds, lens, thr, args.cpu_seconds)
        except Exception as exc:
            line["cpu_baseline"] = {"error": repr(exc)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


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
