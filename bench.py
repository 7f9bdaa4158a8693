#!/usr/bin/env python
"""bench.py -- JPEG XL decode throughput of the B200 hot path (contract: see the round prompt).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

A "step" = one pass of the hot path over one batch of synthetic input: BATCH independent
bench.jxl-shaped frames (2122x1433 lossless Modular RGBA8, 54 groups each), i.e. the workload the
reference's own criterion bench decodes (jpegxl-rs/benches/decode.rs:10-40), replicated so that the
#groups x #frames parallelism fills the GPU. Every rank works on its own batch (weak scaling, no
data-path collective; the final gather of pixels is left to the caller).
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

METRIC = "Mpixels/s decode (lossless Modular RGBA8, bench.jxl shape)"
UNIT = "Mpx/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(data, seconds=15.0, threads=None):
    """The oracle (CPU restatement of libjxl's algorithm; libjxl itself cannot be built here) decoding the same
    frame on the host cores: `threads` Python threads (ctypes releases the GIL), one frame each, for ~`seconds`."""
    import jxlo
    threads = threads or (os.cpu_count() or 1)
    jxlo.lib()
    w = h = 0
    d = jxlo.Decoded(data)
    w, h = d.info.xsize, d.info.ysize
    del d
    count = [0] * threads
    stop = time.time() + seconds

    def work(i):
        while time.time() < stop:
            jxlo.Decoded(data).pixels(4, jxlo.UINT8, raw=True)
            count[i] += 1

    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    frames = sum(count)
    return {"value": frames * w * h / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{frames} bench.jxl frames (2122x1433 RGBA8) in {dt:.1f} s, oracle-CPU (not libjxl)"}, dt / max(frames, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data = open(os.path.join(ROOT, "tests", "golden", "bench.jxl"), "rb").read()
    # bounded: steps x ~ (15 s / steps) so that the whole run stays within a couple of minutes
    per_step = max(2.0, min(15.0, 60.0 / max(args.steps + args.warmup, 1)))
    vals, ms = [], []
    for i in range(args.warmup + args.steps):
        cb, per_frame = cpu_baseline(data, seconds=per_step)
        if i >= args.warmup:
            vals.append(cb["value"])
            ms.append(per_step * 1e3)
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic (replicas of the reference's bench.jxl)",
            "config": {"workload": "decode bench.jxl-shaped frames (2122x1433 lossless Modular RGBA8) on host cores",
                       "note": "libjxl cannot be built in this image (Highway/brotli submodules empty); "
                               "the CPU arm is the scalar oracle restating libjxl 0.11.2"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (jxl_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    data = open(os.path.join(ROOT, "tests", "golden", "bench.jxl"), "rb").read()
    files = [data] * args.batch
    dec = pkg.BatchDecoder(local_rank)
    dec.set_input(files, 4, pkg.JXL_TYPE_UINT8)
    st = dec.stats()
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: inputs and tables already in HBM ----
    for _ in range(args.warmup):
        dec.run(stream)
    dec.wait(stream)
    dec.set_profiling(True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dec.run(stream)
    e1.record()
    barrier()
    clocks = sampler.stop()
    dec.wait(stream)
    ms_total = e0.elapsed_time(e1)
    kernel_ms, runs = dec.kernel_times()
    dec.set_profiling(False)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    pixels_per_step = st.pixels * world
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ----
    outs = [torch.empty(dec.out_size(i), dtype=torch.uint8).pin_memory().numpy() for i in range(args.batch)]
    e2e_steps = max(2, min(args.steps, 5))
    for i in range(1 + e2e_steps):
        if i == 1:
            barrier()
            t0 = time.perf_counter()
        dec.set_input(files, 4, pkg.JXL_TYPE_UINT8)  # host parse + H2D of bitstreams and tables
        dec.run(stream)
        dec.wait(stream)
        dec.read_outputs(outs)                        # D2H into pinned host memory
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = pixels_per_step / e2e_s / 1e6

    # checksum gate: the batch decodes to the golden pixels
    import hashlib
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["bench.jxl"]["sha256"]
    ok = all(hashlib.sha256(outs[i].tobytes()).hexdigest() == g for i in (0, args.batch // 2, args.batch - 1))

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        # algorithmic bytes of the dominant kernel per launch: bitstream read + pixels written
        alg_bytes = st.compressed_bytes + st.output_bytes
        dec_ms = kernel_ms[0] / max(runs, 1)
        achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("batch") == args.batch:
                traffic = tj.get("k_modular_decode_dram_bytes")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic (replicas of the reference's bench.jxl)",
            "config": {"workload": f"decode {args.batch} bench.jxl-shaped frames per GPU (2122x1433 lossless Modular "
                                   f"RGBA8, 54 groups/frame, {st.num_streams} entropy streams)",
                       "batch_per_gpu": args.batch, "l2": "working set %.1f GB per step >> 126 MB L2 (no flush needed)"
                       % ((st.arena_bytes + st.output_bytes + st.compressed_bytes) / 1e9),
                       "golden_checksum_ok": bool(ok)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(st.compressed_bytes), "d2h_bytes_per_step": int(st.output_bytes),
                    "includes": "host parse (threads) + H2D bitstreams/tables + kernels + D2H to pinned host"},
            "gpu_launches": int(st.kernel_launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_modular_decode", "achieved": achieved, "peak": peak,
                         "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": dec_ms,
                         "kernel_share_of_step": dec_ms / ms_per_step if ms_per_step else None,
                         "all_kernels_ms": {"modular_decode": kernel_ms[0] / max(runs, 1),
                                            "group_programs": kernel_ms[1] / max(runs, 1),
                                            "frame_levels": kernel_ms[2] / max(runs, 1),
                                            "write_output": kernel_ms[3] / max(runs, 1)}},
        }
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline(data, seconds=args.cpu_seconds)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
