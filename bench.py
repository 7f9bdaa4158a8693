#!/usr/bin/env python
"""bench.py -- JPEG XL decode throughput of the B200 hot path (contract: see the round prompt).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

A "step" = one pass of the hot path over one batch of synthetic input. Workloads:

  vardct4k (default)  BATCH (256) DIFFERENT lossy 3840x2160 VarDCT frames per GPU (BASELINE.json configs[1]; configs[3]'s
                      512 frames are two GPUs' worth) decoded to RGB8: integer translations of the two committed 4K
                      images, encoded at start-up by this repo's GPU encoder (Workload.make_distinct: replicas of one
                      file flatter the lock-step entropy kernels). --replicas: the two committed fixtures themselves
                      (tests/golden/vardct_4k_{natural,synthetic}.jxl, written by the oracle's restatement of libjxl's
                      AcStrategy search: every transform class), alternating; run as a checked side run (`also`).
  encode4k            BATCH (8) RGB8 3840x2160 frames per GPU encoded to lossy VarDCT at distance 1.0 (configs[2]).
  modular             BATCH bench.jxl-shaped frames (2122x1433 lossless Modular RGBA8, 54 groups each): the input of
                      the reference's own criterion bench (jpegxl-rs/benches/decode.rs:10-40).

Every rank works on its own batch (weak scaling: frames are independent); under torchrun the decoded frames of every
step are gathered on rank 0 with NCCL inside the timed region (the one collective of the path).
`e2e` streams on every handle in flight (JxlB200DecoderPlanBatch / CommitPlan / RunToHost: the next step's files are
parsed while the current step's kernels run, the frames leave for pinned host memory wave by wave); --e2e-serial is the
loop without that. The default run folds short side runs of the other workloads into `also`: the fixtures, encode4k, modular and the
lossless encoder (tools/bench_lossless_enc.py).
"""
import argparse
import ctypes
import hashlib
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

UNIT = "Mpx/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Workload:
    def __init__(self, name, batch):
        g = json.load(open(os.path.join(GOLDEN, "golden.json")))
        self.name = name
        if name == "vardct4k":
            names = ["vardct_4k_natural.jxl", "vardct_4k_synthetic.jxl"]
            self.batch = batch or 256
            self.channels = 3
            self.metric = "Mpixels/s decode (lossy VarDCT 3840x2160 -> RGB8)"
            self.dtype = "f32"
            self.sha = [g[n]["sha256_rgb8"] for n in names]
            self.desc = ("decode %d lossy VarDCT 4K frames per GPU (3840x2160, d=1.0, 135 groups/frame; two fixtures "
                         "alternating: bench image tiled to 4K at %.2f bpp, procedural frame at %.2f bpp) to RGB8"
                         % (self.batch, g[names[0]]["bpp"], g[names[1]]["bpp"]))
            self.data = "synthetic (oracle-encoded 4K fixtures; the reference ships no VarDCT file larger than 40x50)"
        elif name == "encode4k":
            names = ["vardct_4k_natural.jxl", "vardct_4k_synthetic.jxl"]
            self.batch = batch or 8
            self.channels = 3
            self.metric = "Mpixels/s encode (RGB8 3840x2160 -> lossy VarDCT, d = 1.0)"
            self.dtype = "f32"
            self.sha = [g[n]["reencoded_sha256"] for n in names]
            self.decoded_sha = [g[n]["sha256_rgb8"] for n in names]
            self.desc = ("encode %d RGB8 4K frames per GPU (3840x2160; the two decoded 4K fixtures alternating) to lossy "
                         "VarDCT at distance 1.0 (libjxl adaptive quant field, variance-heuristic block sizes 8x8 ... 64x64)" % self.batch)
            self.data = "synthetic (the decoded 4K fixtures)"
        else:
            names = ["bench.jxl"]
            self.batch = batch or 256
            self.channels = 4
            self.metric = "Mpixels/s decode (lossless Modular RGBA8, bench.jxl shape)"
            self.dtype = "int32"
            self.sha = [g["bench.jxl"]["sha256"]]
            self.desc = ("decode %d bench.jxl-shaped frames per GPU (2122x1433 lossless Modular RGBA8, 54 groups/frame)"
                         % self.batch)
            self.data = "synthetic (replicas of the reference's bench.jxl)"
        self.blobs = [open(os.path.join(GOLDEN, n), "rb").read() for n in names]
        self.files = [self.blobs[i % len(self.blobs)] for i in range(self.batch)]
        self.distinct = False

    def make_distinct(self, pkg, device):
        """vardct4k: the batch the GPU arm decodes is `batch` DIFFERENT frames (the two base frames
        x integer translations (13 i mod 256, 7 i mod 256) with wrap-around), written at start-up by this repo's GPU
        encoder at distance 1.0 -- a batch of byte-identical replicas lets the lanes of the lock-step entropy kernels
        (32 streams per warp) run perfectly converged, which no real batch does (tools/exp_distinct.py: AC decode 2.9 x
        faster on replicas). The fixtures themselves (all transform classes, libjxl's AcStrategy search by the oracle
        encoder) are decoded in `also.fixture_replicas`."""
        base = pkg.decode_batch(self.blobs, 3, np.uint8, device=device)
        self.fixture_ok = all(hashlib.sha256(b.tobytes()).hexdigest() == s for b, s in zip(base, self.sha))
        enc = pkg.JxlEncoder(quality=1.0, device=device)
        files = []
        for i0 in range(0, self.batch, 16):
            imgs = [np.ascontiguousarray(np.roll(base[i % 2], ((13 * (i // 2)) % 256, (7 * (i // 2)) % 256), axis=(0, 1)))
                    for i in range(i0, min(self.batch, i0 + 16))]
            files += [r.data for r in enc.encode_batch(imgs, epf_iters=1)]
        del enc
        self.files = files
        self.blobs = files[:8]  # what the CPU arm cycles through
        self.distinct = True
        bpp = 8.0 * sum(len(f) for f in files) / (self.batch * base[0].shape[0] * base[0].shape[1])
        self.desc = ("decode %d DIFFERENT lossy VarDCT 4K frames per GPU (3840x2160, d=1.0, 135 groups/frame, %.2f bpp: the "
                     "two base images x integer translations, encoded at start-up by this repo's GPU encoder -- libjxl "
                     "effort-7 heuristics except the AcStrategy search: DCT 8x8 ... 64x64, adaptive quantisation, "
                     "chroma from luma, custom orders, Gaborish + EPF) to RGB8" % (self.batch, bpp))
        self.data = "synthetic (GPU-encoded translations of two 4K images; no two frames share a bitstream)"

    def check(self, outs):
        if self.distinct:  # no offline hashes for frames encoded at start-up: the CPU arm's oracle decode is the check
            self.out_sha = [hashlib.sha256(outs[i].tobytes()).hexdigest() for i in range(min(8, self.batch))]
            return self.fixture_ok
        idx = sorted({0, 1 % self.batch, self.batch // 2, self.batch - 1})
        return all(hashlib.sha256(outs[i].tobytes()).hexdigest() == self.sha[i % len(self.sha)] for i in idx)


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


def cpu_baseline(wl, seconds=15.0, threads=None):
    """The oracle (CPU restatement of libjxl's algorithm; libjxl itself cannot be built here) decoding the workload's
    frames on the host cores: `threads` Python threads (ctypes releases the GIL), one frame at a time each, for about
    `seconds` (every thread finishes the frame it started)."""
    import jxlo
    threads = threads or (os.cpu_count() or 1)
    jxlo.use_fast_build()  # oracle/libjxlo_fast.so: -O3 -march=x86-64-v3, bit-identical (tests/test_oracle_fast_build.py)
    jxlo.lib()
    d = jxlo.Decoded(wl.blobs[0])
    w, h = d.info.xsize, d.info.ysize
    del d
    count = [0] * threads
    stop = time.time() + seconds
    oracle_sha = {}

    def work(i):
        k = i
        while time.time() < stop:
            px = jxlo.Decoded(wl.blobs[k % len(wl.blobs)]).pixels(wl.channels, jxlo.UINT8, raw=True)
            if k < len(wl.blobs) and k not in oracle_sha:
                oracle_sha[k] = hashlib.sha256(np.asarray(px).tobytes()).hexdigest()
            count[i] += 1
            k += 1

    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    frames = sum(count)
    wl.oracle_sha = oracle_sha
    return {"value": frames * w * h / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{frames} {wl.name} frames ({w}x{h}) in {dt:.1f} s on {threads} threads, oracle-CPU (not libjxl): "
                      "the scalar restatement built -O3 -march=x86-64-v3, one frame per thread. libjxl itself (Highway SIMD) "
                      "cannot be built here; its own design figure is ~400 Mpx/s multithreaded decode "
                      "(jpegxl-src/libjxl/doc/xl_overview.md:7-9), so ratios against this arm are upper bounds"}, dt


def cpu_encode_baseline(wl, images, seconds=15.0, threads=None):
    """The oracle's plain encoder on the host cores, one frame at a time per thread."""
    import jxlo
    threads = threads or (os.cpu_count() or 1)
    jxlo.use_fast_build()
    jxlo.lib()
    count = [0] * threads
    stop = time.time() + seconds

    def work(i):
        k = i
        while time.time() < stop:
            jxlo.encode_vardct(images[k % len(images)], distance=1.0, strategy_mode=2, dc_tree=1)
            count[i] += 1
            k += 1

    t0 = time.time()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.time() - t0
    frames = sum(count)
    h, w = images[0].shape[:2]
    return {"value": frames * w * h / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{frames} 4K frames encoded in {dt:.1f} s on {threads} threads, oracle-CPU encoder (not libjxl)"}, dt


def run_encode(args, wl):
    """Encode workload: `value` from the encoder's device events (kernels + host table build between them), `e2e` the
    wall time of the public call with host buffers (H2D of the pixels, kernels, D2H of the sections, assembly)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    base = pkg.decode_batch(wl.blobs, 3, np.uint8, device=local_rank)  # the inputs: our own decode of the fixtures
    ok = all(hashlib.sha256(b.tobytes()).hexdigest() == s for b, s in zip(base, wl.decoded_sha))
    images = [base[i % len(base)] for i in range(wl.batch)]
    enc = pkg.JxlEncoder(quality=1.0, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(1, args.warmup)):
        outs = enc.encode_batch(images)
    ok = ok and all(hashlib.sha256(outs[i].data).hexdigest() == wl.sha[i % len(wl.sha)] for i in range(wl.batch))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, phases = 0.0, {"tokens_ms": 0.0, "host_tables_ms": 0.0, "emit_ms": 0.0}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        enc.encode_batch(images)
        pt = enc.phase_times()
        for k in phases:
            phases[k] += pt[k] / args.steps
        dev_ms += sum(pt.values())
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([dev_ms / args.steps, wall / args.steps * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_wall = float(t[0].item()), float(t[1].item())
    pixels = sum(a.shape[0] * a.shape[1] for a in images) * world
    in_bytes = sum(a.nbytes for a in images)
    out_bytes = sum(len(o.data) for o in outs)
    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        top = max(phases, key=phases.get)
        alg = in_bytes + out_bytes  # pixels read once + compressed bytes written once
        line = {"metric": wl.metric, "value": pixels / (ms_dev * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": wl.data,
                "config": {"workload": wl.desc, "batch_per_gpu": wl.batch,
                           "l2": "working set %.1f GB per step >> 126 MB L2" % (wl.batch * 0.45),
                           "golden_checksum_ok": bool(ok)},
                "clocks": clocks,
                "e2e": {"value": pixels / (ms_wall * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes),
                        "d2h_bytes_per_step": int(out_bytes),
                        "includes": "H2D of the RGB8 frames + kernels + host histogram / table build + D2H of the sections + "
                                    "codestream assembly"},
                "gpu_launches": args.steps * wl.batch * 9,
                "roofline": {"bound": "hbm", "kernel": top, "achieved": alg / (phases[top] * 1e-3) / 1e9, "peak": peak,
                             "peak_kind": peak_kind, "unit": "GB/s", "frac": alg / (phases[top] * 1e-3) / 1e9 / peak,
                             "traffic": None, "algorithmic_bytes_per_launch": int(alg), "kernel_ms": phases[top],
                             "kernel_share_of_step": phases[top] / ms_dev,
                             "step_achieved": alg / (ms_dev * 1e-3) / 1e9, "step_frac": alg / (ms_dev * 1e-3) / 1e9 / peak,
                             "all_kernels_ms": phases}}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], _ = cpu_encode_baseline(wl, base, seconds=args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, args.batch)
    if wl.name == "encode4k":
        raise SystemExit("--impl reference covers the decode workloads; the CPU encoder is timed as cpu_baseline")
    if wl.name == "vardct4k" and not args.replicas:
        # the same frames as the GPU arm decodes: they are written by the GPU encoder at start-up, so this arm needs the
        # device for its set-up too (nothing of the timed work runs there); without one it decodes the fixtures
        try:
            import torch
            if torch.cuda.is_available():
                import __graft_entry__ as ge
                wl.make_distinct(ge.load_package(), int(os.environ.get("LOCAL_RANK", "0")))
        except Exception as e:
            print("reference arm: set-up on the GPU failed (%s); decoding the fixtures" % e, file=sys.stderr)
    # bounded: the whole --steps K --warmup W run stays within a few minutes
    per_step = max(6.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    per_step = min(per_step, max(0.5, args.cpu_seconds))  # (--cpu-seconds below 6: quick contract checks)
    vals, ms = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_baseline(wl, seconds=per_step)
        if i >= args.warmup:
            vals.append(cb["value"])
            ms.append(dt * 1e3)
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": wl.metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": wl.data,
            "config": {"workload": wl.desc,
                       "cpu_arm": "each step decodes as many of these frames as the host cores finish in %.0f s" % per_step,
                       "note": "libjxl cannot be built in this image (Highway / brotli submodules are empty, no Rust "
                               "toolchain); the CPU arm is the scalar oracle restating libjxl 0.11.2, all host threads"},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="vardct4k", choices=["vardct4k", "modular", "encode4k"])
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: 256 vardct4k, 256 modular, 8 encode4k)")
    ap.add_argument("--e2e-plan-threads", type=int, default=0,
                    help="planning threads per handle in the streaming end-to-end loop (default: host cpus / handles in flight)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loop (default: 6 per handle in flight)")
    ap.add_argument("--no-mallopt", action="store_true", help="leave glibc's malloc thresholds alone")
    ap.add_argument("--e2e-serial", action="store_true",
                    help="end-to-end loop without the streaming calls (parse, kernels and read-back of a handle one after the other)")
    ap.add_argument("--inflight", type=int, default=8,
                    help="decoder handles in flight, each with its own buffers and CUDA stream: the latency-bound "
                         "entropy kernels of one batch overlap the per-pixel kernels of the previous one")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", action="store_true",
                    help="vardct4k: decode replicas of the two committed fixtures instead of frames that all differ")
    ap.add_argument("--no-also", action="store_true", help="skip the short encode / lossless side runs folded into `also`")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    if args.workload == "encode4k":
        run_encode(args, Workload(args.workload, args.batch))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (jxl_b200 has no CPU fallback)")
    # Host side of a rank: the CPUs of its GPU's NUMA node, split between the ranks that share the node -- before any
    # thread or pinned buffer exists, so that the planning threads, the CUDA driver's threads and the first touch of the
    # pinned output buffers all land next to the GPU (8 ranks otherwise run 8 x 32 planning threads on 32 cores and copy
    # across sockets).
    host = bind_rank_to_numa(torch, local_rank, local_world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # The planner builds a few hundred MB of tables per batch in std::vectors; glibc would mmap and unmap them batch
    # after batch (page faults on every first touch, mmap-lock traffic between the planning threads of the handles in
    # flight): keep freed memory in the heap instead. An application-level setting, so it is made here, not in the library.
    try:
        if args.no_mallopt:
            raise OSError
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 32 << 20)        # M_MMAP_THRESHOLD: glibc's maximum
        libc.mallopt(-1, (1 << 31) - 1)   # M_TRIM_THRESHOLD: never trim
    except OSError:
        pass
    wl = Workload(args.workload, args.batch)
    if wl.name == "vardct4k" and not args.replicas:
        wl.make_distinct(pkg, local_rank)
    files = wl.files
    plan_threads = max(1, min(host["cpus"], 32))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # handles in flight: at most --inflight and no more than fit into HBM: the first handle's footprint (arenas + output +
    # bitstreams) is measured, the others may take 85 % of what is free after it (and after rank 0's gather buffers);
    # every rank uses the same number
    free0, _ = torch.cuda.mem_get_info()
    decs = [pkg.BatchDecoder(local_rank)]
    decs[0].set_input(files, wl.channels, pkg.JXL_TYPE_UINT8, threads=plan_threads)
    decs[0].run()
    decs[0].wait()  # (the first wait may regrow the token arena)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    per_handle = max(free0 - free1, 1)
    # ---- the one collective of the path: every step's decoded frames are gathered on rank 0, device to
    # device over NVLink (NCCL gather of the handles' output buffers wrapped as CUDA tensors, no host round trip), on
    # NCCL's own stream so that it overlaps the kernels of the next steps
    out_bytes = decs[0].device_output_bytes()
    gather_bufs = None
    if world > 1 and rank == 0:
        gather_bufs = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
        free1, _ = torch.cuda.mem_get_info()
    fit = 1 + int(0.85 * free1 // per_handle)
    cap = max(1, min(args.inflight, args.steps, fit))
    if world > 1:
        t = torch.tensor([cap], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        cap = int(t.item())
    nfl = cap
    if rank == 0 and nfl < args.inflight:
        print("handles in flight: %d (asked for %d; %d steps; %.1f GB of HBM per handle, %.1f GB free)"
              % (nfl, args.inflight, args.steps, per_handle / 1e9, free1 / 1e9), file=sys.stderr)
    decs += [pkg.BatchDecoder(local_rank) for _ in range(nfl - 1)]
    for d in decs[1:]:
        d.set_input(files, wl.channels, pkg.JXL_TYPE_UINT8, threads=plan_threads)
    dec = decs[0]
    tstreams = [torch.cuda.Stream() for _ in range(nfl)]
    torch.cuda.set_stream(tstreams[0])
    streams = [t.cuda_stream for t in tstreams]
    assert all(x != 0 for x in streams)
    out_tensors = [d.device_output_tensor() for d in decs] if world > 1 else None
    gather_work = [None] * nfl

    def gather_step(h):
        """Enqueues the gather of handle h's output after its kernels; the handle's next run waits for it."""
        with torch.cuda.stream(tstreams[h]):
            gather_work[h] = dist.gather(out_tensors[h], gather_bufs, dst=0, async_op=True)

    def run_step(i):
        h = i % nfl
        if gather_work[h] is not None:
            with torch.cuda.stream(tstreams[h]):
                gather_work[h].wait()  # (stream-side wait: the previous gather has read this handle's output)
            gather_work[h] = None
        decs[h].run(streams[h])
        if world > 1:
            gather_step(h)

    # ---- device-resident throughput: bitstreams and tables already in HBM ----
    for _ in range(max(args.warmup, 1)):
        for h, (d, sx) in enumerate(zip(decs, streams)):
            d.run(sx)
            d.wait(sx)  # (the first wait may regrow the token arena and decode again)
            if world > 1:
                gather_step(h)
    st = dec.stats()
    for d in decs:
        d.set_profiling(True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstreams[0])
    for t in tstreams[1:]:
        t.wait_event(e0)
    for i in range(args.steps):  # step i runs on handle i mod inflight; same-handle steps are ordered by its stream
        run_step(i)
    for h in range(nfl):
        if gather_work[h] is not None:
            with torch.cuda.stream(tstreams[h]):
                gather_work[h].wait()
            gather_work[h] = None
    for t in tstreams[1:]:
        tstreams[0].wait_event(t.record_event())
    e1.record(tstreams[0])
    barrier()
    clocks = sampler.stop()
    for d, sx in zip(decs, streams):
        d.wait(sx)
    ms_total = e0.elapsed_time(e1)
    kernel_ms, runs = {}, 0
    for d in decs:
        km, r = d.kernel_times_ex()
        runs += r
        for k, v in km.items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v
        d.set_profiling(False)
    # One more profiled pass with a single handle and nothing else on the GPU: per-kernel times without the overlap of
    # the other handles (in the timed region a kernel's event-to-event time includes waiting for SMs that another
    # handle's kernels hold, so the classes add up to more than the step).
    barrier()
    alone_ms = {}
    decs[0].set_profiling(True)
    for _ in range(2):
        decs[0].run(streams[0])
        decs[0].wait(streams[0])
    km, r = decs[0].kernel_times_ex()
    alone_ms = {k: v / max(r, 1) for k, v in km.items()}
    decs[0].set_profiling(False)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    pixels_per_step = st.pixels * world
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6

    # the gather alone (nothing else on the GPUs): bytes into rank 0 per second, and the check that what arrived is what
    # every rank decoded
    gather_info = None
    if world > 1:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        g0.record(tstreams[0])
        for _ in range(reps):
            with torch.cuda.stream(tstreams[0]):
                dist.gather(out_tensors[0], gather_bufs, dst=0)
        g1.record(tstreams[0])
        barrier()
        t = torch.tensor([g0.elapsed_time(g1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        okg = True
        if rank == 0:
            n0 = dec.out_size(0)
            first_sha = hashlib.sha256(dec.read_output(0).tobytes()).hexdigest()  # (every rank decodes the same batch)
            for r_ in range(world):  # frame 0 of every rank's share, read from rank 0's gather buffer
                okg = okg and hashlib.sha256(gather_bufs[r_][:n0].cpu().numpy().tobytes()).hexdigest() == first_sha
        gather_info = {"collective": "ncclGather of the decoded frames to rank 0 (device to device, every step, inside the "
                                     "timed region, overlapping the next steps' kernels)",
                       "bytes_per_step_into_rank0": int(out_bytes) * (world - 1), "alone_ms": gms,
                       "alone_gbs_into_rank0": out_bytes * (world - 1) / (gms * 1e-3) / 1e9,
                       "nvlink5_gbs_per_direction": 900.0, "gathered_checksum_ok": bool(okg)}

    # ---- end to end through the public API with host buffers ----
    # Every step: host parse + H2D of that step's bitstreams and tables, the kernels, D2H of the pixels into pinned
    # host memory. Steps run on `inflight` handles from as many host threads (ctypes drops the GIL), so the host
    # parse and the copies of one step overlap the kernels of another -- the same pipelining as above.
    # (pinned host memory: one output set per handle in flight, bounded by this rank's share of the free host memory)
    # (one pinned arena per set, laid out like the device output buffer -- frames back to back, 256-byte aligned -- so
    # that the frames of a wave leave in one copy: JxlB200DecoderRunToHost merges neighbours)
    out_base = dec.device_output(0)
    out_offs = [dec.device_output(i) - out_base for i in range(wl.batch)]
    set_bytes = dec.device_output_bytes()

    def pinned_set():
        arena = torch.empty(set_bytes, dtype=torch.uint8).pin_memory().numpy()
        return [arena[o:o + dec.out_size(i)] for i, o in enumerate(out_offs)]

    try:
        import psutil
        share = psutil.virtual_memory().available / max(1, local_world)
    except Exception:
        share = 64e9
    nfl_e2e = max(1, min(nfl, int(0.4 * share // max(set_bytes, 1))))
    if nfl_e2e < nfl and rank == 0:
        print("e2e: %d of %d handles in flight (pinned output sets of %.1f GB each, %.0f GB of host memory per rank)"
              % (nfl_e2e, nfl, set_bytes / 1e9, share / 1e9), file=sys.stderr)
    out_sets = []
    for _ in range(nfl_e2e):
        try:
            out_sets.append(pinned_set())
        except RuntimeError as e:  # the host refuses to pin more: go on with the sets there are
            if rank == 0:
                print("e2e: pinning stopped after %d output sets (%s)" % (len(out_sets), str(e).splitlines()[0]), file=sys.stderr)
            break
    if not out_sets:  # (pageable buffers as the last resort: the copies then go through the driver's staging)
        out_sets.append([np.empty(dec.out_size(i), dtype=np.uint8) for i in range(wl.batch)])
    nfl_e2e = len(out_sets)
    outs = out_sets[0]

    # The ceiling of the read-back: every rank copies its output buffer from HBM into the pinned set at the same time,
    # nothing else running -- PCIe and host-DRAM write bandwidth shared by the ranks of the node. e2e cannot beat
    # d2h_bytes / this.
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        dec.read_outputs(out_sets[0])
    torch.cuda.synchronize()
    d2h_s = (time.perf_counter() - t0) / 2
    t = torch.tensor([d2h_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d2h_s = float(t.item())
    barrier()

    # enough steps for the handles to fall out of lock step (parse / kernels / read-back of different steps overlap);
    # the ramp-up and the drain stay inside the timed region
    # (six per handle: with three, the drain at the end -- the last handles finishing with the GPU half empty -- was a
    # tenth of the timed region and the figure moved by +- 10 % from run to run; profiles/r2_e2e_streaming.txt)
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(6 * nfl_e2e, min(args.steps, 16))

    # A handle streams (include/jxl_b200.h, PlanBatch / CommitPlan / RunToHost): while its kernels decode step k the
    # host parses step k + 1 into the pending plan, and every wave's frames leave for the pinned buffers as soon as they
    # are rendered; after wait() the upload of the pending plan is all that keeps the handle off the GPU. Every step
    # still does all of it: host parse, H2D of bitstreams and tables, kernels, D2H of every pixel.
    phase_s = {"plan": 0.0, "commit_h2d": 0.0, "launch": 0.0, "wait_incl_d2h": 0.0}
    phase_lock = threading.Lock()
    # Planning runs in the shadow of the handle's kernels, so it does not have to be fast, it has to leave the cores to
    # the threads that launch and copy: the handles share the rank's cores instead of each starting one thread per core
    # (measured: 8 handles x 16 planning threads on 16 cores made a launch call take 300 ms).
    stream_threads = args.e2e_plan_threads if args.e2e_plan_threads > 0 else max(1, host["cpus"] // max(nfl_e2e, 1))

    def e2e_round(n):
        it = iter(range(n))
        lock = threading.Lock()

        def take():
            with lock:
                return next(it, None)

        def serial_worker(h):  # --e2e-serial: set_input -> run -> wait -> read_outputs, nothing ahead (the round-1 loop)
            d, sx = decs[h], streams[h]
            while take() is not None:
                d.set_input(files, wl.channels, pkg.JXL_TYPE_UINT8, threads=plan_threads)
                d.run(sx)
                d.wait(sx)
                d.read_outputs(out_sets[h])

        def worker(h):
            if args.e2e_serial:
                return serial_worker(h)
            d, sx = decs[h], streams[h]
            acc = dict.fromkeys(phase_s, 0.0)
            k = take()
            if k is None:
                return
            t0 = time.perf_counter()
            d.plan(files, wl.channels, pkg.JXL_TYPE_UINT8, threads=stream_threads)
            acc["plan"] += time.perf_counter() - t0
            while k is not None:
                t0 = time.perf_counter()
                d.commit()
                t1 = time.perf_counter()
                d.run_to_host(out_sets[h], sx)
                t2 = time.perf_counter()
                k = take()
                if k is not None:
                    d.plan(files, wl.channels, pkg.JXL_TYPE_UINT8, threads=stream_threads)
                t3 = time.perf_counter()
                d.wait(sx)
                t4 = time.perf_counter()
                acc["commit_h2d"] += t1 - t0
                acc["launch"] += t2 - t1
                acc["plan"] += t3 - t2
                acc["wait_incl_d2h"] += t4 - t3
            with phase_lock:
                for key, v in acc.items():
                    phase_s[key] += v

        ts = [threading.Thread(target=worker, args=(h,)) for h in range(nfl_e2e)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    e2e_round(nfl_e2e)  # warm-up
    for key in phase_s:  # (the warm-up round pays the one-time pinned allocations of the second staging buffers)
        phase_s[key] = 0.0
    barrier()
    t0 = time.perf_counter()
    e2e_round(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = pixels_per_step / e2e_s / 1e6
    if rank == 0:  # where a step's wall time goes (summed over the overlapping handles, warm-up included)
        n_e2e = e2e_steps
        print("e2e phases per step (ms): " + ", ".join("%s %.1f" % (k, v / n_e2e * 1e3) for k, v in phase_s.items()),
              file=sys.stderr)

    ok = wl.check(outs)  # checksum gate: the batch decodes to the golden pixels
    okt = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    ok = bool(okt.item())

    # free the HBM and the pinned sets before the side workloads
    del out_sets, outs, decs, dec, out_tensors, gather_bufs
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        per_run = {k: v / max(runs, 1) for k, v in kernel_ms.items()}
        top = max(alone_ms, key=alone_ms.get) if alone_ms else max(per_run, key=per_run.get)
        # algorithmic bytes per launch: one read of the bitstream + one write of the output pixels
        alg_bytes = st.compressed_bytes + st.output_bytes
        top_ms = alone_ms.get(top, per_run[top])  # the dominant kernel's launch alone on the GPU (CUDA events)
        achieved = alg_bytes / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            e = json.load(open(tpath)).get(wl.name, {})
            if e.get("batch") == wl.batch:
                traffic = e.get("dram_bytes_per_step_all_kernels")
        line = {
            "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": wl.data,
            "config": {"workload": wl.desc, "batch_per_gpu": wl.batch, "entropy_streams": int(st.num_streams),
                       "wave_frames": int(st.wave_frames), "handles_in_flight": nfl,
                       "host": host,
                       "l2": "working set %.1f GB per step >> 126 MB L2 (no flush needed)"
                       % ((st.arena_bytes + st.output_bytes + st.compressed_bytes) / 1e9),
                       "golden_checksum_ok": ok},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(st.compressed_bytes), "d2h_bytes_per_step": int(st.output_bytes),
                    "includes": "host parse (threads, a step ahead of the handle's kernels) + H2D bitstreams/tables + kernels + D2H to pinned host (per wave of rendered frames), "
                                "%d steps in flight, %d planning threads per handle" % (nfl_e2e, stream_threads) + (" [--e2e-serial: nothing ahead, read-back after the kernels]" if args.e2e_serial else ""),
                    "d2h_alone_ms_per_step": d2h_s * 1e3,
                    "d2h_ceiling_gbs_per_gpu": st.output_bytes / d2h_s / 1e9,
                    "d2h_ceiling_value": pixels_per_step / d2h_s / 1e6,
                    "frac_of_d2h_ceiling": d2h_s / e2e_s,
                    "note": "d2h_*: the read-back of one step's pixels alone, all ranks at the same time (PCIe + host DRAM "
                            "shared by the node): the end-to-end figure cannot exceed d2h_ceiling_value"},
            "gpu_launches": int(st.kernel_launches) * args.steps,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak,
                         "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": top_ms,
                         "kernel_share_of_step": min(1.0, top_ms / ms_per_step) if ms_per_step else None,
                         "step_achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                         "step_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "all_kernels_ms_overlapped": per_run, "all_kernels_ms_one_handle_alone": alone_ms,
                         "note": "kernel = the class with the longest launch; kernel_ms = that launch alone on the GPU "
                                 "(one handle, CUDA events on its stream); achieved = algorithmic bytes of the launch's batch "
                                 "/ kernel_ms; step_* = the same bytes / ms_per_step (%d handles overlapping, which is why "
                                 "the step is shorter than the sum of the classes); traffic = DRAM bytes of all kernels of "
                                 "one step (ncu, profiles/traffic.json)" % nfl},
        }
        if gather_info:
            line["gather"] = gather_info
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline(wl, seconds=args.cpu_seconds)
            line["cpu_baseline"] = cb
            if wl.distinct:  # the oracle decoded the same frames: its pixels are the check of the GPU's
                common = [k for k in getattr(wl, "oracle_sha", {}) if k < len(wl.out_sha)]
                line["config"]["frames_checked_against_oracle"] = len(common)
                line["config"]["golden_checksum_ok"] = bool(ok and common and all(wl.oracle_sha[k] == wl.out_sha[k] for k in common))
        elif wl.distinct:
            line["config"]["golden_checksum_ok"] = None  # (no oracle run in this invocation; fixtures decoded ok: see fixture_ok)
            line["config"]["fixture_ok"] = bool(ok)
        if world == 1 and not args.no_also and wl.name == "vardct4k":
            line["also"] = side_workloads(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bind_rank_to_numa(torch, local_rank, local_world):
    """Pins this process to its share of the CPUs next to GPU `local_rank` (sysfs local_cpulist of the GPU's PCI device;
    an even split of the allowed CPUs when sysfs has no answer). Returns what it did for the JSON line."""
    allowed = sorted(os.sched_getaffinity(0))
    info = {"cpus_total": len(allowed), "ranks_on_node": local_world}

    def node_cpus(i):
        try:
            p = torch.cuda.get_device_properties(i)
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
            cpus = []
            for part in txt.split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus += list(range(int(a), int(b) + 1))
                elif part:
                    cpus.append(int(part))
            return tuple(c for c in cpus if c in allowed)
        except Exception:
            return tuple()

    mine = tuple(allowed)
    if local_world > 1:
        n = min(local_world, torch.cuda.device_count())
        nodes = [node_cpus(i) for i in range(n)]
        me = nodes[local_rank] if local_rank < n else tuple()
        if me and len(me) < len(allowed):
            sharers = [i for i in range(n) if nodes[i] == me]
            k, m = sharers.index(local_rank), len(sharers)
            per = max(1, len(me) // m)
            mine = me[k * per:(k + 1) * per] if k < m - 1 else me[k * per:]
            info["numa"] = "GPU-local CPUs split between %d ranks" % m
        else:
            per = max(1, len(allowed) // local_world)
            mine = tuple(allowed[local_rank * per:(local_rank + 1) * per]) or tuple(allowed)
            info["numa"] = "even split (one NUMA node or no sysfs answer)"
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            mine = tuple(allowed)
    info["cpus"] = len(mine)
    return info


def side_workloads(args):
    """The other two workloads of the path, short runs in their own processes after the main measurement (same JSON line
    shape, folded into `also`): the encoder (BASELINE.json configs[2]) and the lossless Modular decode of bench.jxl-shaped
    frames (configs[4]; the input of the reference's own criterion bench)."""
    out = {}
    for name, extra in (("fixture_replicas", ["--workload", "vardct4k", "--replicas", "--steps", "8", "--warmup", "3"]),
                        ("encode4k", ["--workload", "encode4k", "--batch", "16", "--steps", "3", "--warmup", "1"]),
                        ("modular", ["--workload", "modular", "--batch", "128", "--steps", "6", "--warmup", "3", "--inflight", "3"])):
        cmd = [sys.executable, os.path.abspath(__file__), "--no-cpu-baseline", "--no-also"] + extra
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not lines:
                out[name] = {"error": (r.stderr or "no output").strip().splitlines()[-1][:300]}
                continue
            j = json.loads(lines[-1])
            out[name] = {k: j.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "e2e", "config")}
            rf = j.get("roofline") or {}
            out[name]["roofline"] = {k: rf.get(k) for k in ("kernel", "achieved", "frac", "step_frac", "kernel_ms")}
        except Exception as e:  # a side workload never takes the main line down
            out[name] = {"error": str(e)[:300]}
    try:  # the lossless (Modular) encoder: 8 4K RGB8 frames per call, host buffers in, codestreams out
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_lossless_enc.py"), "16", "2"], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=240)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        out["lossless_encode4k"] = json.loads(lines[-1]) if r.returncode == 0 and lines else {
            "error": (r.stderr or "no output").strip().splitlines()[-1][:300]}
    except Exception as e:
        out["lossless_encode4k"] = {"error": str(e)[:300]}
    return out


if __name__ == "__main__":
    main()
