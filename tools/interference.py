#!/usr/bin/env python
"""How much does one kernel class slow another one down when they share the GPU?

Handle A times its `fg` kernel classes (CUDA events on its stream) while N other handles loop their `bg` classes on
their own streams. Uses the profiling hook JxlB200DecoderSetPhaseMask: after one full Run every class can run again on
the data the others left. Prints one JSON line per experiment.

    python tools/interference.py [--batch 256] [--bg-handles 3]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--bg-handles", type=int, default=3)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--quick", action="store_true", help="only: per-pixel classes next to each entropy class, --bg-handles handles")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    wl = bench.Workload("vardct4k", args.batch)
    n = 1 + args.bg_handles
    decs = [pkg.BatchDecoder(0) for _ in range(n)]
    tstreams = [torch.cuda.Stream() for _ in range(n)]
    streams = [t.cuda_stream for t in tstreams]
    for d, s in zip(decs, streams):
        d.set_input(wl.files, 3, pkg.JXL_TYPE_UINT8)
        d.run(s)
        d.wait(s)
        d.run(s)
        d.wait(s)
    PIX = ["dequant_idct", "filters"]
    ENT = ["modular_decode", "dc_finish", "ac_decode"]

    def measure(fg, bg, nbg):
        """ms per Run of classes `fg` on handle 0 while handles 1..nbg loop classes `bg`."""
        stop = threading.Event()

        def loop(i):
            decs[i].set_phase_mask(bg)
            while not stop.is_set():
                decs[i].run(streams[i])
                torch.cuda.current_stream().synchronize() if False else None
                tstreams[i].synchronize()

        ts = [threading.Thread(target=loop, args=(i,)) for i in range(1, 1 + nbg)] if bg else []
        for t in ts:
            t.start()
        time.sleep(0.5 if ts else 0.0)
        decs[0].set_phase_mask(fg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        decs[0].run(streams[0])
        tstreams[0].synchronize()
        e0.record(tstreams[0])
        for _ in range(args.reps):
            decs[0].run(streams[0])
        e1.record(tstreams[0])
        tstreams[0].synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        stop.set()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        for d in decs:
            d.set_phase_mask(None)
        return ms

    exps = []
    for fg in (PIX, ["dequant_idct"], ["filters"], ENT, ["modular_decode"], ["ac_decode"], ["dc_finish"]):
        exps.append((fg, None, 0))
    for nbg in sorted({1, args.bg_handles}):
        for fg in (PIX, ["dequant_idct"], ["filters"]):
            for bg in (["modular_decode"], ["ac_decode"], ["dc_finish"], ENT, PIX):
                exps.append((fg, bg, nbg))
        for fg in (["modular_decode"], ["ac_decode"]):
            for bg in (PIX, ENT):
                exps.append((fg, bg, nbg))
    if args.quick:
        exps = [(PIX, None, 0), (ENT, None, 0), (["modular_decode"], None, 0), (["ac_decode"], None, 0), (["dc_finish"], None, 0)]
        exps += [(PIX, bg, args.bg_handles) for bg in (["modular_decode"], ["ac_decode"], ["dc_finish"], ENT)]
        exps += [(ENT, PIX, args.bg_handles), (ENT, PIX, 1)]
    base = {}
    for fg, bg, nbg in exps:
        ms = measure(fg, bg, nbg)
        key = "+".join(fg)
        if bg is None:
            base[key] = ms
        print(json.dumps({"fg": key, "bg": "+".join(bg) if bg else None, "bg_handles": nbg, "ms": round(ms, 2),
                          "slowdown": round(ms / base[key], 2) if key in base else None}), flush=True)


if __name__ == "__main__":
    main()
