"""Turns the round-2 raw outputs in gpurun_out/ (scratch) into the tracked summaries under profiles/:
  profiles/r2_experiments.md          every bench A/B of the round, one line each
  profiles/r2_ncu_<kernel>.txt        tools/ncu_summary.py of the round's `ncu --set full` captures
  profiles/r2_interference_*.jsonl    tools/interference.py runs (foreground class next to background handles)
"""
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

NOTES = {
    "r2_t0_bench": "start of round 2 (round-1 kernels, replicas of two fixture files)",
    "r2a_bench": "bench line reworked (consistent roofline, NCCL gather, -O3 oracle arm)",
    "r2b_base": "max-shared carve-out preference on every decode kernel", "r2b_nocarve": "... without it",
    "r2b_nonw": "(y, N, W) table path off (JXLB200_NO_NW_LUT=1)",
    "r2d_base": "AC decode 8 warps per CTA", "r2d_onewarp": "AC decode 1 warp per CTA (kept)",
    "r2e_base": "no SM partition", "r2e_sm24": "green contexts: 24 SMs entropy / 124 pixel",
    "r2e_sm32": "green contexts: 32 / 116", "r2e_sm48": "green contexts: 48 / 100",
    "r2g_base": "batch = 256 DIFFERENT GPU-encoded frames (from here on)", "r2g_frame": "CTA per (frame, pass) AC kernel, TMA-staged tables",
    "r2h_n2": "two ranks under torchrun: NCCL gather of the decoded frames inside the timed region",
    "r2i_coop": "k_modular_decode_coop v1 (generic rows, 1 warp per CTA)", "r2i_nocoop": "lock-step kernels (JXLB200_NO_COOP=1)",
    "r2j_coop": "coop v2 (fast rows only without misses -> not taken by the bench streams)", "r2j_nocoop": "lock-step",
    "r2k_coop": "coop v3 (fast rows, speculate-then-repair), 8 handles", "r2k_coop_if6": "... 6 handles", "r2k_coop_if4": "... 4 handles",
    "r2l_turns": "coop v3 + per-pixel phases take turns", "r2l_noturns": "coop v3, no turns",
    "r2l_turns_nocoop": "turns + lock-step DC kernel", "r2l_turns_acw4": "turns + AC decode 4 warps per CTA",
    "r2l_turns_acframe": "turns + CTA per (frame, pass) AC kernel",
}


def short(d):
    return ", ".join("%s %.0f" % (k.replace("_decode", "").replace("dequant_", ""), v) for k, v in d.items() if v > 0.5)


lines = ["# Round-2 bench experiments (B200, `python bench.py --no-cpu-baseline --no-also`, 256 lossy 4K frames per step)", "",
         "Raw lines: gpurun_out/ (scratch) at the time; this table is the tracked copy. `alone` / `overlapped` = per-class kernel ms",
         "of one step with one handle on the GPU / inside the timed region with all handles in flight.", "",
         "| run | what | Mpx/s device | Mpx/s e2e | ms/step | alone | overlapped |", "|---|---|---|---|---|---|---|"]
for f in sorted(glob.glob(os.path.join(G, "r2*.json"))):
    name = os.path.basename(f)[:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    if d.get("impl") == "reference" or "roofline" not in d:
        continue
    r = d["roofline"]
    lines.append("| %s | %s | %.0f | %.0f | %.1f | %s | %s |" % (
        name, NOTES.get(name, ""), d["value"], d["e2e"]["value"], d["ms_per_step"],
        short(r.get("all_kernels_ms_one_handle_alone", {})), short(r.get("all_kernels_ms_overlapped", {}))))
open(os.path.join(P, "r2_experiments.md"), "w").write("\n".join(lines) + "\n")

for rep, out in [("r2c_ncu_k_ac_decode", "r2_ncu_k_ac_decode_b8"), ("r2c_ncu_k_dequant_idct", "r2_ncu_k_dequant_idct_b8"),
                 ("r2c_ncu_k_render_fused", "r2_ncu_k_render_fused_b8"), ("r2i_ncu_coop", "r2_ncu_k_modular_decode_coop_v1_wp_tree_b8"),
                 ("r2k_ncu_coop", "r2_ncu_k_modular_decode_coop_v3_b16"), ("r2m_ncu_coop", "r2_ncu_k_modular_decode_coop_final_b16")]:
    src = os.path.join(G, rep + ".ncu-rep")
    if os.path.exists(src):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), src], capture_output=True, text=True).stdout
        open(os.path.join(P, out + ".txt"), "w").write("# ncu --set full --clock-control none, summary by tools/ncu_summary.py of %s.ncu-rep\n%s" % (rep, txt))
for f in glob.glob(os.path.join(G, "r2*_intf.jsonl")):
    shutil.copy(f, os.path.join(P, "r2_interference_" + os.path.basename(f).split("_")[0] + ".jsonl"))
print("\n".join(lines[-30:]))
