#!/usr/bin/env python3
"""Summarise an ncu report (captured with --set full --import-source on) into a small text file for profiles/:
per kernel the key raw metrics, and per source phase (the '// -----' markers of the kernel sources) the share of
executed instructions and of stall samples.   python tools/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print("# ncu summary of", os.path.basename(rep), "(cold-cache, serialised replays: compare shares, not absolutes)")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("\n## kernel:", d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        if k in d:
            print("  %-82s %s %s" % (k, d[k], u[k]))
    # every other stall reason above 0.05 warps per issue, and the pipe utilisations
    for k in sorted(d):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and k not in KEYS:
            try:
                if float(d[k]) > 0.05:
                    print("  %-82s %s %s" % (k, d[k], u[k]))
            except ValueError:
                pass
    for k in sorted(d):
        if (k.startswith("sm__inst_executed_pipe_") or k.startswith("sm__pipe_")) and k.endswith("pct_of_peak_sustained_active") and k not in KEYS:
            try:
                if float(d[k]) > 2.0:
                    print("  %-82s %s %s" % (k, d[k], u[k]))
            except ValueError:
                pass
    for k in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
              "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed"):
        if k in d:
            print("  %-82s %s %s" % (k, d[k], u[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern, cur, data = None, None, {}
for r in rows:
    if r and r[0] == "Function Name":
        kern = r[1][:60]
    elif r and r[0] == "File Path":
        cur = r[1]
    elif cur and kern and len(r) > 10 and r[0].isdigit():
        try:
            data.setdefault(kern, {}).setdefault(cur, {})[int(r[0])] = (int(r[6]), int(r[7]))
        except ValueError:
            pass
for kern, files in data.items():
    tot = sum(v[0] for f in files.values() for v in f.values()) or 1
    toti = sum(v[1] for f in files.values() for v in f.values()) or 1
    print("\n## source phases of", kern, "(share of stall samples / of executed warp instructions)")
    for f, d in files.items():
        local = f.replace("/root/repo", ROOT)
        if not os.path.exists(local) or "csrc" not in f:
            continue
        lines = open(local).read().split("\n")
        marks = [(i + 1, l.strip().lstrip("/ -=").strip()[:70]) for i, l in enumerate(lines)
                 if l.strip().startswith("// ----------------") or l.strip().startswith("// ================")]
        marks = [(1, "(head)")] + marks + [(len(lines) + 1, "end")]
        for (a, nm), (b, _) in zip(marks, marks[1:]):
            s = sum(v[0] for l, v in d.items() if a <= l < b)
            ii = sum(v[1] for l, v in d.items() if a <= l < b)
            if s * 200 > tot or ii * 200 > toti:
                print("  %-22s %5d-%5d %-72s samples %5.1f%%  inst %5.1f%%" % (os.path.basename(f), a, b, nm, 100.0 * s / tot, 100.0 * ii / toti))
    # the hottest source lines (stall samples), with their share of executed warp instructions
    flat = [(v[0], v[1], f, l) for f, d in files.items() if "csrc" in f for l, v in d.items()]
    flat.sort(reverse=True)
    print("\n## hottest source lines of", kern, "(samples %, inst %, file:line, text)")
    for smp, ins, f, l in flat[:45]:
        local = f.replace("/root/repo", ROOT)
        text = ""
        if os.path.exists(local):
            ls = open(local).read().split("\n")
            text = ls[l - 1].strip()[:110] if l - 1 < len(ls) else ""
        print("  %5.2f%% %5.2f%%  %s:%d  %s" % (100.0 * smp / tot, 100.0 * ins / toti, os.path.basename(f), l, text))
