"""Condense `ncu -i X.ncu-rep --page raw --csv` + `--page source --csv` into the few numbers the design is argued with."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0     # frames transformed in the launch (for per-frame figures)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), " regs/thread", d.get("launch__registers_per_thread"),
          " dyn smem/CTA", d.get("launch__shared_mem_per_block_dynamic"), " CTAs/SM limit (smem, regs)", d.get("launch__occupancy_limit_shared_mem"),
          d.get("launch__occupancy_limit_registers"))
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed.sum.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg"]
    for k in keys:
        if k in d:
            print("  %-75s %s %s" % (k, d[k], u.get(k, "")))
    print("  warp stall reasons, cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
    st = [(k.split("stalled_")[1].split("_per")[0], float(d[k])) for k in hdr if "issue_stalled" in k and "per_issue_active" in k and d[k] not in ("", "n/a")]
    tot = sum(v for _, v in st)
    for name, v in sorted(st, key=lambda t: -t[1]):
        if v > 0.01:
            print("    %-22s %.3f  (%.1f %%)" % (name, v, 100 * v / tot))
    if frames and "smsp__inst_executed.sum" in d:
        print("  per frame: %.0f warp-instructions, %.0f shared-memory wavefronts, %.0f DRAM bytes" % (
            float(d["smsp__inst_executed.sum"]) / frames, float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]) / frames,
            (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u["dram__bytes_read.sum"], 1) / frames))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    byop = collections.Counter(); tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1]); op = m.group(1) if m else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) else op.split(".")[0]
        ie = int(r[idx["Instructions Executed"]]); byop[op] += ie; tot += ie
    print("  executed SASS by opcode (%d static instructions):" % (len(rows) - 2))
    for op, c in byop.most_common(28):
        print("    %-12s %6.2f %%%s" % (op, 100.0 * c / tot, ("  %7.1f per frame" % (c / frames)) if frames else ""))
