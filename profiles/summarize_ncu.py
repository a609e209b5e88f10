#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into the small text/JSON files kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/ncu_r01a_classify [--reads N] [--workload NAME]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max"]


def ncu_csv(rep, *page):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(page), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    reads = int(sys.argv[sys.argv.index("--reads") + 1]) if "--reads" in sys.argv else None
    workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else None
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    summary = {"report": rep, "kernels": []}
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
                try:
                    d[h] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[h] = vals[i]
                d.setdefault("_units", {})[h] = units[i]
        summary["kernels"].append(d)
    k0 = summary["kernels"][0]
    if reads:
        to_b = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        dram = sum(k0[m] * to_b[k0["_units"][m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        summary.update(workload=workload, reads_per_launch=reads, dram_bytes_per_launch=dram, dram_bytes_per_read=dram / reads,
                       warp_inst_per_read=k0["smsp__inst_executed.sum"] / reads)
    # per-source-line instruction shares
    src = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    agg, cur, cur_file, ie, isamp = collections.OrderedDict(), None, None, None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
        elif r[0] not in ("", "Function Name") and len(r) > 2 and r[2] == "-":
            cur = (cur_file, int(r[0]), r[1].strip())
            agg.setdefault(cur, [0, 0, 0])
        elif r[0] == "" and cur is not None:
            try:
                agg[cur][0] += int(r[ie]); agg[cur][1] += int(r[isamp]); agg[cur][2] += 1
            except (ValueError, IndexError):
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    lines = ["# %s" % rep, "# kernel: %s" % k0["kernel"][:100], ""]
    for key in KEYS:
        if key in k0:
            lines.append("%-70s %s %s" % (key, k0[key], k0["_units"][key]))
    for key in sorted(k0):
        if key.startswith("smsp__average_warps_issue_stalled"):
            lines.append("%-70s %.3f" % (key, k0[key]))
    if reads:
        lines += ["", "reads/launch %d  dram bytes/read %.1f  warp-instructions/read %.1f" % (reads, summary["dram_bytes_per_read"], summary["warp_inst_per_read"])]
    lines += ["", "SASS instructions in kernel: %d" % sum(v[2] for v in agg.values()), "", "top source lines by executed warp-instructions:"]
    for (f, l, s), (n, sm, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        lines.append("%5.1f%% inst %5.1f%% samples %4d sass  %s:%d  %s" % (100 * n / tot, 100 * sm / tots, c, f, l, s[:100]))
    open(outbase + ".txt", "w").write("\n".join(lines) + "\n")
    for k in summary["kernels"]:
        k.pop("_units", None)
    json.dump(summary, open(outbase + ".json", "w"), indent=1)
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
