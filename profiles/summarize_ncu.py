#!/usr/bin/env python
"""Turn .ncu-rep captures (gpurun_out/, scratch) into the committed text summaries and roofline_traffic.json.

    python profiles/summarize_ncu.py <tag> <K> <rep> [<rep> ...]       # e.g. r01e 128 gpurun_out/prof_r01e_*.ncu-rep

Per kernel launch in the reports: the metrics the roofline block of bench.py and DESIGN.md quote (duration, DRAM
bytes, tensor-pipe activity, shared-memory wavefronts, registers, grid) -> profiles/<tag>_ncu_full_<rep>.txt, and the
per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, last captured launch of every kernel)
-> profiles/roofline_traffic.json["K<K>"][<kernel>], which bench.py copies into roofline.traffic.
Runs here (no GPU needed): `ncu -i` only reads the report."""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__bytes_(read|write)\.sum\.per_second|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_bytes\.sum|lts__t_sectors_op_read\.sum|lts__t_sectors_op_write\.sum|"
    r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_tensor.*|sm__cycles_active\.avg|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_.*|waves_per_multiprocessor)|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|sass__inst_executed_shared_(loads|stores)|"
    r"smsp__cycles_active\.avg|sm__inst_executed\.sum)$")


def main():
    tag, K, reps = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    tpath = os.path.join(HERE, "roofline_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        name = os.path.basename(rep).replace(".ncu-rep", "")
        lines = ["# %s: ncu --set full --clock-control none --import-source on (tests/ncu_target.py), B200; one block per captured launch" % name,
                 "# (a capture replays every launch ~40 times with serialised, cold-ish caches: compare shares and traffic, not absolutes)", ""]
        for r in data:
            kname = r[col["Kernel Name"]]
            short = re.sub(r"<.*", "", re.sub(r"\(.*", "", kname).split("::")[-1])
            lines.append("==== %s   grid %s  block %s" % (kname, r[col["Grid Size"]] if "Grid Size" in col else "?",
                                                            r[col["Block Size"]] if "Block Size" in col else "?"))
            for h in hdr:
                if KEEP.match(h) and r[col[h]] != "":
                    lines.append("%-95s %s %s" % (h, r[col[h]], units[col[h]]))
            lines.append("")

            def num(metric):
                v, u = float(r[col[metric]].replace(",", "")), units[col[metric]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            try:
                traffic.setdefault("K%d" % K, {})[short] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
                traffic.setdefault("tensor_pipe_active_pct_K%d" % K, {})[short] = round(
                    float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]), 2)
            except (KeyError, ValueError):
                pass
        with open(os.path.join(HERE, "%s_ncu_full_%s.txt" % (tag, name.replace("prof_%s_" % tag, ""))), "w") as f:
            f.write("\n".join(lines))
    traffic["_source"] = "profiles/summarize_ncu.py over the %s captures; bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum" % tag
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
