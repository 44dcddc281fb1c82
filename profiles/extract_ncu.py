#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summaries kept under profiles/.

  python profiles/extract_ncu.py gpurun_out/prof.ncu-rep profiles/out.json "how the capture was made"

Per launch: duration, warp instructions, issue-active %, achieved occupancy, registers, DRAM bytes, shared-memory
wavefronts / bank conflicts, DRAM / SM / FMA / tensor pipe utilisation, L2 hit rate and the six largest warp-stall
reasons per issued instruction.  Metric names are ncu's own (columns of `--page raw --csv`, section prefix dropped).
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_tensor.sum"]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio$|"
                   r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio$")


def main(rep, out, source):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    short = [h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] in ("TriageCompute",) else h for h in hdr]
    kernels = {}
    for r in data:
        if len(r) != len(hdr):
            continue
        rec = dict(zip(short, r))
        name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "").strip()
        grid = rec["Grid Size"].strip("()").split(",")[0].strip()
        key = "%s@grid%s" % (name, grid)
        k = 2
        while key in kernels:
            key = "%s@grid%s#%d" % (name, grid, k)
            k += 1
        entry, stalls = {}, {}
        for h, u, v in zip(short, units, r):
            if h in KEEP and v not in ("", "no data"):
                entry[h] = float(v.replace(",", "")) if re.match(r"^-?[\d.,]+$", v) else v
                if u:
                    entry.setdefault("_units", {})[h] = u
            m = STALL.search(h)
            if m and v not in ("", "no data"):
                stalls[m.group(1) or m.group(2)] = float(v.replace(",", ""))
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        entry["top_stalls_per_issue"] = {k2: round(v2, 4) for k2, v2 in top}
        kernels[key] = entry
    json.dump({"source": source, "kernels": kernels}, open(out, "w"), indent=1)
    print("wrote %s: %d launches" % (out, len(kernels)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
