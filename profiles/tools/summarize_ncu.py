"""Summarise an Nsight Compute report (read here with `ncu -i ... --page raw --csv`) into the text / JSON
files committed under profiles/.  Usage: python profiles/tools/summarize_ncu.py REPORT.ncu-rep OUT_PREFIX"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(rep, prefix):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    txt, js = [], {}
    for r in data:
        name = r[ki]
        txt.append("===== " + name[:110])
        rec = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                txt.append("  %-70s %s %s" % (k, r[i], units[i]))
                rec[k] = [r[i], units[i]]
        rd, wr = rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum")
        if rd and wr:
            tot = float(rd[0].replace(",", "")) * UNIT[rd[1]] + float(wr[0].replace(",", "")) * UNIT[wr[1]]
            rec["dram_bytes_per_launch"] = tot
            txt.append("  %-70s %.0f byte" % ("dram bytes (read + write) per launch", tot))
        js.setdefault(name.split("(")[0].replace("void ", "").strip(), rec)
    open(prefix + "_summary.txt", "w").write("\n".join(txt) + "\n")
    json.dump(js, open(prefix + "_summary.json", "w"), indent=1)
    open(prefix + "_raw.csv", "w").write(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
