"""Extracts the judged metrics of one `ncu --set full` capture into a small markdown + json summary."""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def main(rep, out_md, out_json, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                d[w] = (r[hdr.index(w)], units[hdr.index(w)])
        out.append(d)
    with open(out_md, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (one launch, read with "
                "`ncu -i ... --page raw --csv`).\n\n" % title)
        for d in out:
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % d["kernel"][:90])
            for w in WANT:
                if w in d:
                    f.write("| %s | %s | %s |\n" % (w, d[w][0], d[w][1]))
            f.write("\n")

    def num(d, k):
        v, u = d[k]
        v = float(v.replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v * scale
    d = out[-1]
    js = {"kernel": d["kernel"], "dram_bytes_per_launch": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"),
          "dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum"),
          "source": rep}
    with open(out_json, "w") as f:
        json.dump(js, f, indent=1)
    print(open(out_md).read())


if __name__ == "__main__":
    main(*sys.argv[1:5])
