"""Turns an `ncu --set full --import-source on` report into the small CSV summaries kept under profiles/:
    python scripts/summarize_ncu.py gpurun_out/r1_fwd.ncu-rep profiles/r1_fwd "command line that was profiled"
writes <prefix>_ncu_summary.csv (per-kernel metrics) and <prefix>_stall_summary.csv (warp-stall samples by reason and
the ten hottest SASS instructions).  Needs the `ncu` CLI only (no GPU)."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct",
]


def page(report, name):
    out = subprocess.run(["ncu", "-i", report, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report, prefix, command = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = page(report, "raw")
    header, units = raw[0], raw[1]
    with open(prefix + "_ncu_summary.csv", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; report {report.split('/')[-1]}\n")
        f.write(f"# command: {command}\n")
        f.write("kernel,metric,value,unit\n")
        for row in raw[2:]:
            d = dict(zip(header, row))
            u = dict(zip(header, units))
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            for m in METRICS:
                if m in d:
                    f.write(f"{name},{m},{d[m]},{u[m]}\n")
    src = page(report, "source")
    sections, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1].split("(")[0].replace("void ", ""), "hdr": None, "rows": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    with open(prefix + "_stall_summary.csv", "w") as f:
        f.write(f"# warp-stall samples from the source page of {report.split('/')[-1]} (SASS view)\n")
        seen = set()
        for s in sections:
            if s["name"] in seen or not s["rows"]:
                continue
            seen.add(s["name"])
            ix = {n: i for i, n in enumerate(s["hdr"])}
            cols = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
            total = sum(int(r[ix["# Samples"]] or 0) for r in s["rows"])
            f.write(f"kernel,{s['name']}\ntotal samples,{total}\nstall reason,samples\n")
            for n, c in sorted(((sum(int(r[ix[c]] or 0) for r in s["rows"]), c) for c in cols), reverse=True):
                if n:
                    f.write(f"{c},{n}\n")
            f.write("hottest instructions: samples,share,SASS,dominant stall\n")
            for r in sorted(s["rows"], key=lambda r: -int(r[ix["# Samples"]] or 0))[:10]:
                n = int(r[ix["# Samples"]] or 0)
                top = max(cols, key=lambda c: int(r[ix[c]] or 0))
                f.write(f"{n},{100 * n / max(total, 1):.1f}%,{' '.join(r[ix['Source']].split())[:70]},{top}\n")


if __name__ == "__main__":
    main()
