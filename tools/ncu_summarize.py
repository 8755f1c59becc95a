"""Turn an .ncu-rep (ncu --set full) into the small CSV kept under profiles/.

    python tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_X.csv "comment line" ...

Runs `ncu -i REP --page raw --csv` (works without a GPU) and keeps the columns the roofline discussion needs."""
import csv
import io
import subprocess
import sys

KEEP = [
    "ID", "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    comments = sys.argv[3:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(k) for k in KEEP if k in head]
    with open(out, "w", newline="") as f:
        for c in comments:
            f.write("# " + c + "\n")
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(body)} launches, {len(idx)} columns")


if __name__ == "__main__":
    main()
