"""DRAM traffic per launch of the headline kernel from an `ncu --set full` report -> profiles/traffic.json (read by bench.py
for `roofline.traffic`):   python scripts/ncu_traffic.py gpurun_out/r02_lin.ncu-rep
Also copies the raw metric rows of every lin_kernel launch next to it (profiles/r02_ncu_lin_raw.csv) so the number can be re-derived."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
keep = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
keep = [k for k in keep if k in idx]
out_rows = [keep, [units[idx[k]] for k in keep]]
best = None
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if "lin_kernel" not in name:
        continue
    out_rows.append([r[idx[k]] for k in keep])
    if "lin_kernel<5, 0" in name or "(int)5, (int)0" in name:   # Double Sphere, pixel residual: the bench headline
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * SCALE[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * SCALE[units[idx["dram__bytes_write.sum"]]]
        best = {"linearize_ds_pixel_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "kernel": name,
                "source": os.path.basename(rep) + " (ncu --set full --clock-control none, 100 M points, one launch)"}
tag = os.path.basename(rep).split("_")[0]
with open(os.path.join(root, "profiles", f"{tag}_ncu_lin_raw.csv"), "w", newline="") as f:
    csv.writer(f).writerows(out_rows)
if best:
    json.dump(best, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(best))
