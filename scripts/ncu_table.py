"""Markdown table of the key `ncu --set full` metrics of every kernel in a report:  python scripts/ncu_table.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
MODELS = {0: "pinhole", 1: "rad_tan", 2: "kannala_brandt", 3: "ucm", 4: "eucm", 5: "double_sphere", 6: "fov"}


def f(r, k, scale=1.0, fmt="{:.1f}"):
    if k not in idx or r[idx[k]] in ("", "n/a"):
        return "-"
    return fmt.format(float(r[idx[k]].replace(",", "")) * scale)


def unit(k):
    return rows[1][idx[k]] if k in idx else ""


print("| kernel | model | duration us | DRAM % of ncu peak | FP64 pipe % | issue active % | regs | warps active % | DRAM read GB | DRAM write GB | long_scoreboard / issue | top other stall |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = name.split("<")[0].replace("void ", "")
    mid = name.split("<")[1].split(",")[0].split(">")[0].strip() if "<" in name else ""
    model = MODELS.get(int(mid), mid) if mid.isdigit() else mid
    dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    du = unit("gpu__time_duration.sum")
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
    def gb(k):
        if k not in idx or r[idx[k]] in ("", "n/a"): return "-"
        v = float(r[idx[k]].replace(",", "")); u = unit(k)
        return "{:.2f}".format(v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9))
    stalls = [(h.split("stalled_")[1].replace("_per_issue_active.ratio", ""), float(r[idx[h]].replace(",", ""))) for h in hdr
              if "issue_stalled" in h and "per_issue_active" in h and r[idx[h]] not in ("", "n/a")]
    stalls = [s for s in stalls if s[0] not in ("selected",)]
    ls = dict(stalls).get("long_scoreboard", 0.0)
    other = max((s for s in stalls if s[0] != "long_scoreboard"), key=lambda s: s[1], default=("-", 0.0))
    print(f"| `{short}` | {model} | {dur_us:.0f} | {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{f(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')} | {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{f(r, 'launch__registers_per_thread', fmt='{:.0f}')} | {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | "
          f"{gb('dram__bytes_read.sum')} | {gb('dram__bytes_write.sum')} | {ls:.2f} | {other[0]} {other[1]:.2f} |")
