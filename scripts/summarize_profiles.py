"""Turns the ncu CSVs of scripts/profile_step.py into the tracked summaries under profiles/.
  python scripts/summarize_profiles.py <tag> <launches.csv> <full_raw.csv>
writes profiles/<tag>_step_shares.txt, profiles/<tag>_ncu_full_summary.csv, profiles/<tag>_traffic.json"""
import csv, json, os, re, sys
from collections import OrderedDict

tag, launches, full = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"void |lpgnn::|\(anonymous namespace\)::|unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("lpgnn::", "")


# ---- launch list (one metric per row)
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    k = short(r[4])
    t = float(r[-1]) / 1e3
    a = agg.setdefault(k, [0.0, 0])
    a[0] += t; a[1] += 1
tot = sum(a[0] for a in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_step_shares.txt"), "w") as f:
    what = sys.argv[4] if len(sys.argv) > 4 else "bf16 prediction step (lpgnn_predict_basis)"
    f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of ONE C2 {what} = {len(rows)} launches\n")
    f.write("# (scripts/profile_step.py inside cudaProfilerStart/Stop).  Cold-cache, serialised: compare SHARES, not absolutes.\n\n")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{t:9.1f} us  {100 * t / tot:5.1f} %  x{c:<2d} {k}\n")
    f.write(f"{tot:9.1f} us  total\n")

if full == "-":
    print(open(os.path.join(ROOT, "profiles", f"{tag}_step_shares.txt")).read())
    sys.exit(0)

# ---- full capture (--page raw: one row per launch, one column per metric)
raw = list(csv.reader(open(full)))
hi = [i for i, r in enumerate(raw) if r and r[0] == "ID"][0]
hdr, units, data = raw[hi], raw[hi + 1], raw[hi + 2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "gpc__cycles_elapsed.avg.per_second"]
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for d in data:
        w.writerow([short(d[i]) if hdr[i] == "Kernel Name" else d[i] for i in idx])

col = {h: i for i, h in enumerate(hdr)}
def to_bytes(d, k):
    v, u = float(d[col[k]]), units[col[k]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
traffic = OrderedDict()
for d in data:
    k = short(d[col["Kernel Name"]])
    traffic.setdefault(k, []).append(int(to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum")))
json.dump({"_comment": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), ncu --set full --clock-control none, "
                       "one C2 step (scripts/profile_step.py; precision in the file name); launches in step order", "per_kernel": traffic},
          open(os.path.join(ROOT, "profiles", f"{tag}_traffic_raw.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_step_shares.txt")).read())
for k, v in traffic.items():
    print(k, [round(x / 1e6, 1) for x in v])
