"""Turns the raw ncu exports under gpurun_out/ into the tracked summaries under profiles/:

  profiles/r01_launches.csv         per-launch gpu__time_duration of one `bench.py` run (as exported by ncu)
  profiles/r01_launch_shares.md     kernel shares of a train step from that launch list
  profiles/r01_kernel_metrics.json  per-kernel metrics of the `--set full` captures (dram bytes per launch, pipe
                                    utilisation, ...); bench.py reads `dram_bytes` from here for roofline.traffic
  profiles/r01_kernel_metrics.md    the same, readable

Usage: python scripts/summarize_profiles.py [--launches gpurun_out/r01_launches_v2.csv] [--raw a.csv b.csv ...]
(raw = `ncu -i X.ncu-rep --page raw --csv`)."""
import argparse, collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "r01_launches_final.csv"))
ap.add_argument("--reps", nargs="*", default=[])
ap.add_argument("--tag", default="r01")
args = ap.parse_args()
prof = os.path.join(ROOT, "profiles")
os.makedirs(prof, exist_ok=True)

def short(name):
    n = name.split("(")[0].replace("void ", "").replace("gte::", "")
    return n

# ---- launch list -------------------------------------------------------------------------------
if os.path.exists(args.launches):
    rows = [r for r in csv.reader(open(args.launches)) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    shutil.copy(args.launches, os.path.join(prof, f"{args.tag}_launches.csv"))
    ours = [r for r in data]
    # the last replayed step: find the last k_adam launch and the one before it
    idx = [i for i, r in enumerate(ours) if "k_adam" in r[ki]]
    lines = [f"# {args.tag}: kernel shares of one train step (ncu --metrics gpu__time_duration.sum --clock-control none)\n",
             f"Source: `profiles/{args.tag}_launches.csv` ({len(data)} launches of `bench.py`; per-launch times under ncu are serialised and",
             "cold-cache, so only the SHARES are comparable with the event-timed numbers in the bench line).\n"]
    if len(idx) >= 2:
        step = ours[idx[-2] + 1: idx[-1] + 1]
        agg = collections.OrderedDict()
        for r in step:
            a = agg.setdefault(short(r[ki]), [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi]) / 1e3
        tot = sum(a[1] for a in agg.values())
        lines.append(f"One step = {len(step)} launches, {tot:.1f} us of kernel time under ncu.\n")
        lines.append("| kernel | launches | us | share |\n|---|---:|---:|---:|")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |")
    open(os.path.join(prof, f"{args.tag}_launch_shares.md"), "w").write("\n".join(lines) + "\n")
    print("wrote launch shares")

# ---- full captures -----------------------------------------------------------------------------------
WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_insts",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}
metrics = {}
for rep in args.reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 20]
    if len(rows) < 3:
        print("no data in", rep); continue
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        rec = {"capture": os.path.basename(rep)}
        for m, k in WANT.items():
            if m in hdr and r[hdr.index(m)] != "":
                try:
                    v = float(r[hdr.index(m)].replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(m)]
                if k in ("dram_read", "dram_write"): v *= UNIT.get(u, 1.0)
                if k == "time_us": v *= UNIT.get(u, 1.0)
                rec[k] = v
        rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
        name = short(r[ki])
        key = f"{name} grid={int(rec.get('grid', 0))}"
        metrics.setdefault(key, []).append(rec)
if metrics:
    out = {}
    md = [f"# {args.tag}: `ncu --set full --clock-control none` captures (per launch)\n",
          "| kernel | capture | us | DRAM read MB | DRAM write MB | L2 % | LSU wavefront % | issue % | tensor pipe % | regs |",
          "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for key, recs in metrics.items():
        out[key] = recs
        for rec in recs:
            md.append(f"| `{key}` | {rec['capture']} | {rec.get('time_us', 0):.1f} | {rec.get('dram_read', 0) / 1e6:.1f} | "
                      f"{rec.get('dram_write', 0) / 1e6:.1f} | {rec.get('l2_pct', 0):.1f} | "
                      f"{rec.get('lsu_wavefronts_pct', 0):.1f} | {rec.get('issue_active_pct', 0):.1f} | "
                      f"{rec.get('tensor_pipe_pct', 0):.1f} | {int(rec.get('regs', 0))} |")
    json.dump(out, open(os.path.join(prof, f"{args.tag}_kernel_metrics.json"), "w"), indent=1)
    open(os.path.join(prof, f"{args.tag}_kernel_metrics.md"), "w").write("\n".join(md) + "\n")
    print("wrote kernel metrics for", len(out), "kernels")
