"""Turn gpurun_out/*.ncu-rep + the launch list into the committed summaries under profiles/."""
import csv, io, json, os, subprocess, sys, collections
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = "profiles"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
lines = ["# ncu summaries, round %s" % R, "",
         "Command profiled: `python bench.py --steps 2 --warmup 3 --quick` (24-qubit H12-chain UCCSD energy+gradient), one B200,",
         "`ncu --set full --clock-control none --import-source on`; numbers under ncu are cold-cache and serialised:",
         "use them for traffic, occupancy, stall reasons and SHARES of the step, never as bench values.", ""]
def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]
traffic = {}
for name in ("tile_fwd", "tile_adj", "happly", "dense_fwd", "sp2f", "sp2a", "sp2", "csweep_fwd", "csweep_adj", "sigma", "rdm_gram"):
    rep = "gpurun_out/%s_%s.ncu-rep" % (R, name)
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
    with open(os.path.join(OUT, "%s_%s_metrics.csv" % (R, name)), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEYS + stall)
        for r in rows:
            w.writerow([r[hdr.index("Kernel Name")][:60]] + [r[hdr.index(k)] if k in hdr else "" for k in KEYS + stall])
    lines.append("## %s (`%s`)" % (name, os.path.basename(rep)))
    for r in rows:
        kn = r[hdr.index("Kernel Name")].split("(")[0]
        g = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
        u = lambda k: units[hdr.index(k)] if k in hdr else ""
        lines.append("")
        lines.append("* kernel `%s`, grid %s x %s threads, %s regs, %s %s dynamic smem, occupancy limits: smem %s / regs %s CTAs per SM" % (
            kn, g("launch__grid_size"), g("launch__block_size"), g("launch__registers_per_thread"),
            g("launch__shared_mem_per_block_dynamic"), u("launch__shared_mem_per_block_dynamic"),
            g("launch__occupancy_limit_shared_mem"), g("launch__occupancy_limit_registers")))
        lines.append("  * duration %s %s; DRAM read %s %s + write %s %s (DRAM throughput %s%% of peak); L2 hit %s%%" % (
            g("gpu__time_duration.sum"), u("gpu__time_duration.sum"), g("dram__bytes_read.sum"), u("dram__bytes_read.sum"),
            g("dram__bytes_write.sum"), u("dram__bytes_write.sum"),
            g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("lts__t_sector_hit_rate.pct")))
        lines.append("  * SM throughput %s%%, issue active %s%%, warps active %s%% of peak, %s warp instructions, %s threads/instr" % (
            g("sm__throughput.avg.pct_of_peak_sustained_elapsed"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("smsp__inst_executed.sum"),
            g("smsp__thread_inst_executed_per_inst_executed.ratio")))
        try:
            def _b(k):
                v = float(g(k).replace(",", "")); un = u(k).lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(un, 1)
            key = kn.replace("void ", "").split("<")[0] + ("<true>" if ("<1" in kn or "(bool)1" in kn) else ("<false>" if ("<0" in kn or "(bool)0" in kn) else ""))
            t = traffic.setdefault(key, [])
            t.append(_b("dram__bytes_read.sum") + _b("dram__bytes_write.sum"))
        except Exception:
            pass
        top = sorted(((float(r[hdr.index(s)] or 0), s) for s in stall), reverse=True)[:4]
        lines.append("  * top stalls (warps per issue): " + ", ".join(
            "%s %.2f" % (s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, s in top))
    lines.append("")
# launch list
ll = "gpurun_out/launches_%s.csv" % R
if os.path.exists(ll):
    txt = open(ll).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if unit == "ns": v /= 1e3
        elif unit == "ms": v *= 1e3
        elif unit == "s": v *= 1e6
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines += ["## launch list (`launches_%s.csv`: %d launches = 2 timed steps)" % (R, len(rows)), "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
    lines.append("")
    with open(os.path.join(OUT, "launches_%s.csv" % R), "w") as f:
        f.write(txt[start:])
if traffic:
    tj = os.path.join(OUT, "traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    for k, v in traffic.items():
        cur[k] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v),
                  "source": "ncu --set full, %s (cold cache, serialised)" % R}
    json.dump(cur, open(tj, "w"), indent=1, sort_keys=True)
if len(sys.argv) > 2:
    lines += ["", open(sys.argv[2]).read()]
open(os.path.join(OUT, "%s_summary.md" % R), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
