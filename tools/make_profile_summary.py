"""Turn ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profile_summary.py <tag> <launches.csv> <kernel-report.ncu-rep> [more reports ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
with open(os.path.join(out, tag + "_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), launches], stdout=subprocess.PIPE, text=True).stdout)
traffic = {}
import re
for rep in reps:
    base = os.path.splitext(os.path.basename(rep))[0]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    seen = set()
    for row in rows[2:]:
        d = dict(zip(rows[0], zip(rows[1], row)))
        kname = d.get("Kernel Name", ("", "?"))[1]
        short = re.sub(r"[^A-Za-z0-9_]+", "_", kname.split("(")[0].replace("void ", "").replace("<unnamed>::", "")).strip("_")
        if short in seen or float(d["gpu__time_duration.sum"][1].replace(",", "")) < 0.02:      # one (the first) launch per kernel; skip trivial ones
            continue
        seen.add(short)
        with open(os.path.join(out, tag + "_" + short + ".txt"), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on (%s), kernel: %s\n" % (os.path.basename(rep), kname))
            for k in KEYS:
                if k in d:
                    f.write("%-72s %s %s\n" % (k, d[k][1], d[k][0]))
            st = {k: float(v[1].replace(",", "")) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v[1]}
            tot = sum(st.values()) or 1
            f.write("\n# warp stall reasons (share of samples)\n")
            for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
                f.write("%-40s %5.1f%%\n" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
            f.write("\n# hottest CUDA source lines\n")
            fn = kname.split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0].split("::")[-1]
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "30", fn], stdout=subprocess.PIPE, text=True).stdout)
        def num(k):
            u, v = d[k]; v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[short] = {"kernel": kname, "bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "ms": float(d["gpu__time_duration.sum"][1].replace(",", ""))}
print(json.dumps(traffic, indent=1))
