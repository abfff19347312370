"""Turn gpurun_out/final (tools/run_final.sh) into the tracked summaries under profiles/ (prefix rNN_final_; the round
is argv[1], default 2)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUND = int(sys.argv[1]) if len(sys.argv) > 1 else 2
SRC, DST, PRE = os.path.join(ROOT, "gpurun_out", "final"), os.path.join(ROOT, "profiles"), f"r{ROUND:02d}_final_"


def copy(src, dst):
    shutil.copyfile(os.path.join(SRC, src), os.path.join(DST, PRE + dst))


def short_name(k):
    return re.sub(r"^void |\(.*$|mvldm::|<unnamed>::|unnamed>::|at::native::.*?::", "", k)[:46]


def read_launches(name):
    """ncu --csv raw rows (one per launch x metric) -> list of {name, us, dram_bytes} in launch order"""
    rows = [l for l in open(os.path.join(SRC, name)) if l.startswith('"')]
    rd = list(csv.DictReader(io.StringIO("".join(rows))))
    out, by_id = [], {}
    for r in rd:
        e = by_id.get(r["ID"])
        if e is None:
            e = by_id[r["ID"]] = {"name": r["Kernel Name"], "us": 0.0, "dram_bytes": 0.0}
            out.append(e)
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        elif r["Metric Name"].startswith("dram__bytes"):
            e["dram_bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return out


def one_step(launches):
    # one whole DDIM step = build_inputs_kernel ... ddim_step_kernel; take the last complete one (the launches after it
    # belong to bench.py's per-op profiling pass)
    ends = [i for i, r in enumerate(launches) if "ddim_step_kernel" in r["name"]]
    begins = [i for i, r in enumerate(launches) if "build_inputs_kernel" in r["name"] and i < ends[-1]]
    return launches[begins[-1]:ends[-1] + 1]


def launch_summary(name, header):
    step = one_step(read_launches(name))
    agg = collections.OrderedDict()
    for r in step:
        a = agg.setdefault(short_name(r["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["us"]
        a[2] += r["dram_bytes"]
    tot = sum(a[1] for a in agg.values())
    out = [header, ""]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:46s} launches {a[0]:4d} {a[1]:9.1f} us {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0]:6.1f} us"
                   f"  dram {a[2] / 1e6:9.1f} MB ({a[2] / max(a[0], 1) / 1e6:7.2f} MB/launch)")
    out.append(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches, "
               f"dram {sum(a[2] for a in agg.values()) / 1e6:.1f} MB")
    return "\n".join(out) + "\n"


def traffic_json(name, source):
    """per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of one DDIM step: bench.py's
    roofline.traffic reads this file"""
    step = one_step(read_launches(name))
    agg = {}
    for r in step:
        k = short_name(r["name"]).split("<")[0]
        a = agg.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += r["dram_bytes"]
        a["us"] += r["us"]
    for a in agg.values():
        a["dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]
    agg["source"] = source
    return agg


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu_summary(rep, header):
    p = subprocess.run(["ncu", "-i", os.path.join(SRC, rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    r = list(csv.reader(io.StringIO(p.stdout)))
    h, units, row = r[0], r[1], r[2]
    out = [header, "", f"Kernel Name = {row[h.index('Kernel Name')]}"]
    for w in WANT:
        if w in h:
            out.append(f"{w} = {row[h.index(w)]} {units[h.index(w)]}")
    for i, k in enumerate(h):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(row[i]) >= 0.3:
                    out.append(f"{k} = {row[i]} {units[i]}")
            except ValueError:
                pass
    return "\n".join(out) + "\n"


def main():
    for a, b in [("bench_n1.json", "bench_n1.json"), ("bench_cfg.json", "bench_cfg.json"),
                 ("bench_cfg_two_forwards.json", "bench_cfg_two_forwards.json"), ("bench_variant_b.json", "bench_variant_b.json"),
                 ("bench_8scenes.json", "bench_8scenes_per_gpu.json"), ("bench_reference.json", "bench_reference_cpu.json"),
                 ("bench_standard.json", "bench_standard_transformer.json"), ("bench_vae.json", "bench_vae.json"),
                 ("clocks.csv", "clocks.csv"), ("scale_check.txt", "scale_check.txt"), ("prof_ops.txt", "ops_and_timelines.txt"),
                 ("gn_graph_bench.txt", "groupnorm_in_graph.txt"), ("gemm_sweep.txt", "gemm_config_sweep.txt"),
                 ("excess_v8.txt", "per_op_vs_floor.txt"), ("micro.txt", "micro_mufu_pdl.txt"), ("launches_cold.csv", "launches_cold.csv"),
                 ("launches_warm.csv", "launches_warm.csv"), ("pytest_gpu.log", "pytest_gpu.log"), ("smoke.log", "smoke.log")]:
        if os.path.exists(os.path.join(SRC, a)):
            copy(a, b)
    cmd = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none {} "
           "python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-config4 --view-sharded-views 0")
    b = json.loads(open(os.path.join(SRC, "bench_n1.json")).read().strip().splitlines()[-1])
    hdr = (f"round {ROUND} final, one DDIM step = 1 scene x 8 views, no CFG, the timed step of\n`" + cmd + "`\n"
           f"bench.py (CUDA-graph replay, CUDA events, not under ncu): {b['ms_per_step']:.2f} ms/step = {b['value']:.1f} steps/s. "
           "ncu times are serialised{}: compare shares.  dram = dram__bytes_read.sum + dram__bytes_write.sum.\n")
    if os.path.exists(os.path.join(SRC, "launches_warm.csv")):
        open(os.path.join(DST, PRE + "launches_warm.summary.txt"), "w").write(
            launch_summary("launches_warm.csv", hdr.format("--cache-control none", " (caches kept warm)")))
        json.dump(traffic_json("launches_warm.csv", f"profiles/{PRE}launches_warm.csv: `" + cmd.format("--cache-control none") + "`, "
                               "the last whole DDIM step (caches not flushed between launches, as in the real step)"),
                  open(os.path.join(DST, PRE + "traffic.json"), "w"), indent=1)
    if os.path.exists(os.path.join(SRC, "launches_cold.csv")):
        open(os.path.join(DST, PRE + "launches_cold.summary.txt"), "w").write(
            launch_summary("launches_cold.csv", hdr.format("", " and cold-cache")))
    for rep, dst, what in [
            ("ncu_gemm_conv_l0.ncu-rep", "ncu_gemm_conv_l0.txt", "ncu --set full --clock-control none --import-source on -k regex:gemm_tc, "
             "tools/prof_gemm.py 8 320 320 32 (level-0 conv3x3 320->320 at 8 views: M=8192 N=320 K=2880, 128 tiles of 128x160)"),
            ("ncu_attn_l0.ncu-rep", "ncu_attn_l0.txt", "ncu --set full --clock-control none --import-source on -k regex:attn64q, "
             "tools/prof_attn.py 1 8192 40 (joint attention of one scene: 8 views x 1024 tokens, 8 heads, d=40 padded to 64)"),
            ("ncu_gn_flat_l0.ncu-rep", "ncu_groupnorm_l0.txt", "ncu --set full --clock-control none --import-source on -k regex:gn_flat, "
             "tools/prof_gn.py 8 1024 320 0 (GroupNorm+SiLU of one level-0 tensor: 8 images x 1024 px x 320 ch)")]:
        if os.path.exists(os.path.join(SRC, rep)):
            open(os.path.join(DST, PRE + dst), "w").write(ncu_summary(rep, what))


if __name__ == "__main__":
    main()
