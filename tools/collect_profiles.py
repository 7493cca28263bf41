"""Copy what a tools/gpu_final.sh visit left in gpurun_out/ into profiles/ under a round/version tag:
bench lines, the ncu launch list, text summaries of the two full ncu captures, DRAM traffic of the deflate launch."""
import csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = "gpurun_out", "profiles"
def cp(src, dst):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(P, dst)); print("copied", dst)
cp("bench.json", f"{tag}_bench.json"); cp("bench_ref.json", f"{tag}_bench_ref.json"); cp("launches.csv", f"{tag}_launches_bench.csv")
cp("extra.json", f"{tag}_extra.json"); cp("pcie_duplex.json", f"{tag}_pcie_duplex.json")
cp("pytest_gpu.log", f"{tag}_pytest_gpu.log"); cp("phases_window.json", f"{tag}_phases_window_kernel.json")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
def summarize(rep, out, title):
    path = os.path.join(G, rep)
    if not os.path.exists(path): return None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3: return None
    h, u, v = rows[0], rows[1], rows[2]
    d = {k: (v[i], u[i]) for i, k in enumerate(h)}
    with open(os.path.join(P, out), "w") as f:
        f.write(title + "\n" + d.get("Kernel Name", ("", ""))[0] + "\n\n")
        for k in KEYS:
            if k in d: f.write(f"{k:90s} {d[k][1]:16s} {d[k][0]}\n")
        for k in sorted(d):
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"): f.write(f"{k:90s} {d[k][1]:16s} {d[k][0]}\n")
    print("wrote", out)
    return d
d = summarize("prof_deflate.ncu-rep", f"{tag}_ncu_deflate_summary.txt",
              "ncu --set full --clock-control none --import-source on; one launch over a 512 MiB device-resident call of `bench.py --gib 0.5` (8192 windows of 64 KiB)")
summarize("prof_lz4.ncu-rep", f"{tag}_ncu_lz4_summary.txt",
          "ncu --set full --clock-control none --import-source on; one launch of the LZ4 window kernel over 512 MiB (tools/gpu_perf_extra.py, EXTRA_ONLY=lz4)")
summarize("prof_inflate.ncu-rep", f"{tag}_ncu_inflate_summary.txt",
          "ncu --set full --clock-control none --import-source on; one launch of tools/gpu_inflate_bench.py (8192 gzip-ext members of 64 KiB made by our compressor, 512 MiB out)")
if d:
    def num(k):
        val, unit = d[k]; x = float(val.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    b = json.load(open(os.path.join(G, "bench.json"))) if os.path.exists(os.path.join(G, "bench.json")) else {}
    sys.path.insert(0, os.getcwd())
    import bench
    json.dump({"deflate_dram_bytes_per_launch": int(rd + wr), "kernel": d.get("Kernel Name", ("", ""))[0], "dram_read": int(rd), "dram_write": int(wr),
               "launch": "512 MiB device-resident call, 8192 windows of 64 KiB", "capture": f"profiles/{tag}_ncu_deflate_summary.txt", "source_sha": bench.kernel_source_sha(),
               "source": f"ncu --set full capture summarised in profiles/{tag}_ncu_deflate_summary.txt",
               "algorithmic_bytes": b.get("roofline", {}).get("bytes_per_launch")}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("traffic", int(rd + wr))
