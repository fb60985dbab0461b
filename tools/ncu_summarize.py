"""Turn the ncu outputs of tools/gpu_round.sh into the small, committed summaries under profiles/:
  launches CSV (gpu__time_duration per launch)  ->  per-kernel share table (JSON)
  --set full report (.ncu-rep)                  ->  per-kernel DRAM traffic / throughput / occupancy (JSON)
usage: python tools/ncu_summarize.py TAG [PREFIX] (reads gpurun_out/*_TAG.*, writes profiles/PREFIX_ncu_summary_TAG.json; PREFIX default r2)"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else "r2"
out = {}

lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
        k = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        agg[k] += v
        cnt[k] += 1
    tot = sum(agg.values())
    out["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 1100 python bench.py "
                                     "--steps 2 --warmup 3 --no-cpu-baseline --no-graph",
                          "launches": sum(cnt.values()), "total_us": tot,
                          "kernels": [{"kernel": k, "launches": cnt[k], "us": round(v, 1), "share": round(v / tot, 4)}
                                      for k, v in agg.most_common()]}

rp = os.path.join(ROOT, "gpurun_out", f"ncu_top_{tag}.ncu-rep")
rc = os.path.join(ROOT, "gpurun_out", f"ncu_top_{tag}.raw.csv")    # exported on the GPU box when the report is too big to copy
if os.path.exists(rp) or os.path.exists(rc):
    txt = (open(rc).read() if os.path.exists(rc) else
           subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
            "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
            "smsp__inst_executed.sum": "warp_instructions"}
    idx = {k: hdr.index(k) for k in want if k in hdr}
    caps = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")}
        for k, name in want.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    e[name] = float(r[i].replace(",", ""))
                    e[name + "_unit"] = units[i]
                except ValueError:
                    pass
        caps.append(e)
    out["full_captures"] = {"command": "ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05|dwconv_|attn_|ln_ "
                                       "python tools/kernel_bench.py t5nn t5nt dwf dwb attnf attnb lnb", "captures": caps}

os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
dst = os.path.join(ROOT, "profiles", f"{prefix}_ncu_summary_{tag}.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst)
