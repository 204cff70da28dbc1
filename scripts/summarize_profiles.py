"""Turn the ncu outputs of scripts/profile_r0N.sh (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r02      (after `gpurun -- bash scripts/profile_r02.sh r02`)"""
import csv
import gzip
import json
import re
import shutil
import subprocess
import sys
from collections import defaultdict

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01f"
OUT = "gpurun_out"
PEAK_HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def read_ncu_csv(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("spff::<unnamed>::", "").replace("<unnamed>::", "")
    return name.strip()[:110]


# 1) launch list ------------------------------------------------------------------------------------------------
rows = read_ncu_csv(f"{OUT}/launches_{TAG}.csv")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = short(r["Kernel Name"])
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"]) / 1e3   # ns -> us
total = sum(v[1] for v in agg.values())
import os
plain_path = f"{OUT}/prof_plain_{TAG}.json" if os.path.exists(f"{OUT}/prof_plain_{TAG}.json") else f"{OUT}/prof_plain.json"
plain = json.loads(open(plain_path).read().strip().splitlines()[-1])
ROUND = "2" if TAG.startswith("r02") else "1"
own = sum(v[1] for k, v in agg.items() if not k.startswith("at::"))
with open(f"profiles/{TAG}_launches_summary.md", "w") as f:
    f.write(f"# Round {ROUND} — ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
            f"Command: `SPFF_BENCH_SAMPLES=256 python bench.py --steps 1 --warmup 3 --no-cpu` (`--no-extras` from round 2 on; one sample group of 256 slices per "
            f"step; the first 3000 launches ≈ 5 train steps). Per-launch times are cold-cache and serialised: compare SHARES.\n"
            f"Raw list: `profiles/{TAG}_launches.csv.gz`. The same command without ncu: {plain['ms_per_step']:.1f} ms/step, "
            f"breakdown by C-ABI entry point (CUDA events) below the table.\n\n| share | launches | avg us | kernel |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f"| {100 * us / total:.2f}% | {n} | {us / n:.1f} | `{k}` |\n")
    f.write(f"\nTotal {len(rows)} launches, {total / 1e3:.1f} ms of kernel time; kernels of libspff_b200.so: {100 * own / total:.1f}% of it "
            f"(the rest: torch fills / copies / tiny table ops).\n\n")
    conv = sum(v[1] for k, v in agg.items() if k.startswith("conv3_fprop") or k.startswith("conv3_halo"))
    wg = sum(v[1] for k, v in agg.items() if k.startswith("conv3_wgrad"))
    b = plain["breakdown_ms"]
    tot = plain["breakdown_total_ms"]
    f.write(f"Share of the step: conv3_halo_kernel + conv3_fprop_kernel (fwd + dgrad) {100 * conv / total:.1f}%, conv3_wgrad(+reduce) {100 * wg / total:.1f}% under ncu; "
            f"bench.py's CUDA-event breakdown of the same build and command: fwd+dgrad "
            f"{b['spff_conv3d_k3_fwd_stats'] + b['spff_conv3d_k3_dgrad'] + b.get('spff_conv3d_k3_dgrad_stats', 0):.1f} ms and wgrad {b['spff_conv3d_k3_wgrad']:.1f} ms of {tot:.1f} ms "
            f"({100 * (b['spff_conv3d_k3_fwd_stats'] + b['spff_conv3d_k3_dgrad'] + b.get('spff_conv3d_k3_dgrad_stats', 0)) / tot:.1f}% / {100 * b['spff_conv3d_k3_wgrad'] / tot:.1f}%).\n\n"
            f"CUDA-event breakdown (ms per step of 256 slices): " + ", ".join(f"{k[5:]} {v}" for k, v in list(b.items())[:14]) + "\n")
with open(f"{OUT}/launches_{TAG}.csv", "rb") as src, gzip.open(f"profiles/{TAG}_launches.csv.gz", "wb") as dst:
    shutil.copyfileobj(src, dst)

# 2) conv DRAM traffic per launch ---------------------------------------------------------------------------------
rows = read_ncu_csv(f"{OUT}/conv_traffic_{TAG}.csv")
per = defaultdict(dict)
for r in rows:
    per[r["ID"]]["k"] = short(r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
    per[r["ID"]][r["Metric Name"]] = v * scale
fam = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    k = "conv3_fprop_kernel" if d["k"].startswith("conv3_fprop") else "conv3_halo_kernel" if d["k"].startswith("conv3_halo") else ("conv3_wgrad_kernel" if d["k"].startswith("conv3_wgrad_kernel") else d["k"])
    a = fam[k]
    a[0] += 1
    a[1] += d.get("dram__bytes_read.sum", 0)
    a[2] += d.get("dram__bytes_write.sum", 0)
    a[3] += d.get("gpu__time_duration.sum", 0)
traffic = {k: {"launches": a[0], "dram_read_bytes_per_launch": a[1] / a[0], "dram_write_bytes_per_launch": a[2] / a[0],
               "avg_us": a[3] / a[0]} for k, a in fam.items()}
json.dump({"command": "SPFF_BENCH_SAMPLES=256 python bench.py --steps 1 --warmup 3 --no-cpu", "samples_per_launch": 256,
           "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every conv3_* launch of the run", "kernels": traffic},
          open(f"profiles/{TAG}_conv_traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))

# 3) --set full captures -------------------------------------------------------------------------------------------
CAPTURES = ("halo", "rows", "wgrad", "wgrad_kh", "bw_maxpool", "bw_convt", "bw_head", "stem") if ROUND == "2" else ("bwd_reduce4", "bwd_apply4", "norm_act", "norm_act_pool", "maxpool_bwd", "wgrad32", "stem", "head_loss")
FULL_NAME = f"profiles/{TAG}_conv_kernels_ncu_full.md" if ROUND == "2" else f"profiles/{TAG}_bandwidth_kernels_ncu_full.md"
WANT = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
with open(FULL_NAME, "w") as f:
    f.write(f"# Round {ROUND} — `ncu --set full --clock-control none --import-source on` captures\n\n"
            "Command: `SPFF_BENCH_SAMPLES=256 python bench.py --steps 1 --warmup 3 --no-cpu`; first launches of each kernel in a step "
            "(level-1 tensors: 256 slices x 5 x 128 x 128 positions x 32 channels bf16 = 1.342 GB each). Times under ncu are cold-cache and "
            "serialised; mind the unit printed beside each metric (reads of the reduce kernels are GB, their writes MB). "
            f"HBM peak (measured copy): {PEAK_HBM:.0f} GB/s.\n\n")
    for name in CAPTURES:
        path = f"{OUT}/prof_{name}_{TAG}.ncu-rep"
        try:
            if os.path.exists(f"{OUT}/prof_{name}_{TAG}.csv") and os.path.getsize(f"{OUT}/prof_{name}_{TAG}.csv") > 0:
                out = open(f"{OUT}/prof_{name}_{TAG}.csv").read()      # exported on the GPU box (profile_r02.sh)
            else:
                out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        except Exception as e:
            f.write(f"## {name}\n\n(no capture: {type(e).__name__})\n\n")
            continue
        rr = list(csv.reader(out.splitlines()))
        hdr, units = rr[0], rr[1]
        f.write(f"## {name}\n\n| metric | " + " | ".join(f"launch {i}" for i in range(len(rr) - 2)) + " |\n|---|" + "---|" * (len(rr) - 2) + "\n")
        f.write("| kernel | " + " | ".join(f"`{short(r[hdr.index('Kernel Name')])[:60]}`" for r in rr[2:]) + " |\n")
        for m in WANT:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| {m} ({units[i]}) | " + " | ".join(r[i] for r in rr[2:]) + " |\n")
        f.write("\n")
print("ok")
