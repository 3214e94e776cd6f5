"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked text summaries under profiles/.
   python tools/summarize_profiles.py <tag>        (reads gpurun_out/launches_<tag>.csv and gpurun_out/<tag>_*.ncu-rep)
Runs in the build container (ncu -i needs no GPU)."""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
]


def launch_summary(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.isfile(path):
        return
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
            rows.append((int(row["ID"]), row["Kernel Name"].split("(")[0], ns, row["Grid Size"], row["Block Size"]))
    adam = [i for i, r in enumerate(rows) if "adam_kernel" in r[1]]
    seg = rows[adam[-2] + 1:adam[-1] + 1]          # one full step: after the previous Adam up to and including this one
    tot = sum(r[2] for r in seg)
    agg = collections.OrderedDict()
    for _, k, ns, _, _ in seg:
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += ns
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none   python bench.py --no-graph --steps 1 --warmup 3 "
           "--no-kernel-timing --no-cpu-baseline",
           f"# one full training step (8 labeled + 8 unlabeled clips, --bv; both passes batched = 32 clip-passes), "
           f"launch IDs {seg[0][0]}..{seg[-1][0]}: {len(seg)} launches, sum of kernel durations {tot / 1e6:.3f} ms",
           "# ncu serialises launches with cold caches: compare SHARES with the live CUDA-event numbers of the bench line",
           f"{'kernel':62s} {'launches':>8s} {'ms':>9s} {'share':>7s}"]
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n[:62]:62s} {c:8d} {ns / 1e6:9.3f} {100 * ns / tot:6.2f}%")
    open(os.path.join(OUT, f"{tag}_launch_summary.txt"), "w").write("\n".join(out) + "\n")
    with open(os.path.join(OUT, f"{tag}_launch_list_eager_step.csv"), "w") as f:
        f.write("id,kernel,duration_ns,grid,block\n")
        for i, k, ns, g, b in seg:
            f.write(f'{i},"{k}",{ns:.0f},"{g}","{b}"\n')
    print("\n".join(out[:14]))


def full_summary(tag):
    out = []
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        out.append(f"===== {os.path.basename(rep)}   (ncu --set full --clock-control none, same bench command; one row block per launch)")
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0]
            out.append(f"--- {name}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
            for key, label in KEYS:
                if key in h:
                    i = h.index(key)
                    out.append(f"    {label:26s} {r[i]:>16s} {units[i]}")
            if "dram__bytes_read.sum" in h:
                def gb(k):
                    i = h.index(k)
                    v = float(r[i].replace(",", ""))
                    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(units[i], 1e-9)
                out.append(f"    {'traffic (read+write)':26s} {gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'):16.3f} GB")
    if out:
        open(os.path.join(OUT, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
        print("\n".join(out))


def traffic_summary(tag):
    """gpurun_out/dram_<tag>.csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:igemm` over the
    conv launches of ONE step -> profiles/<round>_traffic.json (read by bench.py for roofline.traffic)."""
    import json
    path = os.path.join(ROOT, "gpurun_out", f"dram_{tag}.csv")
    if not os.path.isfile(path):
        return
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    tot, ids = 0.0, set()
    per = collections.OrderedDict()
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(row["Metric Value"].replace(",", ""))
            b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1)
            tot += b
            ids.add(row["ID"])
            per[row["ID"]] = per.get(row["ID"], 0.0) + b
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:igemm (one step, {len(ids)} launches), "
                     f"gpurun_out/dram_{tag}.csv", "clips_per_gpu": 16, "launches": len(ids), "igemm_dram_bytes_per_step": tot,
           "largest_launches_bytes": sorted(per.values(), reverse=True)[:8]}
    json.dump(out, open(os.path.join(OUT, f"{tag[:3]}_traffic.json"), "w"), indent=1)      # r02b -> r02_traffic.json
    print(out)


if __name__ == "__main__":
    t = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    launch_summary(t)
    full_summary(t)
    traffic_summary(t)
