"""One-call profiling recipe for the training step (run ON the GPU box, inside ONE gpurun call):

    gpurun --timeout 1500 -- 'python tools/profile_step.py r02a'          # then here: python tools/summarize_profiles.py r02a

1. ncu launch list (gpu__time_duration.sum) of `bench.py --no-graph --steps 1 --warmup 3 ...` -> gpurun_out/launches_<tag>.csv
2. from that list: the last complete step (between two adam_kernel launches), its top kernels, and for each the
   ordinal of its longest launches among the launches of the same kernel name (what `ncu -k NAME --launch-skip` counts)
3. `ncu --set full` of those launches -> gpurun_out/<tag>_<kernel>.ncu-rep   (source import only for the igemm kernels;
   total size kept under gpurun's 64 MiB return limit)
4. DRAM byte pass over the conv launches of that step -> gpurun_out/dram_<tag>.csv
Numbers printed by bench.py under ncu are never bench values."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
BENCH = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-graph", "--steps", "1", "--warmup", "3", "--no-kernel-timing",
         "--no-cpu-baseline"]


def read_launches(path):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            # base function name: ncu's -k matches it without template arguments ("igemm_fprop_kernel<0>" -> "igemm_fprop_kernel")
            name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0].strip()
            rows.append((int(row["ID"]), name, v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)))
    return rows


def plan_captures(rows, top_kernels=6, per_kernel=(("igemm_fprop_kernel", 4), ("igemm_wgrad_kernel", 2))):
    """-> (step slice, [(kernel short name, ordinal among same-name launches, duration ns)], conv skip, conv count)."""
    adam = [i for i, r in enumerate(rows) if "adam_kernel" in r[1]]
    if len(adam) < 2:
        raise SystemExit("launch list holds fewer than two optimiser launches: cannot delimit a step")
    lo, hi = adam[-2] + 1, adam[-1] + 1
    ordinal, seen = [], collections.Counter()
    for _, name, _ in rows:
        ordinal.append(seen[name])
        seen[name] += 1
    tot = collections.Counter()
    for i in range(lo, hi):
        tot[rows[i][1]] += rows[i][2]
    want = dict(per_kernel)
    caps = []
    for name, _ in tot.most_common(top_kernels):
        short = name.split("::")[-1]
        n = next((c for k, c in want.items() if k in name), 1)
        best = sorted((i for i in range(lo, hi) if rows[i][1] == name), key=lambda i: -rows[i][2])[:n]
        caps += [(short, ordinal[i], rows[i][2]) for i in sorted(best)]
    conv = [i for i in range(lo, hi) if "igemm" in rows[i][1]]
    conv_before = sum(1 for i in range(lo) if "igemm" in rows[i][1])
    return (lo, hi), caps, conv_before, len(conv)


def ncu(args, log):
    cmd = ["ncu", "--clock-control", "none"] + args + BENCH
    with open(os.path.join(OUT, log), "w") as f:
        return subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=900).returncode


def main(tag):
    os.makedirs(OUT, exist_ok=True)
    lst = os.path.join(OUT, f"launches_{tag}.csv")
    ncu(["--metrics", "gpu__time_duration.sum", "--csv", "--log-file", lst], f"ncu_list_{tag}.log")
    rows = read_launches(lst)
    (lo, hi), caps, conv_skip, conv_count = plan_captures(rows)
    print(f"step = launches {rows[lo][0]}..{rows[hi - 1][0]} ({hi - lo} launches, {sum(r[2] for r in rows[lo:hi]) / 1e6:.3f} ms)")
    budget = 56 << 20
    for short, ordn, ns in caps:
        rep = os.path.join(OUT, f"{tag}_{short}_{ordn}")
        args = ["--set", "full", "-k", f"regex:{short}", "--launch-skip", str(ordn), "--launch-count", "1", "-f", "-o", rep]
        if "igemm" in short:
            args = ["--import-source", "on"] + args
        ncu(args, f"ncu_{tag}_{short}_{ordn}.log")
        size = os.path.getsize(rep + ".ncu-rep") if os.path.isfile(rep + ".ncu-rep") else 0
        budget -= size
        print(f"captured {short} #{ordn} ({ns / 1e6:.3f} ms): {size >> 20} MiB")
        if budget < (8 << 20):
            print("stopping captures: gpurun returns at most 64 MiB")
            break
    ncu(["--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "-k", "regex:igemm", "--launch-skip", str(conv_skip),
         "--launch-count", str(conv_count), "--csv", "--log-file", os.path.join(OUT, f"dram_{tag}.csv")], f"ncu_dram_{tag}.log")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "prof")
