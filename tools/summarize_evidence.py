#!/usr/bin/env python
"""Turns the raw files tools/gpu_evidence.sh leaves in gpurun_out/ into the tracked summaries under profiles/:
r02_launches.csv + r02_launches_summary.txt, r02_traffic.csv + traffic.json, r02_bench*.json."""
import collections, csv, io, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def rows_of(path):
    return list(csv.DictReader(io.StringIO("".join(l for l in open(path) if l.startswith('"')))))


# ---- launch list
rd = rows_of(os.path.join(G, "r02_launches.csv"))
tot, cnt = collections.Counter(), collections.Counter()
for x in rd:
    if x["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(x["Metric Unit"], 1)
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("(anonymous namespace)", "<unnamed>")[:120]
    tot[n] += v
    cnt[n] += 1
T = sum(tot.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1",
       f"total {int(T)} ns over {sum(cnt.values())} launches"]
for n, v in tot.most_common(16):
    out.append(f"{100 * v / T:5.1f}%  {int(v):12d} ns  x {cnt[n]:3d}  {n}")
st = sum(v for n, v in tot.items() if "stream_kernel" in n)
out.append(f"stream_kernel instantiations together: {100 * st / T:.1f}% of all device time in this short run (the rest: graph "
           "construction = CUB sorts/scans + plan kernels, torch's input generation, host-pipeline column copies)")
open(os.path.join(P, "r02_launches_summary.txt"), "w").write("\n".join(out) + "\n")
shutil.copy(os.path.join(G, "r02_launches.csv"), os.path.join(P, "r02_launches.csv"))

# ---- DRAM traffic per feature length (last call of each F: 3 warm-up + 2 timed calls, two kernels each)
shutil.copy(os.path.join(G, "r02_traffic.csv"), os.path.join(P, "r02_traffic.csv"))
by = collections.OrderedDict()
for x in rows_of(os.path.join(G, "r02_traffic.csv")):
    by.setdefault(x["ID"], {"name": x["Kernel Name"]})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
ks = list(by.values())
N, M, Z = 1261888, 509632, 2226407
per, tot_b, tot_alg = {}, 0.0, 0
for i, F in enumerate((32, 64, 128, 256, 512)):
    a, b = ks[i * 10 + 8], ks[i * 10 + 9]
    assert ", 0, 0, 3," in a["name"] and ", 0, 1, 3," in b["name"], (a["name"], b["name"])
    alg = 8 * F * N + 4 * Z + 12 * M + 4 * N + 4
    d = sum(k[m] for k in (a, b) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    stage = lambda k: {"dram_read": k["dram__bytes_read.sum"], "dram_write": k["dram__bytes_write.sum"],
                       "us_under_ncu": round(k["gpu__time_duration.sum"] / 1e3, 1)}
    per[str(F)] = {"stage_A": stage(a), "stage_B": stage(b), "dram_bytes": d, "algorithmic_bytes": alg, "ratio": round(d / alg, 3)}
    tot_b += d
    tot_alg += alg
json.dump({"source": "round 2: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                     "-k regex:stream_kernel, python tools/tune.py --features 32,64,128,256,512 --iters 2 (profiles/r02_traffic.csv); "
                     "last call of each F (stage A + stage B kernels)",
           "per_F": per, "dram_bytes_per_step": tot_b, "algorithmic_bytes_per_step": tot_alg, "ratio": round(tot_b / tot_alg, 3)},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
# ---- bench lines
for n in ("r02_bench", "r02_bench_reference"):
    line = [l for l in open(os.path.join(G, n + ".json")) if l.startswith("{")][-1]
    open(os.path.join(P, n + ".json"), "w").write(line)
print("traffic ratio", round(tot_b / tot_alg, 3), "| stream kernels", round(100 * st / T, 1), "% of the short run")
