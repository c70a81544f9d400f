#!/usr/bin/env python
"""Summarise an `ncu --csv --page raw` capture of one warm step (scripts/one_step.py under --profile-from-start off):
per kernel family launches, device time, DRAM bytes (= dram__bytes.sum.per_second x duration) and the achieved DRAM
throughput / tensor-pipe activity ncu reports.    python scripts/summarize_ncu_step.py <raw.csv> [--launches]"""
import csv
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 2:]
    idx = {n: i for i, n in enumerate(h)}
    cn, cd = idx["Kernel Name"], idx["gpu__time_duration.sum"]
    cbw, cpct = idx["dram__bytes.sum.per_second"], idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    ctp = idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
    cg = idx.get("launch__grid_size")
    recs = []
    for r in data:
        if len(r) < len(h):
            continue
        name = r[cn].split("(")[0].replace("void ", "").replace("rv::", "")
        ns = f(r[cd])
        recs.append((name, ns / 1e3, f(r[cbw]) * ns / 1e9 / 1e6, f(r[cbw]) / 1e9, f(r[cpct]), f(r[ctp]), r[cg] if cg is not None else ""))
    tot = sum(r[1] for r in recs)
    print(f"# {len(recs)} launches, {tot:.1f} us of serialised cold-cache device time (compare shares, not absolutes)")
    print("kernel,launches,us,share_pct,dram_MB,dram_GBs_mean,dram_pct_of_peak_max,tensor_pipe_pct_max")
    agg = {}
    for name, us, mb, gbs, pct, tp, _ in recs:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] += mb
        a[3] = max(a[3], pct)
        a[4] = max(a[4], tp)
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{name}\",{a[0]},{a[1]:.1f},{100 * a[1] / tot:.1f},{a[2]:.1f},{a[2] / a[1] * 1e3 if a[1] else 0:.0f},{a[3]:.1f},{a[4]:.1f}")
    if "--launches" in sys.argv:
        print("\n# every launch in order")
        print("kernel,us,dram_MB,dram_GBs,dram_pct_of_peak,tensor_pipe_pct,grid")
        for name, us, mb, gbs, pct, tp, g in recs:
            print(f"\"{name}\",{us:.1f},{mb:.1f},{gbs:.0f},{pct:.1f},{tp:.1f},{g}")


if __name__ == "__main__":
    main()
