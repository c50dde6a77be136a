"""Kernel shares of one step from an ncu launch list (gpu__time_duration.sum per launch) next to the in-bench CUDA-event
shares (`roofline.per_kernel_ms` of a bench line).
    python scripts/launch_shares.py profiles/r2_launches.csv profiles/r2_bench_1gpu.json > profiles/r2_launch_shares.md"""
import csv
import json
import sys

SHORT = [("k_final_tc", "final"), ("k_final", "final"), ("k_scan_sums", "scan_sums"), ("k_scan", "scan"), ("k_scatter", "scatter"),
         ("k_canon", "canon"), ("k_pair_adjacency", "pair_adjacency"), ("k_adjacency_t", "adjacency"), ("k_hop<2", "hop0"),
         ("k_hop<(int)2", "hop0"), ("k_hop<3", "hop0"), ("k_hop<(int)3", "hop0"), ("k_hop<1", "hop_last"), ("k_hop<(int)1", "hop_last")]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = {}, {}
    for r in rows:
        if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
        if "k_final" in r[ik] and not (", 1>(" in r[ik] or "(bool)1>(" in r[ik]):
            continue                      # the open readout (chunked launches of the host-buffer e2e loop) is not in the step graph
        for key, short in SHORT:
            if key in r[ik]:
                tot[short] = tot.get(short, 0.0) + v
                cnt[short] = cnt.get(short, 0) + 1
                break
    mean = {k: tot[k] / cnt[k] for k in tot}
    s = sum(mean.values())
    bench = json.load(open(sys.argv[2]))["roofline"]["per_kernel_ms"]
    sb = sum(bench.values())
    print("# Kernel shares of one step: ncu launch list vs in-bench CUDA events (N = 1M, K = 3, H = 32)\n")
    print("ncu: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 8 --warmup 3` (cold-cache,")
    print("serialised launches: shares, not absolutes); bench: `roofline.per_kernel_ms` of the committed bench line.\n")
    print("| kernel | launches in the list | ncu mean us | ncu share | bench us | bench share |")
    print("|---|---|---|---|---|---|")
    for k in sorted(mean, key=lambda k_: -mean[k_]):
        b = bench.get(k)
        print(f"| {k} | {cnt[k]} | {mean[k]:.1f} | {100 * mean[k] / s:.1f} % | " + (f"{b * 1e3:.1f} | {100 * b / sb:.1f} % |" if b else "- | - |"))


if __name__ == "__main__":
    main()
