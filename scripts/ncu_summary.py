"""Turn an `ncu --set full` report into the markdown summary and the per-kernel DRAM traffic table kept under profiles/.
    python scripts/ncu_summary.py gpurun_out/r1_v10.ncu-rep profiles/r1_v10_ncu_full_summary.md profiles/r1_traffic.json
Runs here (no GPU needed): it only reads the report with `ncu -i ... --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
SHORT = {"k_final_tc": "final", "k_final": "final", "k_scan_sums": "scan_sums", "k_scan": "scan", "k_scatter": "scatter",
         "k_canon": "canon", "k_adjacency_t": "adjacency", "k_pair_adjacency": "pair_adjacency", "k_hop<2": "hop0", "k_hop<3": "hop0",
         "k_hop<(int)2": "hop0", "k_hop<(int)3": "hop0", "k_hop<1": "hop_last", "k_hop<(int)1": "hop_last"}


def main():
    rep, out_md, out_json = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    md = [f"# ncu --set full, one launch of each step kernel ({rep.split('/')[-1]}; N=1M, K=3, H=32, d~5)", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_\" -s 49 -c 7 python scripts/ncu_step.py 1000000 9`",
          "(cold-cache, serialised launches: compare shares, not absolutes)", ""]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        md += [f"## {name}", "", "| metric | value |", "|---|---|"]
        for w in WANT:
            if w in ix:
                md.append(f"| {w} | {r[ix[w]]} {units[ix[w]]} |")
        stalls = [(h, float(r[i])) for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled")
                  and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        stalls.sort(key=lambda x: -x[1])
        tot = sum(v for _, v in stalls) or 1.0
        md.append("| top stalls (share of stalled warp-cycles per issue) | " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {100 * v / tot:.0f}%"
            for h, v in stalls[:5]) + " |")
        md.append("")
        for key, short in SHORT.items():
            if key in name:
                rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
                scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                traffic[short] = rd * scale[units[ix["dram__bytes_read.sum"]]] + wr * scale[units[ix["dram__bytes_write.sum"]]]
                break
    open(out_md, "w").write("\n".join(md) + "\n")
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_sha
    traffic["_kernel_source_sha"] = kernel_source_sha()     # bench.py drops `roofline.traffic` when the kernels have changed since
    json.dump(traffic, open(out_json, "w"), indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
