"""Per-source-line hot spots of the step kernels from an `ncu --set full --import-source on` report (runs here, no GPU):
    python scripts/ncu_source_hot.py gpurun_out/r2v_step.ncu-rep > profiles/r2_source_hotspots.md
For every kernel: the source lines (file:line, -lineinfo) with the most warp-stall samples and the most executed warp
instructions, with the dominant stall reasons of the line."""
import csv
import io
import os
import subprocess
import sys

# (ncu --kernel-name pattern on the base name, substring of the demangled function name, title)
KERNELS = [("k_pair_adjacency", "", "k_pair_adjacency"), ("k_final_tc", "", "k_final_tc"), ("k_hop", "k_hop<(int)2", "k_hop<2,1> (hop 0)"),
           ("k_hop", "k_hop<(int)1", "k_hop<1,0> (last hop)"), ("k_canon", "", "k_canon"), ("k_scatter", "", "k_scatter")]
TOP = 14


def source_rows(rep, pattern, func_substr):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + pattern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out, path, hdr, func = [], None, None, ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            func = r[1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-" and func_substr in func:
            out.append((path, hdr, r))
    return out


def main():
    rep = sys.argv[1]
    print(f"# Source-line hot spots of the step kernels ({os.path.basename(rep)}; N = 1M, K = 3, H = 32)\n")
    print("`ncu --set full --import-source on` (sampled warp stalls + executed warp instructions per source line, `-lineinfo`); "
          "made by `scripts/ncu_source_hot.py`.  Inlined device functions are attributed to their own file:line.\n")
    for pat, func_substr, title in KERNELS:
        rows = source_rows(rep, pat, func_substr)
        if not rows:
            continue
        hdr = rows[0][1]
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        agg = {}
        for path, _, r in rows:
            key = (os.path.basename(path), int(r[0]), r[1].strip())
            a = agg.setdefault(key, [0, 0, {}])
            a[0] += int(r[i_s] or 0)
            a[1] += int(r[i_i] or 0)
            for j, h in stall_cols:
                v = int(r[j] or 0)
                if v:
                    a[2][h] = a[2].get(h, 0) + v
        tot_s = sum(a[0] for a in agg.values()) or 1
        tot_i = sum(a[1] for a in agg.values()) or 1
        print(f"## {title}: {tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples\n")
        print("| file:line | source | samples | instructions | top stall reasons |")
        print("|---|---|---|---|---|")
        keys = sorted(agg, key=lambda k: -agg[k][0])[:TOP]
        more = [k for k in sorted(agg, key=lambda k: -agg[k][1])[:TOP // 2] if k not in keys]
        for key in keys + more:
            s, i, st = agg[key]
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            src = key[2].replace("|", "\\|")
            src = src if len(src) <= 90 else src[:87] + "..."
            print(f"| {key[0]}:{key[1]} | `{src}` | {100 * s / tot_s:.1f} % | {100 * i / tot_i:.1f} % | "
                  + ", ".join(f"{h[6:]} {100 * v / max(s, 1):.0f}%" for h, v in top) + " |")
        print()


if __name__ == "__main__":
    main()
