"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares (markdown).
  python tools/launch_shares.py gpurun_out/r02_launches_hot.csv "title" > profiles/r02_step_kernel_shares.md"""
import collections
import csv
import re
import sys


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "kernel shares"
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
            rows.append((r["Kernel Name"], us))
    tot = sum(u for _, u in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, u in rows:
        n = re.sub(r"\(anonymous namespace\)::|vtm::|<unnamed>::", "", n)
        n = re.sub(r"\(.*", "", n)[:70]
        agg[n][0] += 1
        agg[n][1] += u
    ours = ("gemm_kernel", "flash_attn", "fa_combine", "normalize_split", "gather_rows", "radix_", "compose_maps", "decode_match",
            "reduce_", "sim_argmax")
    print(f"# {title}\n")
    print(f"{len(rows)} launches, {tot / 1000:.2f} ms of GPU time (ncu times are cold-cache and serialised: read the SHARES)\n")
    print("| kernel | launches | total us | share | ours |")
    print("|---|---|---|---|---|")
    mine = 0.0
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        own = any(o in n for o in ours)
        mine += u if own else 0
        if u / tot >= 0.003:
            print(f"| `{n}` | {c} | {u:.1f} | {100 * u / tot:.1f} % | {'x' if own else ''} |")
    print(f"\nKernels of this library: {100 * mine / tot:.1f} % of the listed GPU time.")


if __name__ == "__main__":
    main()
