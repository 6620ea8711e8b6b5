"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code paths (B200_PROFILING.md): run here, no GPU.
  python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import os
import re
import subprocess
import sys
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vidtome_b200", "libvidtome_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA", "ATOMG", "REDG", "USETMAXREG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, sizes = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.+?);", line)
        if m and cur:
            sizes[cur] += 1
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            op = ins.split()[0]
            for w in WANT:
                if op.startswith(w):
                    counts[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# r02 — SASS evidence per kernel of libvidtome_b200.so (cuobjdump -sass, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma kind::f16, `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = TMA tensor load, `SYNCS` = mbarrier ops,")
    print("`FFMA2` = packed fp32 FMA, `USETMAXREG` = setmaxnreg; `HMMA` (legacy mma.sync) must be absent.\n")
    cols = [w for w in WANT if any(c[w] for c in counts.values())] + (["HMMA"] if not any(c["HMMA"] for c in counts.values()) else [])
    cols = list(dict.fromkeys(cols))
    print("| kernel | SASS instrs | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::|vtm::|<unnamed>::", "", dn)
        short = re.sub(r"\(.*$", "", short)[:90]
        print(f"| `{short}` | {sizes[name]} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")


if __name__ == "__main__":
    main()
