"""Per-kernel counts of the Blackwell tensor-core / TMA / tensor-memory instructions in the built library: the evidence that
a kernel really is tcgen05 + TMEM + TMA (B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, UTMALDG / UTMASTG = TMA tensor
load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk).
usage: python tools/sass_summary.py > profiles/sass_summary.txt   (no GPU needed: cuobjdump -sass on liblssvc_b200.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lssvc_b200", "csrc", "liblssvc_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCCP", "HMMA", "FFMA", "SYNCS", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, regs, cur = collections.OrderedDict(), {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: cuobjdump -sass / -res-usage, sm_100a; columns = static instruction counts per kernel")
    print(f"# {'kernel':70s} {'instr':>6s} {'regs':>4s} {'smem':>6s} " + " ".join(f"{k:>7s}" for k in MNEMONICS))
    tot = collections.Counter()
    for k, c in counts.items():
        short = re.sub(r"\(anonymous namespace\)::", "", names.get(k, k))
        short = re.sub(r"\(.*", "", short)[:70]
        r = regs.get(k, (0, 0))
        print(f"{short:72s} {c['_total']:6d} {r[0]:4d} {r[1]:6d} " + " ".join(f"{c[m]:7d}" for m in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':72s} {tot['_total']:6d} {'':4s} {'':6s} " + " ".join(f"{tot[m]:7d}" for m in MNEMONICS))


if __name__ == "__main__":
    main()
