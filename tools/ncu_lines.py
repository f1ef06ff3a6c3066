"""Per-source-line and per-warp-role summary of an ncu report captured with --import-source on (-lineinfo build).

usage: python tools/ncu_lines.py report.ncu-rep [top_n]
Roles are the regions between the '// -------- <role>' banner comments of the kernel source.
"""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path = None
    hdr = None
    lines = []          # (file, line_no, text, samples, instr, stall dict)
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue
        idx = {h: i for i, h in enumerate(hdr)}
        num = lambda v: int(v) if v.strip().isdigit() else 0
        stalls = {h[6:]: num(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        if not r[0].strip().isdigit():
            continue
        lines.append((path, int(r[0]), r[1], num(r[idx["# Samples"]]), num(r[idx["Instructions Executed"]]), stalls))
    total = sum(l[3] for l in lines)
    tinstr = sum(l[4] for l in lines)
    print(f"total samples {total}, warp instructions {tinstr}")
    print(f"--- top {top_n} lines by samples")
    for f, n, text, s, ins, st in sorted(lines, key=lambda l: -l[3])[:top_n]:
        top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{100 * s / max(total, 1):5.1f}% {ins:10d}  {f.split('/')[-1]}:{n:<4d} {text.strip()[:70]:70s} [{top}]")
    # ---- per role (main kernel file only)
    main_file = Counter(l[0] for l in lines if l[3]).most_common(1)[0][0]
    try:
        src = open(main_file).read().splitlines()
    except OSError:
        return
    banners = [(i + 1, re.sub(r"[-/ ]+", " ", t).strip()) for i, t in enumerate(src) if re.match(r"\s*// -{20,} \S", t)]
    role_s, role_i, role_st = Counter(), Counter(), defaultdict(Counter)
    for f, n, text, s, ins, st in lines:
        role = "other files" if f != main_file else "prologue"
        if f == main_file:
            for ln, name in banners:
                if n >= ln:
                    role = name
        role_s[role] += s
        role_i[role] += ins
        for k, v in st.items():
            role_st[role][k] += v
    # ---- per role by SASS address: instructions inlined from headers are attributed to the role whose code surrounds them
    sass = []   # (address, file, line, samples, instr, stalls)
    cur = None
    hdr2 = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1]
            continue
        if r and r[0] == "Line No":
            hdr2 = r
            continue
        if hdr2 is None or len(r) < len(hdr2):
            continue
        idx = {h: i for i, h in enumerate(hdr2)}
        num = lambda v: int(v) if v.strip().isdigit() else 0
        if r[0].strip().isdigit():
            cur = (path, int(r[0]))
            continue
        if cur and r[2].startswith("0x"):
            stalls = {h[6:]: num(r[i]) for i, h in enumerate(hdr2) if h.startswith("stall_") and "Not Issued" not in h}
            sass.append((int(r[2], 16), cur[0], cur[1], num(r[idx["# Samples"]]), num(r[idx["Instructions Executed"]]), stalls, r[3].strip()))
    sass.sort(key=lambda t: t[0])
    def role_of_line(n):
        role = "prologue"
        for ln, name in banners:
            if n >= ln:
                role = name
        return role
    arole_s, arole_i, arole_st = Counter(), Counter(), defaultdict(Counter)
    last = "prologue"
    for a, f, n, smp, ins, st, txt in sass:
        if f == main_file:
            last = role_of_line(n)
        arole_s[last] += smp
        arole_i[last] += ins
        for k, v in st.items():
            arole_st[last][k] += v
    tot2 = sum(arole_s.values())
    print("--- per role, by SASS address (inlined helpers included)")
    for role, smp in arole_s.most_common():
        top = ", ".join(f"{k}:{v}" for k, v in arole_st[role].most_common(5) if v)
        print(f"{100 * smp / max(tot2, 1):5.1f}% samples  {arole_i[role]:10d} instr  {role[:44]:44s} [{top}]")
    print("--- per role, by source line")
    for role, s in role_s.most_common():
        top = ", ".join(f"{k}:{v}" for k, v in role_st[role].most_common(4) if v)
        print(f"{100 * s / max(total, 1):5.1f}% samples  {100 * role_i[role] / max(tinstr, 1):5.1f}% instr ({role_i[role]:10d})  {role[:44]:44s} [{top}]")


if __name__ == "__main__":
    main()
