"""Per-kernel SASS mnemonic counts of the in-tree libfpose.so (the .so itself is git-ignored): which kernels are
tcgen05 / TMEM / TMA code and which are plain SIMT.  Written to profiles/sass_summary_r02.txt.

    python tools/sass_summary.py
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "foundationpose_b200", "lib", "libfpose.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "HMMA", "ATOMS", "ATOMG", "RED", "MUFU", "LDG", "STG",
         "ACQBULK", "CCTL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            if op in WATCH:
                kernels[cur][op] += 1
                if op in ("UTCHMMA", "UTMALDG", "UTCBAR") and ".2CTA" in m.group(2):
                    kernels[cur][op + ".2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS summary of foundationpose_b200/lib/libfpose.so (sm_100a), `cuobjdump -sass` mnemonic counts per kernel",
             "# tcgen05 = UTCHMMA (MMA), LDTM/STTM (TMEM load/store), UTCBAR (commit); TMA = UTMALDG / UTMASTG; HMMA = mma.sync (none expected)", ""]
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*$", "", dn.replace("(anonymous namespace)::", ""))[:110]
        watch = ", ".join(f"{k} {v}" for k, v in sorted(c.items()) if k != "_total")
        lines.append(f"{short}\n    instructions {c['_total']}: {watch or '-'}")
        tot.update(c)
    lines += ["", "TOTAL: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items()))]
    path = os.path.join(ROOT, "profiles", "sass_summary_r02.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines[-3:]))
    print("wrote", path)


if __name__ == "__main__":
    main()
