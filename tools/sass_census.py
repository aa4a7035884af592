"""Opcode census of libvqb200.so per kernel (cuobjdump -sass; runs without a GPU):
    python tools/sass_census.py > profiles/r02_sass_census.md
The Blackwell-native instructions (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier;
HMMA would be the legacy mma.sync path (none here)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vq-vae-transformer-arc-welding_b200", "libvqb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "USETMAXREG", "FFMA2",
         "FADD2", "FMUL2", "FMNMX3", "MUFU", "HMMA", "LDGSTS", "ATOMS", "RED", "STL", "LDL"]
kern, counts, total = None, {}, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print("# SASS opcode census of libvqb200.so (sm_100a), static instruction counts per kernel\n")
print("`python tools/sass_census.py`, nvcc " + subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.split("release ")[1].split(",")[0] + "\n")
print("| kernel | instr | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for k in sorted(counts, key=lambda k: -total[k]):
    print(f"| `{k}` | {total[k]} | " + " | ".join(str(counts[k][w]) if counts[k][w] else "" for w in WATCH) + " |")
