"""cuobjdump -sass of the shipped .so -> per-kernel instruction histogram (markdown), profiles/r02_sass_histogram.md."""
import subprocess, re, collections, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "talkshow_b200", "libtalkshow_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FFMA", "LDGSTS", "LDS", "STS", "LDG", "STG",
        "BAR", "MUFU", "REDG", "ATOMG"]
kern = None; hist = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_n"] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c in ("UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "BAR", "MUFU", "UTCHMMA", "LDTM") and op.startswith(c)):
                hist[kern][c] += 1
                break
names = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
rows = sorted(zip(names, hist.values()), key=lambda kv: -kv[1]["_n"])
print("# SASS instruction histogram of the shipped `talkshow_b200/libtalkshow_b200.so` (round 2, final build)\n")
print("`python scratch/sass_histogram.py` (`cuobjdump -sass`), instructions per kernel (kernels with >= 200 instructions).")
print("Blackwell-native evidence (B200_PROFILING.md): `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor loads,")
print("`UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops, `FFMA2` = packed fp32 FMA; `HMMA` = legacy mma.sync (attention, positional conv).\n")
print("| kernel | instr | " + " | ".join(COLS) + " |")
print("|---|---:|" + "---:|" * len(COLS))
for n, h in rows:
    if h["_n"] < 200: continue
    print("| `%s` | %d | " % (n[:110], h["_n"]) + " | ".join(str(h[c]) if h[c] else "" for c in COLS) + " |")
