"""Per-kernel SASS evidence for profiles/: counts of the Blackwell-native mnemonics in the shipped libarx.so
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG = bulk-copy / TMA engine, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, MUFU.EX2, packed FP32 FFMA2/FADD2/FMUL2, legacy HMMA).  usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "isbfsar_b200", "libarx.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMUL2", "HMMA", "LDGSTS"]
cur, counts, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = {p: 0 for p in pats}
        counts[cur]["instructions"] = 0
        order.append(cur)
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instructions"] += 1
        for p in pats:
            if re.search(r"\b" + re.escape(p), line):
                counts[cur][p] += 1
print("SASS mnemonic counts per kernel, cuobjdump -sass isbfsar_b200/libarx.so (sm_100a)")
print("%-58s %7s " % ("kernel", "instr") + " ".join("%8s" % p for p in pats))
tot = {p: 0 for p in pats}
for k in sorted(order, key=lambda k: -counts[k]["UTCHMMA"]):
    c = counts[k]
    print("%-58s %7d " % (k[:58], c["instructions"]) + " ".join("%8d" % c[p] for p in pats))
    for p in pats:
        tot[p] += c[p]
print("%-58s %7s " % ("TOTAL", "") + " ".join("%8d" % tot[p] for p in pats))
