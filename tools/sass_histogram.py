"""SASS opcode histogram of the built library, per kernel (cuobjdump -sass): the evidence for which hardware paths the
kernels use -- DMMA (FP64 tensor), UBLKCP / SYNCS (bulk async copies completing on mbarriers: the TMA engine),
LDGSTS (cp.async), scalar DFMA/DMUL/DADD.  usage: sass_histogram.py [library.so]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gptools_b200/csrc/libgptb200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for l in txt.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        per[cur][m.group(2)] += 1
KEYS = ["DMMA", "UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "LDGSTS", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "BAR"]
tot = collections.Counter()
print("%-52s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
for name, c in per.items():
    n = sum(c.values())
    if n == 0:
        continue
    tot.update(c)
    print("%-52s %7d " % (name[:52], n) + " ".join("%7d" % c[k] for k in KEYS))
print("%-52s %7d " % ("TOTAL", sum(tot.values())) + " ".join("%7d" % tot[k] for k in KEYS))
print("\nmnemonics seen with the prefixes U (uniform datapath / TMA), SYNCS (mbarrier), DMMA:")
for k in sorted(tot):
    if k.startswith(("UBLKCP", "UTMA", "SYNCS", "DMMA", "LDGSTS", "UTC", "LDTM", "STTM")):
        print("   %-12s %d" % (k, tot[k]))
