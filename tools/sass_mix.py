"""SASS instruction mix of a device function of the batched kernel, attributed by -lineinfo source lines.
usage: sass_mix.py <object.o> <kernel-name-substring> <function-name> [...]      (development aid; needs cuobjdump + nvdisasm)
e.g.   python tools/sass_mix.py gptools_b200/csrc/batched4.o Li2E gen_ktot_tab grad_tab"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, ksub = os.path.abspath(sys.argv[1]), sys.argv[2]
names = sys.argv[3:]
src_path = obj[:-2] + ".cu"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
counts = collections.defaultdict(collections.Counter)
fn = line = None
for l in txt:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and fn and ksub in fn:
        counts[line][m.group(2).split(".")[0]] += 1
src = open(src_path).read().split("\n")
base = os.path.basename(src_path)
for name in names:
    start = next(i + 1 for i, l in enumerate(src) if re.match(r"(__device__|__global__).*\b%s\(" % name, l))
    end = next(j + 1 for j in range(start, len(src)) if src[j] == "}")
    tot = collections.Counter()
    for (f, ln), c in counts.items():
        if f == base and start <= ln <= end:
            tot.update(c)
    fp64 = sum(v for k, v in tot.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print("%s (lines %d-%d): %d instructions, %d scalar FP64, %d branches: %s" % (
        name, start, end, sum(tot.values()), fp64, tot["BRA"], dict(tot.most_common(14))))
