"""Print the SASS around the first dense run of a mnemonic (default DMMA) in one kernel of an object file.
usage: sass_window.py <object.o> <kernel-name-substring> [mnemonic] [which-run] [lines-after]     (development aid)"""
import os
import re
import subprocess
import sys
import tempfile

obj, ksub = os.path.abspath(sys.argv[1]), sys.argv[2]
mn = sys.argv[3] if len(sys.argv) > 3 else "DMMA"
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
after = int(sys.argv[5]) if len(sys.argv) > 5 else 90
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
for s in re.split(r"\n(?=\s*\.text\.)", txt):
    if ksub in s.split("\n")[0]:
        lines = s.split("\n")
        idx = [k for k, l in enumerate(lines) if mn in l]
        print("%d lines, %d %s" % (len(lines), len(idx), mn))
        runs = []
        k = 0
        while k < len(idx) - 16:
            if idx[k + 15] - idx[k] < 60:
                runs.append(idx[k])
                while k < len(idx) - 1 and idx[k + 1] - idx[k] < 60:
                    k += 1
            k += 1
        print("dense runs start at lines", runs)
        if runs:
            st = max(0, runs[min(which, len(runs) - 1)] - 30)
            print("\n".join(l[:120] for l in lines[st:st + 30 + after]))
        break
