"""Per-kernel SASS mnemonic counts of the in-tree library (cuobjdump -sass): the instructions that show which hardware
paths each kernel uses -- UBLKCP (TMA bulk copy), SYNCS (mbarrier), ATOMS.POPC.INC (merged shared atomics), IDP.4A (dp4a),
IMMA (tensor-core integer MMA), LDS/STS/LDG/STG widths, REDUX, spills (LDL/STL)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "qvz_b200/csrc/libqvz_gpu.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print(f"# cuobjdump -sass {lib}: cubins for {', '.join(arch)}")
watch = ["UBLKCP", "SYNCS", "ATOMS.POPC.INC", "ATOMS", "ATOMG", "RED", "IDP.4A", "IMMA", "LDS.128", "LDS.64", "LDS", "STS", "LDG.E.128", "LDG", "STG",
         "PRMT", "VABSDIFF4", "SHF", "REDUX", "LDL", "STL", "LDC", "UTMALDG", "UTCMMA", "BAR.SYNC"]
cur, counts, total = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in watch:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
                break
for k, c in counts.items():
    if not k.startswith("qvz_"):
        continue
    print(f"{k}: {total[k]} instructions; " + ", ".join(f"{w} {n}" for w, n in c.items() if n))
