"""SASS opcode summary of libidb_b200.so (cuobjdump -sass): tcgen05 / TMEM / TMA / mbarrier / atomics counts, whole library and
per kernel.  usage: python tools/sass_summary.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "faceposegenerator_b200", "libidb_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
CLASSES = [("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"), ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"),
           ("UTC*MMA (tcgen05.mma) .2CTA", r"\bUTC[A-Z]*MMA[.\w]*\.2CTA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
           ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("SYNCS (mbarrier)", r"\bSYNCS"),
           ("RED/ATOM (global atomics)", r"\b(RED|ATOMG|ATOM)\b"), ("MUFU", r"\bMUFU"), ("HMMA (legacy mma.sync)", r"\bHMMA")]
total = collections.Counter()
per = []
chunks = re.split(r"\n\s*Function : ", sass)[1:]
for name, chunk in zip(names, chunks):
    c = collections.Counter()
    for label, rx in CLASSES:
        c[label] = len(re.findall(rx, chunk))
    total.update(c)
    per.append((name, c))
print("SASS opcode summary of faceposegenerator_b200/libidb_b200.so (cuobjdump -sass, sm_100a), HEAD of round 2 (tools/sass_summary.py)")
print("mnemonics: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA (cp.async.bulk.tensor), HMMA = legacy mma.sync (must be 0)\n")
print("whole library:")
for label, _ in CLASSES:
    print(f"{total[label]:8d}  {label}" + ("  <- none" if label.startswith("HMMA") and total[label] == 0 else ""))
print("\nper kernel (kernels with tensor-core / TMEM / TMA instructions):")
for name, c in per:
    if c["UTC*MMA (tcgen05.mma)"] or c["LDTM (tcgen05.ld)"] or c["UTMALDG (TMA load)"]:
        short = {"LDTM": c["LDTM (tcgen05.ld)"], "STTM": c["STTM (tcgen05.st)"], "RED/ATOM": c["RED/ATOM (global atomics)"],
                 "UTC*MMA": c["UTC*MMA (tcgen05.mma)"], "UTCBAR": c["UTCBAR (tcgen05.commit)"], "UTMALDG": c["UTMALDG (TMA load)"],
                 "UTMASTG": c["UTMASTG (TMA store)"]}
        print(f"  {name}\n      " + ", ".join(f"{k} {v}" for k, v in short.items() if v))
