"""Print the opcode sequence around the MUFU.EX2 clusters of a cuobjdump -sass dump (one letter per instruction):
shows how the compiler interleaved the exponent pipe with the FMA/ALU work.  Usage: python tools/sass_seq.py dump.sass"""
import collections
import re
import sys

ops = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m:
        ops.append(m.group(2).strip())


def name(o):
    t = o.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


names = [name(o) for o in ops]
short = {"MUFU.EX2": "M", "FADD2": "a", "FFMA2": "f", "FMNMX3": "3", "FMNMX": "x", "F2FP.BF16.F32.PACK_AB": "p",
         "F2FP.F16.F32.PACK_AB": "p", "IMAD": "i", "IMAD.MOV.U32": "m", "LDTM.x32": "L", "STTM.x32": "S", "NOP": "n",
         "R2UR": "u", "FMUL": "*", "FFMA": "F", "FADD": "A", "MOV": "v", "STL": "$", "LDL": "%"}
mu = [i for i, n in enumerate(names) if n == "MUFU.EX2"]
clusters, start, prev = [], mu[0], mu[0]
for i in mu[1:]:
    if i - prev > 60:
        clusters.append((start, prev))
        start = i
    prev = i
clusters.append((start, prev))
for a, b in clusters:
    if b - a < 20:
        continue
    c = collections.Counter(names[a:b + 1])
    print(f"--- instructions {a}..{b}: {dict(c.most_common(10))}")
    seq = "".join(short.get(n, "?") for n in names[max(0, a - 20):b + 40])
    for i in range(0, len(seq), 120):
        print(seq[i:i + 120])
