"""SASS census of the built library (cuobjdump -sass meteotools_b200/lib/fastcompute.so): per kernel that uses TMA
or cp.async the number of instructions and of UTMALDG / SYNCS / LDGSTS / DFMA / LDS / STG, plus two excerpts (the
plane loads of k_stencil_tma's producer thread, the pair step of k_setdata_tma's phase C).
    python profiles/sass_census.py > profiles/sass_r02.txt"""
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "meteotools_b200", "lib", "fastcompute.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((m.group(1), m.group(2).strip()))
dem = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(funcs, dem))


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    return n.split("(")[0]


def count(ins, op):
    return sum(1 for _, s in ins if re.match(r"(@!?U?P\d+\s+)?" + op + r"\b", s))


print("cuobjdump -sass meteotools_b200/lib/fastcompute.so  (sm_100a cubins only; built by meteotools_b200/build.py)")
print("architectures in the fat binary:", archs)
tot = {op: sum(count(i, op) for i in funcs.values()) for op in ("UTMALDG", "SYNCS", "LDGSTS")}
print(f"whole library: UTMALDG (TMA tensor loads) {tot['UTMALDG']}, SYNCS (mbarrier) {tot['SYNCS']}, "
      f"LDGSTS (cp.async) {tot['LDGSTS']}, kernels {len(funcs)}\n")
print("kernels that use TMA or cp.async:  name | instructions | UTMALDG | SYNCS | LDGSTS | DFMA | LDS | STG")
for f, ins in funcs.items():
    if count(ins, "UTMALDG") or count(ins, "LDGSTS"):
        c = [count(ins, op) for op in ("UTMALDG", "SYNCS", "LDGSTS", "DFMA", "LDS", "STG")]
        print(f"  {short(names[f]):90s} | {len(ins):5d} | " + " | ".join(f"{x:3d}" for x in c))


def excerpt(pred, title, start_op, before, after):
    for f, ins in funcs.items():
        if pred(names[f]):
            idx = [i for i, (_, s) in enumerate(ins) if start_op in s]
            if idx:
                print(f"\n{title} ({short(names[f])}):")
                for a, s in ins[max(idx[0] - before, 0): idx[0] + after]:
                    print(f"    /*{a}*/ {s}")
            return


excerpt(lambda n: "k_stencil_tma<0, false, true" in n and n.rstrip(")").find(", 16>") > 0,
        "k_stencil_tma, Cartesian, all outputs, row shift 16: the producer thread's plane loads (csrc/diag_tma.cu, load_3d)",
        "UTMALDG", 12, 60)
excerpt(lambda n: "k_setdata_tma<double, false, false, 5, true>" in n,
        "k_setdata_tma<double, ..., PP = 5, EXACT>: one pair step of phase C, shared-source-cell instantiation "
        "(csrc/setdata_tma.cu: records, five integer instructions for the four corner addresses, the lerps of "
        "fastcompute.f90:100-102, one multiply-add per row address, 16-byte streaming stores)",
        "STG.E.EF.128", 95, 6)
