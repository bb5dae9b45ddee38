#!/usr/bin/env python
"""SASS census of libmpcc_b200.so -> profiles/r2_sass_census.md (instruction mnemonics per kernel, an excerpt of k_mlp's inner loop)."""
import collections, re, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
txt = subprocess.run(["cuobjdump", "-sass", str(ROOT / "mpcc_manipulator_b200" / "libmpcc_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)[1:]
lines = ["# Round 2 - SASS census of libmpcc_b200.so (sm_100a)", "",
         "`cuobjdump -sass mpcc_manipulator_b200/libmpcc_b200.so` (tools/sass_census.py), instruction mnemonics per kernel.  The FP64 tensor instruction is `DMMA.8x8x4`",
         "(mma.sync.m8n8k4.f64).  tcgen05 (`UTC*MMA`), TMEM (`LDTM` / `STTM`) and TMA tensor loads (`UTMALDG`) do not appear: tcgen05 has no f64 kind, so the native path of",
         "an fp64 contraction on sm_100a is DMMA (DESIGN.md 3); `LDGSTS` is cp.async.",
         "", "| kernel | SASS instructions | DMMA | DFMA | DMUL+DADD | MUFU | LDGSTS | LDL+STL (local) | BAR | SHFL | UTC*MMA / LDTM / UTMALDG |", "|---|---|---|---|---|---|---|---|---|---|---|"]
exc = None
for f in funcs:
    name = f.split('\n')[0].strip()
    ins = re.findall(r'^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', f, flags=re.M)
    c = collections.Counter(i.split('.')[0] for i in ins)
    dm = sum(1 for i in ins if i.startswith('DMMA'))
    t5 = sum(1 for i in ins if i.startswith(('UTC', 'LDTM', 'STTM', 'UTMALDG')))
    pretty = re.search(r'(k_[a-z0-9_]+?)(?=E|P|N4|$)', re.sub(r'^_Z(N4mpcc)?\d+', '', name))
    pretty = pretty.group(1) if pretty else name[:40]
    lines.append(f"| `{pretty}` | {len(ins)} | {dm} | {c['DFMA']} | {c['DMUL'] + c['DADD']} | {c['MUFU']} | {c['LDGSTS']} | {c['LDL'] + c['STL']} | {c['BAR']} | {c['SHFL']} | {t5} |")
    if 'k_mlp' in name and exc is None:
        body = f.split('\n')
        idx = [i for i, l in enumerate(body) if 'DMMA' in l]
        exc = body[idx[40] - 6: idx[40] + 30]
lines += ["", "Excerpt of `k_mlp`'s inner loop (LDS.128 of B fragments feeding a run of DMMA.8x8x4):", "", "```"] + \
         [re.sub(r'\s+/\* 0x[0-9a-f]+ \*/', '', l).rstrip() for l in exc if '/*' in l and not re.match(r'^\s+/\* 0x', l)] + ["```"]
(ROOT / "profiles" / "r2_sass_census.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
