#!/usr/bin/env python
"""Where do k_sqp_warp's warp-stall samples go?  Joins an `ncu --set full --import-source on` report (source page, SASS
order) with nvdisasm's inline chains of the SAME build and prints samples / stall_no_instruction / executed instructions
per outer path and per innermost function.

    cuobjdump -xelf all build/k_sqp_warp.o ; nvdisasm -gi k_sqp_warp.sm_100a.cubin > dis.txt
    ncu -i report.ncu-rep --page source --csv -k regex:k_sqp_warp > src.csv
    python tools/ncu_hotspots.py src.csv dis.txt
"""
import collections, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_footprint import chains_of, ranges
import bisect

def main():
    src, dis = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    col = {n: hdr.index(n) for n in ["# Samples", "stall_no_inst", "Instructions Executed", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving"]}
    chains = chains_of(dis, lambda line: 'k_sqp_warp' in line and 'solve_ocp' not in line and 'r255' not in line)
    if len(chains) != len(data):
        sys.exit(f"report ({len(data)} instr) and disassembly ({len(chains)}) are different builds")
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mpcc_manipulator_b200', 'csrc')
    R = {f: ranges(os.path.join(root, f)) for f in ['sqp_warp.cuh', 'dev_stage.cuh', 'dev_track.cuh', 'dev_panda.cuh', 'dev_qp.cuh']}
    def fn(f, l):
        if f not in R: return f
        st = R[f]; i = bisect.bisect_right([s[0] for s in st], l) - 1
        return st[i][1] if i >= 0 else '?'
    keys = ('solve', 'factor', 'solve_step', 'gradient', 'ineq_steps', 'eval_horizon', 'run', 'issue_tile', 'qp_box_infeasible', 'gather_point', 'stream_constraints')
    outer = collections.defaultdict(lambda: [0] * 8); inner = collections.defaultdict(lambda: [0] * 8)
    tot = 0
    for r, ch in zip(data, chains):
        vals = [int(r[col[n]] or 0) for n in col] + [1]
        tot += vals[0]
        names = [fn(f, l) for f, l in ch if f == 'sqp_warp.cuh']
        top = [n for n in names if n in keys]
        ko = ' > '.join(reversed(top[-3:])) if top else '?'
        ki = fn(*ch[0]) if ch else '?'
        for d, k in ((outer, ko), (inner, ki)):
            a = d[k]
            for i, v in enumerate(vals): a[i] += v
    for name, d in (("outer path", outer), ("innermost function", inner)):
        print(f"{name:44s} {'samp%':>6s} {'noinst%':>7s} {'longsb%':>7s} {'wait%':>6s} {'shortsb%':>8s} {'exec(M)':>8s} {'instrs':>6s}")
        for k, a in sorted(d.items(), key=lambda x: -x[1][0])[:18]:
            print(f"{k:44s} {100*a[0]/tot:6.2f} {100*a[1]/tot:7.2f} {100*a[3]/tot:7.2f} {100*a[4]/tot:6.2f} {100*a[5]/tot:8.2f} {a[2]/1e6:8.1f} {a[7]:6d}")
        print()
    print("total samples", tot)

if __name__ == '__main__':
    main()
