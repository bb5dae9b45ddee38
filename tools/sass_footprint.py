#!/usr/bin/env python
"""Static instruction footprint of k_sqp_warp by source function / call site (no GPU needed).

    cuobjdump -xelf all build/k_sqp_warp.o ; nvdisasm -gi k_sqp_warp.sm_100a.cubin > dis.txt ; python tools/sass_footprint.py dis.txt

The SQP kernel is bound by instruction fetch (ncu: stall_no_instruction), so the bytes of SASS a warp walks through per
interior-point iteration matter; this prints where they are.
"""
import bisect, collections, re, sys, os

def chains_of(path, want):
    chains, cur, pending, infun = [], [], [], False
    for line in open(path):
        if line.startswith('.text.'):
            infun = want(line)
        m = re.match(r'\s*//## File "(.*?)", line (\d+)(?: inlined at "(.*?)", line (\d+))?', line)
        if m:
            if not pending: pending = [(os.path.basename(m.group(1)), int(m.group(2)))]
            if m.group(3): pending.append((os.path.basename(m.group(3)), int(m.group(4))))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
            if pending: cur, pending = pending, []
            if infun: chains.append(cur)
    return chains

def ranges(path):
    starts = []
    for i, l in enumerate(open(path), 1):
        m = re.search(r'MPCC_(?:HDNI|HDN|HD|D)\b[^;(]*?\b(\w+)\s*\(', l)
        if m and not l.strip().startswith('//'): starts.append((i, m.group(1)))
    return starts

def main():
    dis = sys.argv[1]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mpcc_manipulator_b200', 'csrc')
    R = {f: ranges(os.path.join(root, f)) for f in ['sqp_warp.cuh', 'dev_stage.cuh', 'dev_track.cuh', 'dev_panda.cuh', 'dev_qp.cuh']}
    def fn(f, l):
        if f not in R: return f
        st = R[f]; i = bisect.bisect_right([s[0] for s in st], l) - 1
        return st[i][1] if i >= 0 else '?'
    chains = chains_of(dis, lambda line: 'k_sqp_warp' in line and 'solve_ocp' not in line)
    inner, outer = collections.Counter(), collections.Counter()
    for ch in chains:
        if not ch: inner['?'] += 1; continue
        inner[fn(*ch[0])] += 1
        names = [fn(f, l) for f, l in ch if f == 'sqp_warp.cuh']
        top = [n for n in names if n in ('solve', 'factor', 'solve_step', 'gradient', 'ineq_steps', 'eval_horizon', 'run', 'issue_tile', 'qp_box_infeasible', 'gather_point')]
        outer[tuple(reversed(top[-3:])) if top else ('?',)] += 1
    tot = len(chains)
    print(f"k_sqp_warp: {tot} instructions = {tot * 16 / 1024:.0f} KB")
    print("by innermost function:"); [print(f"  {k:28s} {v:6d} {v * 16 / 1024:7.1f} KB") for k, v in inner.most_common(16)]
    print("by outer path:"); [print(f"  {' > '.join(k):44s} {v:6d} {v * 16 / 1024:7.1f} KB") for k, v in outer.most_common(16)]

if __name__ == '__main__':
    main()
