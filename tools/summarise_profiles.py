#!/usr/bin/env python
"""Regenerate the tracked ncu summaries under profiles/ from the scratch captures in gpurun_out/ (tools/gpu_profile_round.sh).

    python tools/summarise_profiles.py            # writes profiles/r1_launches.md, r1_ncu_full_kmlp_ksqp.md, dram_traffic.json
"""
import collections, csv, io, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

CYCLE = ("k_prologue", "k_kin", "k_mlp", "k_order", "k_sqp_warp", "k_sim_step")

def launches():
    text = (OUT / "launches_r1_final.csv").read_text()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        a = agg.setdefault(r["Kernel Name"], [0, 0.0]); a[0] += 1; a[1] += ms
    per_cycle = {k: v for k, v in agg.items() if any(c in k for c in CYCLE)}
    n_cycles = max(v[0] for k, v in per_cycle.items() if "k_order" in k)
    def ms_per_cycle(k, v):  # k_kin / k_mlp also run once outside the cycles (probes): use their mean; k_sqp_warp runs twice per cycle
        return v[1] / v[0] if ("k_kin" in k or "k_mlp" in k) else v[1] / n_cycles
    cyc_ms = sum(ms_per_cycle(k, v) for k, v in per_cycle.items())
    lines = ["# Round 1 - ncu launch list (final build of the round)", "",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (B = 4096, N = 20), run after the same command exited 0 without ncu.",
             "Per-launch times are cold-cache and serialised: compare SHARES with the live CUDA-event numbers of bench.py (profiles/bench_r1_single_gpu.json: `kernels_ms`, `roofline.kernel_share_of_step`).",
             "`k_sqp_warp` is launched twice per cycle (exclusive-SM launch for predicted stragglers, often empty, + main launch): its row is per cycle.", "",
             "| kernel | launches | ms per cycle (mean per launch outside the cycle) | share of cycle |", "|---|---|---|---|"]
    for k, (n, tot) in agg.items():
        if k in per_cycle: lines.append(f"| `{k[:70]}` | {n} | {ms_per_cycle(k, (n, tot)):.3f} | {100 * ms_per_cycle(k, (n, tot)) / cyc_ms:.2f}% |")
        else: lines.append(f"| `{k[:70]}` | {n} | {tot / n:.3f} | (outside the cycle) |")
    lines += ["", f"Kernels of one control cycle: {cyc_ms:.2f} ms under ncu ({n_cycles} cycles captured).  The first cycles after `reset()` are cold starts (every instance runs several SQP iterations), which is why `k_sqp_warp`'s time here is above the closed-loop average of bench.py."]
    try:
        live = json.loads((OUT / "plain.log").read_text().strip().splitlines()[-1])
        lines += ["", "Live CUDA-event shares of the SAME command run without ncu just before (timed cycles only): "
                  + ", ".join(f"{k} {100 * v:.1f}%" for k, v in live["roofline"]["kernel_share_of_step"].items())
                  + f" of {live['ms_per_step']:.2f} ms per step.  (The ncu list also contains the untimed warm-up cycles 0-2, which are cold starts, and serialises the two SQP launches.)"]
    except Exception:
        pass
    (PROF / "r1_launches.md").write_text("\n".join(lines) + "\n")

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]

def full(reading):
    rep = OUT / "prof_r1_final.ncu-rep"
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    lines = ["# Round 1 - ncu `--set full` capture of the two main kernels (one launch each, final build of the round)", "",
             'Command: `ncu --set full --clock-control none --import-source on -k regex:"k_mlp|k_sqp_warp" -s 6 -c 3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (B = 4096, N = 20, one B200), run after the same command exited 0 without ncu.',
             "Times under ncu are cold-cache and serialised (not bench values). Raw report: gpurun_out/prof_r1_final.ncu-rep (scratch).", ""]
    traffic = {}
    # the SQP kernel is launched twice per cycle (exclusive-SM launch, often empty, + main launch; two builds): keep the longest capture
    best = {}
    for r in rows[2:]:
        kn = "k_sqp_warp" if "k_sqp_warp" in r[name_col] else "k_mlp" if "k_mlp" in r[name_col] else None
        if not kn: continue
        dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms"), 1.0)
        if kn not in best or dur > best[kn][0]: best[kn] = (dur, r)
    for kn, (dur, r) in best.items():
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        lines += [f"(captured launch: `{r[name_col][:60]}`)", ""]
        lines += [f"## `{kn}`", "", "| metric | value | unit |", "|---|---|---|"]
        for m in KEEP:
            if m in d: lines.append(f"| {m} | {d[m]} | {u[m]} |")
        t = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        traffic[kn] = t
        lines += ["", f"DRAM traffic per launch (read + write): {t / 1e9:.3f} GB", ""]
    lines += ["## Reading", ""] + reading
    (PROF / "r1_ncu_full_kmlp_ksqp.md").write_text("\n".join(lines) + "\n")
    (PROF / "dram_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

READING = (ROOT / "profiles" / "r1_reading.md")

if __name__ == "__main__":
    launches()
    full(READING.read_text().splitlines() if READING.exists() else [])
    print("profiles/ refreshed")
