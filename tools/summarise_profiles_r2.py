#!/usr/bin/env python
"""Round 2: regenerate the tracked ncu summaries under profiles/ from the scratch captures in gpurun_out/
(tools/gpu_profile_round2.sh).  Writes profiles/r2_launches.md, profiles/r2_ncu_full_steady.md, profiles/dram_traffic.json."""
import collections, csv, io, json, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
SHA = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
CMD = "python bench.py --steps 3 --warmup 30 --no-cpu-baseline --no-secondary   (no settle phase in that build: 33 closed-loop cycles from the cold start)"


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def launches():
    text = (OUT / "r2_launches.csv").read_text()
    rows = list(csv.DictReader(io.StringIO(text[text.index('"ID"'):])))
    per = collections.OrderedDict()
    seq = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = num(r["Metric Value"]); unit = r["Metric Unit"]
        ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v
        name = r["Kernel Name"].split("(")[0].replace("mpcc::", "")
        seq.append((name, ms))
        a = per.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
    # steady state: the last 3 cycles (each cycle: k_prologue ... k_sim_step)
    idx = [i for i, (n, _) in enumerate(seq) if n == "k_prologue"]
    steady = seq[idx[30]:idx[33]]   # cycles 30..32 of the device-resident loop (the e2e pass of bench.py follows and starts cold again)
    cyc = collections.OrderedDict()
    for n, ms in steady:
        cyc[n] = cyc.get(n, 0.0) + ms / 3
    own = {k: v for k, v in cyc.items() if k.startswith("k_")}
    tot = sum(own.values())
    live = json.loads([l for l in (OUT / "r2_plain.log").read_text().splitlines() if l.startswith("{")][-1])
    lines = ["# Round 2 - ncu launch list (steady state)", "",
             f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv {CMD}`, run after the same command exited 0 without ncu (build {SHA}).",
             f"{len(seq)} launches captured, {len(idx)} control cycles; the table is the mean of cycles 30..32 after the cold start (steady state: one SQP iteration per instance).",
             "Per-launch times under ncu are serialised and cold-cache: compare SHARES with the live CUDA-event numbers (right column, same command without ncu).", "",
             "| kernel | ms per cycle (ncu) | share of the cycle's own kernels | live CUDA events, ms (share of step) |", "|---|---|---|---|"]
    lk, ls = live["kernels_ms"], live["roofline"]["kernel_share_of_step"]
    for k, v in cyc.items():
        lv = f"{lk[k]:.3f} ({100 * ls[k]:.1f} %)" if k in lk else ""
        lines.append(f"| `{k}` | {v:.3f} | {100 * v / tot:.2f} % |" + f" {lv} |" if k in own else f"| `{k[:60]}` (torch: L2 flush / copies of the harness) | {v:.3f} | - | |")
    lines += ["", f"Own kernels of one steady-state cycle under ncu: {tot:.2f} ms; live step {live['ms_per_step']:.2f} ms.", "",
              "All launches of the capture:", "", "| kernel | launches | total ms |", "|---|---|---|"]
    for k, (n, t) in per.items():
        lines.append(f"| `{k[:70]}` | {n} | {t:.3f} |")
    (PROF / "r2_launches.md").write_text("\n".join(lines) + "\n")


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
        ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "shared fp64+DMMA pipe active"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 (DFMA) pipe active"),
        ("sm__inst_executed_pipe_fp64.sum", "warp instructions on the fp64 pipe (DFMA etc.)"), ("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "warp instructions DMMA"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct", "issue slots busy"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__occupancy_limit_shared_mem", "CTAs / SM (shared-memory limit)")]


def full(rep, title, note):
    raw = subprocess.run(["ncu", "-i", str(OUT / rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"## {title}", "", note, ""]
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("mpcc::", "")
        out += [f"### `{name}`", "", "| metric | value |", "|---|---|"]
        for m, label in WANT:
            if m in idx and r[idx[m]] not in ("", "n/a"):
                out.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        stalls = sorted([(num(r[i]) or 0.0, h) for h, i in idx.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")], reverse=True)[:6]
        out += ["", "Largest warp-stall reasons (cycles per issued instruction): " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls), ""]
        rd, wr = num(r[idx["dram__bytes_read.sum"]]), num(r[idx["dram__bytes_write.sum"]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        b = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
        dur = r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]]
        traffic[name] = {"dram_bytes_per_launch": b, "kernel_duration_under_ncu": dur, "capture": note, "build": SHA}
    return out, traffic


def main():
    launches()
    a, t1 = full("r2_prof_steady.ncu-rep", "C2 steady state (4096 x N = 20): cycle 31 after the cold start",
                 f"`ncu --set full --clock-control none --import-source on -k regex:\"k_mlp|k_sqp_warp\" -s 93 -c 3 {CMD}`: the 94th..96th matching launches = "
                 "one whole steady-state cycle (k_sqp_warp main launch, k_mlp, and the -- empty -- exclusive-SM launch k_sqp_warp_r255).")
    b, t2 = full("r2_prof_cta.ncu-rep", "C5 latency mode (64 x N = 40): the CTA-per-instance kernel",
                 "`ncu --set full --clock-control none --import-source on -k regex:k_sqp_cta -s 60 -c 1 python bench.py --config c5 --steps 60 --warmup 20 --no-cpu-baseline`")
    (PROF / "r2_ncu_full_steady.md").write_text("# Round 2 - ncu --set full captures\n\nRegenerate with tools/summarise_profiles_r2.py from gpurun_out/*.ncu-rep (tools/gpu_profile_round2.sh).\n\n" + "\n".join(a + b) + "\n")
    t1.update(t2)
    t1.pop("k_sqp_warp_r255", None)
    (PROF / "dram_traffic.json").write_text(json.dumps(t1, indent=1) + "\n")
    print("written:", "profiles/r2_launches.md, profiles/r2_ncu_full_steady.md, profiles/dram_traffic.json")


if __name__ == "__main__":
    main()
