"""Summarise an .ncu-rep (and optionally a launch-list csv) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [gpurun_out/launches.csv] > profiles/rN_<kernel>.txt
    python tools/ncu_summary.py --json <envs> <summary-file-name> gpurun_out/prof.ncu-rep   # updates profiles/current.json
The second form records the executed FLOP and the DRAM bytes of the k_step_physics launch in profiles/current.json,
which bench.py reads for `fp32` and `roofline.traffic` (so those always describe the build that is benchmarked).
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_"


def write_json(envs, source, rep):
    import json
    import os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "current.json")
    cur = json.load(open(path)) if os.path.isfile(path) else {}

    def val(r, name):
        i = H.index(name)
        x = float(r[i])
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units[i], 1.0)

    for r in data:
        for kern in ("k_step_physics", "k_post_fused"):
            if kern in r[ki]:
                cyc = val(r, "sm__cycles_elapsed.max")
                fl = sum(val(r, f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") * w
                         for o, w in (("ffma", 2), ("fadd", 1), ("fmul", 1))) * cyc
                cur[kern] = {"envs": int(envs), "flop": fl, "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                             "duration_us_under_ncu": val(r, "gpu__time_duration.sum") / (1.0 if units[H.index("gpu__time_duration.sum")] == "us" else 1e3),
                             "source": source}
    json.dump(cur, open(path, "w"), indent=1)
    print(json.dumps(cur))


def main():
    if sys.argv[1] == "--json":
        return write_json(sys.argv[2], sys.argv[3], sys.argv[4])
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    print(f"# ncu summary of {rep} ({len(data)} launches captured; --set full --clock-control none)")
    for r in data:
        print(f"\n## {r[ki][:100]}")
        for i, h in enumerate(H):
            if h in KEYS:
                print(f"{h:90s} {r[i]:>16s} {units[i]}")
        st = [(float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]) for i, h in enumerate(H)
              if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i]]
        print("stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
        try:
            cyc = float(r[H.index("sm__cycles_elapsed.max")])
            fl = sum(float(r[H.index(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed")]) * w
                     for o, w in (("ffma", 2), ("fadd", 1), ("fmul", 1))) * cyc
            print(f"executed FP32 FLOP this launch (ffma*2 + fadd + fmul): {fl:.4e}")
        except Exception:  # noqa: BLE001
            pass
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        H = rows[hdr]
        ki, vi = H.index("Kernel Name"), H.index("Metric Value")
        agg = collections.defaultdict(list)
        for r in rows[hdr + 1:]:
            if len(r) > vi:
                try:
                    agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
                except ValueError:
                    pass
        tot = sum(sum(v) for v in agg.values())
        print(f"\n# launch list {sys.argv[2]} (gpu__time_duration.sum, cold-cache and serialised: compare shares)")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            print(f"{k:72s} n={len(v):4d} mean={sum(v) / len(v) / 1000:9.2f} us share={sum(v) / tot * 100:5.1f}%")
        # the two launches of one fused env step (dyros_task_step), by their mean durations (k_crossenv in the list
        # belongs to the staged steps bench.py runs to time k_step_physics on its own)
        step = {n: [sum(v) / len(v) for k, v in agg.items() if n in k] for n in ("k_step_physics", "k_post_fused")}
        if all(step.values()):
            t = sum(v[0] for v in step.values())
            print("# one fused env step = " + " + ".join(f"{n} {v[0] / 1000:.1f} us ({v[0] / t * 100:.0f} %)" for n, v in step.items())
                  + f" = {t / 1000:.1f} us of kernel time (k_ffma_peak: the FP32-peak microbenchmark bench.py runs outside the timed region)")


if __name__ == "__main__":
    main()
