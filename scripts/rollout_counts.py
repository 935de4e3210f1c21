"""profiles/rollout_counts.json from ncu metric CSVs of ONE rollout launch per BASELINE config:
    ncu --metrics <METRICS> --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 --csv \
        --log-file gpurun_out/counts_cfg<i>.csv python scripts/prof_cfg.py <i> 3
    python scripts/rollout_counts.py gpurun_out/counts_cfg{0,1,2,3,4}.csv
(the first 11 rollout launches of prof_cfg.py are the single-instance pipeline_init / env steps)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline_configs import BASELINE, ENV_CFG  # noqa: E402

METRICS = ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum")


def num(v, unit):
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0,
             "us": 1e-3, "ns": 1e-6, "ms": 1.0}.get(unit, 1.0)
    return x * scale


def main():
    path = os.path.join(ROOT, "profiles", "rollout_counts.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for f in sys.argv[1:]:
        ci = int("".join(ch for ch in os.path.basename(f) if ch.isdigit())[-1])
        b = BASELINE[ci]
        lines = [l for l in open(f, newline="") if not l.startswith("==")]
        m, meta = {}, {}
        for r in csv.DictReader(lines):
            m[r["Metric Name"]] = num(r["Metric Value"], r.get("Metric Unit", ""))
            meta = r
        n_frames = int(round(ENV_CFG[b["env"]].get("dt", 0.02) / ENV_CFG[b["env"]].get("timestep", 0.02))) if b["env"] == "allegro_reorient" else 1
        steps = (b["N"] + 1) * (b["Hs"] + 1) * n_frames
        flop = (m["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
                + 2 * m["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"])
        out[b["name"]] = {
            "kernel": meta["Kernel Name"][:48], "grid": meta["Grid Size"], "block": meta["Block Size"],
            "physics_steps_per_launch": steps, "flop_per_launch": flop, "flop_per_physics_step": flop / steps,
            "warp_inst_per_physics_step": m["smsp__inst_executed.sum"] / steps,
            "thread_inst_per_warp_inst": m["smsp__thread_inst_executed.sum"] / m["smsp__inst_executed.sum"],
            "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "ncu_time_ms": m["gpu__time_duration.sum"],
            "source": f"ncu --metrics {METRICS.split(',')[0]},... one launch of scripts/prof_cfg.py {ci} "
                      f"(scripts/rollout_counts.py; round 2, kernel with the peeled Newton loop)"}
    json.dump(out, open(path, "w"), indent=1)
    for k, v in out.items():
        print(k, round(v["warp_inst_per_physics_step"]), round(v["flop_per_physics_step"]), v["ncu_time_ms"])


if __name__ == "__main__":
    main()
