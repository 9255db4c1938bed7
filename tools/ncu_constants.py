"""Turn an ncu CSV of ONE bench run into profiles/ncu_constants.json (read by bench.py).

On the GPU box (one command; the run under ncu is never a bench value):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
        --clock-control none --csv --log-file gpurun_out/rXX_step.csv \
        python bench.py --no-e2e --no-cpu --no-prof --no-extras --steps 4 --warmup 3
Here:
    python tools/ncu_constants.py gpurun_out/rXX_step.csv 7 "rXX: <what>"     (7 = steps + warmup of that run)

The constants are stamped with the SHA-256 of the kernel sources (bench.kernel_sources_sha) and the git commit:
bench.py prints roofline.traffic / roofline.issue only while the stamp matches the sources it runs.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PROF_NAME = [("k_bb_lanes", "k_bb_lanes"), ("k_front", "k_front"), ("k_demod_spec", "k_demod_spec"),
             ("k_resample", "k_resample"), ("k_rds_front", "k_rds_front"), ("k_rotfir_lanes<1,", "k_rotfir_lp29"),
             ("k_rotfir_lanes<2,", "k_rotfir_rdslp"), ("k_rotfir_lanes<0,", "k_rotfir_rdsmf"), ("k_rds_pll", "k_rds_pll"),
             ("k_audio_tail", "k_audio_tail"), ("k_rds_slice", "k_rds_slice"), ("k_demod_repair", "k_demod_fix"),
             ("k_demod_fix", "k_demod_fix"), ("k_if_level", "k_if_level"), ("k_osc", "k_osc"), ("k_res_taps", "k_res_taps"),
             ("k_tails", "k_tails")]


def main():
    path, blocks, what = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = [l for l in open(path) if l.startswith('"')]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    launches = collections.Counter()
    for r in csv.DictReader(rows):
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "inst": 1.0}.get(unit, 1.0)
        per[name][r["Metric Name"]] += v * scale
        if r["Metric Name"] == "gpu__time_duration.sum":
            launches[name] += 1
    out = {"dram_bytes_per_launch": {}, "warp_instructions_per_launch": {}, "isolated_ms_per_launch": {}}
    tot_b = tot_i = tot_ms = 0.0
    for name, m in per.items():
        key = next((p for frag, p in PROF_NAME if frag in name), None)
        if key is None or "synth" in name or name.startswith("void at::"):
            continue
        n = launches[name]
        b = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        tot_b += b / blocks
        tot_i += m["smsp__inst_executed.sum"] / blocks
        tot_ms += m["gpu__time_duration.sum"] / blocks
        out["dram_bytes_per_launch"][key] = out["dram_bytes_per_launch"].get(key, 0.0) + b / blocks
        out["warp_instructions_per_launch"][key] = out["warp_instructions_per_launch"].get(key, 0.0) + m["smsp__inst_executed.sum"] / blocks
        out["isolated_ms_per_launch"][key] = out["isolated_ms_per_launch"].get(key, 0.0) + m["gpu__time_duration.sum"] / blocks
        del n
    out["dram_bytes_per_step"] = tot_b
    out["warp_instructions_per_step"] = tot_i
    out["isolated_ms_per_step"] = tot_ms
    out["kernel_sources_sha"] = bench.kernel_sources_sha()
    out["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    out["capture"] = what
    out["blocks_in_capture"] = blocks
    with open(os.path.join(ROOT, "profiles", "ncu_constants.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({k: out[k] for k in ("dram_bytes_per_step", "warp_instructions_per_step", "isolated_ms_per_step")}))
    for k, v in sorted(out["isolated_ms_per_launch"].items(), key=lambda kv: -kv[1]):
        print(f"{k:<18s} {v:8.4f} ms  {out['warp_instructions_per_launch'][k] / 1e6:8.1f} M instr  {out['dram_bytes_per_launch'][k] / 1e6:8.1f} MB")


if __name__ == "__main__":
    main()
