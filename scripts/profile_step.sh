#!/usr/bin/env bash
# Launch list of the pre-training mix (one task cycle = 4 steps) and of the ITM fine-tuning step, plus `ncu --set full`
# captures of the attention kernels and of the GEMM forms inside the running step.  Everything is reduced to text ON
# THE BOX (gpurun_out/ is capped at 64 MiB; .ncu-rep files of a whole step are larger than that).
#   gpurun --timeout 1500 -- "UC2_COMMIT=$(git rev-parse --short HEAD) bash scripts/profile_step.sh"
set -u
out=gpurun_out/profile
tmp=/tmp/uc2_profile
mkdir -p "$out" "$tmp"
B="python bench.py --primary-only --no-cpu-baseline"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$tmp/launches_pretrain.csv" \
    $B --workload pretrain --steps 4 --warmup 3 > "$out/launches_pretrain.log" 2>&1
python scripts/summarize_launches.py "$tmp/launches_pretrain.csv" 4 > "$out/launches_pretrain.txt"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$tmp/launches_itm.csv" \
    $B --workload itm --steps 2 --warmup 3 > "$out/launches_itm.log" 2>&1
python scripts/summarize_launches.py "$tmp/launches_itm.csv" 1 > "$out/launches_itm.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_.*_tc_kernel -s 48 -c 2 -f \
    -o "$out/attn_step" $B --workload itm --steps 2 --warmup 3 > "$out/ncu_attn.log" 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 450 -c 24 -f \
    -o "$tmp/gemm_step" $B --workload pretrain --steps 4 --warmup 3 > "$out/ncu_gemm.log" 2>&1
python scripts/ncu_summary.py "$tmp/gemm_step.ncu-rep" > "$out/gemm_step_ncu.txt"
python - "$out/gemm_step_ncu.txt" "$out/ncu_traffic.json" <<'PY'
import json, subprocess, sys
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n = 0.0, 0
for line in open(sys.argv[1]):
    t = line.split()
    if "gemm_bf16_kernel" in line:
        n += 1
    elif len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and n:
        tot += float(t[1]) * scale.get(t[2], 1.0)
import os
rev = os.environ.get("UC2_COMMIT") or "unknown"     # .git does not travel to the GPU box: pass UC2_COMMIT=$(git rev-parse --short HEAD)
json.dump({"gemm_bytes_per_launch": tot / max(n, 1), "launches": n, "commit": rev,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the gemm_bf16_kernel launches of one ncu --set full "
                   "capture inside the running pre-training step (scripts/profile_step.sh)"}, open(sys.argv[2], "w"))
PY
python scripts/ncu_summary.py "$out/attn_step.ncu-rep" > "$out/attn_step_ncu.txt"
du -sh "$out"; ls -la "$out"
