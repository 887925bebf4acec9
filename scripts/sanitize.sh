#!/usr/bin/env bash
# compute-sanitizer memcheck + racecheck + synccheck over the kernel unit tests at their smallest shapes (the tools
# slow kernels down 10-100x).  Logs are reduced to their summaries:  gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
set -u
out=gpurun_out/sanitize
mkdir -p "$out"
run() {   # tool, name, pytest selection...
    local tool=$1 name=$2; shift 2
    echo "=== $tool $name" | tee -a "$out/summary.txt"
    timeout 600 compute-sanitizer --tool "$tool" --print-limit 20 python -m pytest -x -q "$@" > "$out/${tool}_${name}.log" 2>&1
    echo "exit $?" | tee -a "$out/summary.txt"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error" "$out/${tool}_${name}.log" | tail -n 6 | tee -a "$out/summary.txt"
}
for tool in memcheck racecheck synccheck; do
    run $tool attention tests/test_attention_tc_gpu.py -k "matches_reference and (2-16 or 5-33 or 3-161)"
    run $tool attention_dropout tests/test_attention_tc_gpu.py -k "dropout and 5-33"
    run $tool ce tests/test_ce_fused_gpu.py -k "5-33 or 129-1000"
done
run memcheck optim tests/test_optim_gpu.py -k "deferred or bit_level"
run memcheck gemm tests/test_gemm_gpu.py -k "not big" 
