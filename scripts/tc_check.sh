#!/usr/bin/env bash
# Parity + timing of the tcgen05 attention kernels on the GPU box: gpurun --timeout 900 -- 'bash scripts/tc_check.sh'
set -u
out=gpurun_out/tc_check
mkdir -p "$out"
run() {   # name, command...
    local name=$1; shift
    echo "=== $name" | tee -a "$out/summary.txt"
    timeout 300 "$@" > "$out/$name.log" 2>&1
    echo "exit $?" | tee -a "$out/summary.txt"
    tail -n 8 "$out/$name.log" | tee -a "$out/summary.txt"
}
run tests python -m pytest tests/test_attention_tc_gpu.py -x -q
run bench_mma_sync env UC2_ATTN_TCGEN05=0 python scripts/attn_bench.py
run bench_tcgen05 env UC2_ATTN_TCGEN05=1 python scripts/attn_bench.py
if [ "${1:-}" = "ncu" ]; then
    UC2_ATTN_TCGEN05=1 timeout 600 ncu --set full --clock-control none --import-source on \
        -k regex:attention_.*_tc_kernel -c 2 -f -o "$out/attn_tc" python scripts/attn_bench.py > "$out/ncu.log" 2>&1
    echo "ncu exit $?" | tee -a "$out/summary.txt"
fi
