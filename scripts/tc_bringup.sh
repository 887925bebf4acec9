#!/usr/bin/env bash
# First GPU call of the next round: bring up the experimental tcgen05 attention kernels (csrc/attention_tc.cu).
# Every stage runs in its own process under `timeout` (a pipeline bug traps after ~4 s and poisons that process's
# CUDA context only); logs go to gpurun_out/tc_bringup/.  Usage on the GPU box:
#     gpurun --timeout 900 -- 'bash scripts/tc_bringup.sh'
set -u
out=gpurun_out/tc_bringup
mkdir -p "$out"
export UC2_TEST_EXPERIMENTAL=1
run() {   # name, command...
    local name=$1; shift
    echo "=== $name" | tee -a "$out/summary.txt"
    timeout 300 "$@" > "$out/$name.log" 2>&1
    echo "exit $?" | tee -a "$out/summary.txt"
    tail -n 6 "$out/$name.log" | tee -a "$out/summary.txt"
}
run fwd_small   python -m pytest tests/test_attention_tc_gpu.py -x -q -k "forward_matches and (2-16 or 5-33)"
run fwd_all     python -m pytest tests/test_attention_tc_gpu.py -x -q -k "forward or switch or rejects"
run bwd_small   python -m pytest tests/test_attention_tc_gpu.py -x -q -k "backward_matches and (2-16 or 5-33)"
run bwd_all     python -m pytest tests/test_attention_tc_gpu.py -x -q -k "backward"
run end_to_end  python -m pytest tests/test_attention_tc_gpu.py -x -q -k "end_to_end"
# the backward with four instead of two element-wise warps per TMEM lane quarter (16 warps; tuning knob)
UC2_ATTN_TC_BWD_SPLIT=4 run bwd_all_split4 python -m pytest tests/test_attention_tc_gpu.py -x -q -k "backward"
# memcheck of one small forward + backward if anything above failed
if grep -q "exit [1-9]" "$out/summary.txt"; then
    run sanitizer compute-sanitizer --tool memcheck python -m pytest tests/test_attention_tc_gpu.py -x -q -k "2-16"
fi
# timing, default kernels vs the tc kernels
run bench_mma_sync python scripts/attn_bench.py
UC2_ATTN_TCGEN05=1 run bench_tcgen05 python scripts/attn_bench.py
UC2_ATTN_TCGEN05=1 UC2_ATTN_TC_BWD_SPLIT=4 run bench_tcgen05_split4 python scripts/attn_bench.py
# the whole ITM step both ways (the second only means something if every stage above exited 0)
run bench_itm_default python bench.py --steps 20 --warmup 5 --no-cpu-baseline
UC2_ATTN_TCGEN05=1 run bench_itm_tcgen05 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
# one full ncu capture of each tc kernel (never a timing source)
UC2_ATTN_TCGEN05=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:attention_.*_tc_kernel -c 2 -f -o "$out/attn_tc" python scripts/attn_bench.py > "$out/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$out/summary.txt"
