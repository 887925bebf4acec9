#!/usr/bin/env bash
# Round-end evidence pass (text outputs only; gpurun_out is capped at 64 MiB).
set -u
out=gpurun_out/final
mkdir -p "$out"
python -m pytest tests/test_model_gpu.py tests/test_full_size_gpu.py -q 2>&1 | tail -3 > "$out/model_tests.txt"
python scripts/attn_bench.py > "$out/attn_bench_tcgen05.txt" 2>&1
UC2_ATTN_TCGEN05=0 python scripts/attn_bench.py > "$out/attn_bench_mma_sync.txt" 2>&1
python scripts/attn_fwd_shapes.py > "$out/attn_fwd_shapes.txt" 2>&1
python scripts/gemm_bench.py --tokens 10240 > "$out/gemm_bench_10240.txt" 2>&1
python scripts/gemm_bench.py --tokens 19200 > "$out/gemm_bench_19200.txt" 2>&1
timeout 600 ncu --set full --clock-control none -k regex:attention_bwd_tc_kernel -s 12 -c 1 -f -o /tmp/attn_bwd_step \
    python bench.py --primary-only --no-cpu-baseline --workload itm --steps 2 --warmup 3 > "$out/ncu_attn_bwd.log" 2>&1
python scripts/ncu_summary.py /tmp/attn_bwd_step.ncu-rep > "$out/attn_bwd_step_ncu.txt" 2>&1
bash scripts/sanitize.sh > "$out/sanitize_stdout.txt" 2>&1
cp gpurun_out/sanitize/summary.txt "$out/sanitize_summary.txt" 2>/dev/null
rm -f gpurun_out/sanitize/*.log
ls -la "$out"
