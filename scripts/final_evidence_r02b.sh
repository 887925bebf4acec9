#!/usr/bin/env bash
# Round-2b closing pass on one box: the whole GPU suite, smoke(), the default bench line (as the driver runs it), a short
# reference-arm line, the launch list of the final build, memcheck of the GEMM tests under the dynamic tile order.
set -u
out=gpurun_out/final2
mkdir -p "$out"
( time python -m pytest tests -m gpu -x -q ) > "$out/gpu_tests.txt" 2>&1
tail -n 4 "$out/gpu_tests.txt"
python __graft_entry__.py smoke > "$out/smoke.txt" 2>&1; tail -n 1 "$out/smoke.txt"
( time python bench.py ) > "$out/bench_default.json" 2> "$out/bench_default.err"; tail -n 4 "$out/bench_default.err"
( time python bench.py --impl reference --steps 2 --warmup 1 ) > "$out/bench_reference.json" 2> "$out/bench_reference.err"; tail -n 4 "$out/bench_reference.err"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/launches_pretrain.csv \
    python bench.py --primary-only --no-cpu-baseline --workload pretrain --steps 4 --warmup 3 > "$out/launches_pretrain.log" 2>&1
python scripts/summarize_launches.py /tmp/launches_pretrain.csv 4 > "$out/launches_pretrain.txt"
UC2_GEMM_SCHED=dynamic timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q tests/test_gemm_gpu.py -k "not big" \
    > "$out/memcheck_gemm_dynamic.log" 2>&1
grep -E "ERROR SUMMARY|passed|failed" "$out/memcheck_gemm_dynamic.log" | tail -n 3 | tee "$out/memcheck_gemm_dynamic.txt"
rm -f "$out/memcheck_gemm_dynamic.log"
ls -la "$out"
