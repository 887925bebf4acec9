echo "== auto"; python scripts/gemm_bench.py --reps 10
echo "== c2 256"; python scripts/gemm_bench.py --reps 10 --ctas 2 --block-n 256
echo "== c1 256"; python scripts/gemm_bench.py --reps 10 --ctas 1 --block-n 256
echo "== c1 128"; python scripts/gemm_bench.py --reps 10 --ctas 1 --block-n 128
