python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 0 > gpurun_out/r02_bench_hess1.json 2> gpurun_out/r02_bench_hess1.err
tail -c 1500 gpurun_out/r02_bench_hess1.json; tail -5 gpurun_out/r02_bench_hess1.err
