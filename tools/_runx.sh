python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_n.log 2>&1
tail -3 gpurun_out/r2_pytest_n.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
