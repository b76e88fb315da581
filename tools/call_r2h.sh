mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -8) | tee gpurun_out/r2h_pytest.log
timeout 600 python bench.py --no-eager-baseline > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; grep '^{' gpurun_out/r2h_bench_c2.json | cut -c1-2300; tail -3 gpurun_out/r2h_bench_c2.err
cat gpurun_out/parity_b8.json
