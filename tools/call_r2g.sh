mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err; grep '^{' gpurun_out/r2g_bench_c2.json | cut -c1-1500; tail -3 gpurun_out/r2g_bench_c2.err
timeout 600 python bench.py --config c4 --steps 2 --warmup 3 > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; grep '^{' gpurun_out/r2g_bench_c4.json | cut -c1-900; tail -3 gpurun_out/r2g_bench_c4.err
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err; grep '^{' gpurun_out/r2g_bench_c5.json | cut -c1-900; tail -3 gpurun_out/r2g_bench_c5.err
timeout 300 python bench.py --impl reference > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; grep '^{' gpurun_out/r2g_bench_ref.json | cut -c1-600
