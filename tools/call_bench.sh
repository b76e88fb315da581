T=${1:-bench}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -5) > gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
cat gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench_ref.json; cat gpurun_out/${T}_bench.json
