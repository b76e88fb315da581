# full GPU suite + the default bench (both arms optional): the round-end check
T=${1:-full}
mkdir -p gpurun_out
(timeout ${2:-300} python -m pytest tests -m gpu -x -q -W ignore --durations=12 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 200 python bench.py --steps 3 --warmup 3 ${3:---no-cpu-baseline} > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
