# round-end check: full GPU suite, smoke(), the default bench line (what the driver runs)
T=${1:-final}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
(timeout 120 python __graft_entry__.py --smoke 2>&1 | grep -v Warn | tail -3) > gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_smoke.log
timeout 240 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep '^{' gpurun_out/${T}_bench.json | cut -c1-700; tail -2 gpurun_out/${T}_bench.err
