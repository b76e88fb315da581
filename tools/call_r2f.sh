mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -W ignore -k "folded or geglu" 2>&1 | tail -8) | tee gpurun_out/r2f_pytest_ops.log
(timeout 1500 python -m pytest tests -m gpu -x -q -W ignore --durations=8 2>&1 | tail -22) | tee gpurun_out/r2f_pytest.log
P="timeout 300 python tools/gpu_probe_plan.py"
( $P; UPGPT_LN_FOLD=0 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2f_probe.jsonl
