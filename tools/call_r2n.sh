mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -W ignore 2>&1 | tail -5) | tee gpurun_out/r2n_pytest_ops.log
(timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -5) | tee gpurun_out/r2n_pytest.log
timeout 300 python tools/gpu_gemm_timeline2.py 2>&1 | grep -v Warn | grep -e "--- GEGLU" | tee gpurun_out/r2n_geglu_timeline.txt
P="timeout 300 python tools/gpu_probe_plan.py"
( UPGPT_CALIBRATE=1 $P; UPGPT_CALIBRATE=1 UPGPT_GEMM_NO_EPI2=1 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2n_probe.jsonl
