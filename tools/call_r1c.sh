mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r1c_pytest.log
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/r1c_timeline.txt 2>&1
cat gpurun_out/r1c_pytest.log
