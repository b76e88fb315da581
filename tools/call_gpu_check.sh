mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r1k_pytest.log
timeout 600 python tools/gpu_breakdown.py > gpurun_out/r1k_breakdown.json 2> gpurun_out/r1k_breakdown.err
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/r1k_timeline.txt 2>&1
cat gpurun_out/r1k_pytest.log gpurun_out/r1k_breakdown.json; tail -5 gpurun_out/r1k_breakdown.err
