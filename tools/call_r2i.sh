mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2i_pytest.log
timeout 300 python tools/gpu_gemm_timeline2.py 2>&1 | grep -v Warn | tee gpurun_out/r2i_gemm_timeline.txt
