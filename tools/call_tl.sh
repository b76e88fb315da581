mkdir -p gpurun_out
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/r1f_timeline.txt 2>&1
