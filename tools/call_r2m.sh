mkdir -p gpurun_out
timeout 300 python tools/gpu_gemm_timeline2.py 2>&1 | grep -v Warn | grep -A2 -e "--- GEGLU" | tee gpurun_out/r2m_geglu_timeline.txt
