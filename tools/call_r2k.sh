mkdir -p gpurun_out
timeout 300 python tools/gpu_gemm_timeline2.py 2>&1 | grep -v Warn | grep -A1 -e "--- attn.out L0\|--- q L0\|--- ff2 L0" | tee gpurun_out/r2k_gemm_timeline.txt
P="timeout 300 python tools/gpu_probe_plan.py"
( UPGPT_GN_VERBOSE=1 UPGPT_CALIBRATE=1 $P; UPGPT_CALIBRATE=1 UPGPT_GN_FORCE16=1 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2k_probe.jsonl
