mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -5) | tee gpurun_out/r2l_pytest.log
bash tools/call_evidence.sh r02
python tools/summarize_hot_path.py gpurun_out/r02_launches.csv 297 95 | tee gpurun_out/r02_hot_path_launches.txt | head -14
python tools/dump_program.py gpurun_out/r02_launches.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/r02_gemm_time_by_shape.txt
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q -W ignore -k "gemm_plain or conv3x3 or geglu or folded or groupnorm or layernorm or attention_head_pairs or small_kernels" 2>&1 | tail -12) | tee gpurun_out/r02_sanitizer_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q -W ignore -k "gemm_plain or folded or groupnorm or attention_head_pairs" 2>&1 | tail -12) | tee gpurun_out/r02_sanitizer_racecheck.log
