mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2c_pytest.log
timeout 600 python tools/gpu_probe_batch.py 2>&1 | grep -v Warn | tee gpurun_out/r2c_batch_probe.json
P="timeout 300 python tools/gpu_probe_plan.py"
( UPGPT_TF_PLANES=x1 $P; UPGPT_TF_PLANES=x1 UPGPT_GEMM_2CTA=1 $P; UPGPT_PRECISION=fp16 $P; UPGPT_PRECISION=fp16 UPGPT_GEMM_2CTA=1 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2c_probe.jsonl
