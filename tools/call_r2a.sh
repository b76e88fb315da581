# round 2, call A: baseline sanity + scheduling / precision-plan probes (no kernel changes)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
P="timeout 300 python tools/gpu_probe_plan.py"
( $P; UPGPT_PDL=0 $P; UPGPT_PAR_SKIP=1 $P; UPGPT_TF_PLANES=x1 $P; UPGPT_PAR_SKIP=1 UPGPT_TF_PLANES=x1 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2a_probe.jsonl
(timeout 600 python -m pytest tests/test_gpu_hotpath.py -x -q -W ignore -k "b8 or full_size" -s 2>&1 | tail -8) | tee gpurun_out/r2a_pytest_b8.log
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2a_pytest.log
