mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -W ignore -k "attention or folded or gemm_plain" 2>&1 | tail -15) | tee gpurun_out/r2d_pytest_ops.log
(UPGPT_LN_FOLD=0 timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2d_pytest_nofold.log
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2d_pytest.log
P="timeout 300 python tools/gpu_probe_plan.py"
( UPGPT_LN_FOLD=0 $P; UPGPT_LN_FOLD=0 UPGPT_ATTN_D32=0 $P; UPGPT_LN_FOLD=0 UPGPT_ATTN_NO_COMPACT=1 $P; $P; UPGPT_TF_PLANES=x1 UPGPT_PAR_SKIP=1 $P ) 2>&1 | grep -v Warn | tee gpurun_out/r2d_probe.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/r2d_launches_warm.csv python tools/prof_hot_path.py > gpurun_out/r2d_prof.log 2>&1
tail -1 gpurun_out/r2d_prof.log
python tools/summarize_hot_path.py gpurun_out/r2d_launches_warm.csv 297 95 | tee gpurun_out/r2d_hot_path_warm.txt | head -16
python tools/dump_program.py gpurun_out/r2d_launches_warm.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/r2d_gemm_shapes_warm.txt; head -60 gpurun_out/r2d_gemm_shapes_warm.txt
