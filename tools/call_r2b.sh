# round 2, call B: GPU suite after the host refactor (shared weight store, LRU engines, per-device state, fork/join) + a WARM launch list
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -6) | tee gpurun_out/r2b_pytest.log
UPGPT_PRECISION=mixed timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/r2b_launches_warm.csv python tools/prof_hot_path.py > gpurun_out/r2b_prof.log 2>&1
tail -1 gpurun_out/r2b_prof.log
python tools/summarize_hot_path.py gpurun_out/r2b_launches_warm.csv 345 95 | tee gpurun_out/r2b_hot_path_warm.txt | head -30
python tools/dump_program.py gpurun_out/r2b_launches_warm.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/r2b_gemm_shapes_warm.txt; head -70 gpurun_out/r2b_gemm_shapes_warm.txt
