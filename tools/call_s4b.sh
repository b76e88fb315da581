T=${1:-s4b}
mkdir -p gpurun_out
(timeout 120 python -m pytest tests/test_gpu_zclip.py -x -q -W ignore 2>&1 | tail -15) > gpurun_out/${T}_pytest_clip.log; cat gpurun_out/${T}_pytest_clip.log
timeout 120 python tools/gpu_clip_timing.py 2>&1 | grep -v Warn | tail -25
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_hot_path.py > gpurun_out/${T}_prof.log 2>&1
tail -1 gpurun_out/${T}_prof.log
python tools/summarize_hot_path.py gpurun_out/${T}_launches.csv 345 95 > gpurun_out/${T}_hot_path.txt 2>&1; head -30 gpurun_out/${T}_hot_path.txt
python tools/dump_program.py gpurun_out/${T}_launches.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/${T}_gemm_shapes.txt; head -5 gpurun_out/${T}_gemm_shapes.txt
