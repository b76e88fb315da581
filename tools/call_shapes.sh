mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/r1r_launches.csv python tools/prof_hot_path.py > gpurun_out/r1r_prof.log 2>&1
tail -1 gpurun_out/r1r_prof.log
python tools/dump_program.py gpurun_out/r1r_launches.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/r1r_gemm_shapes.txt
cat gpurun_out/r1r_gemm_shapes.txt
