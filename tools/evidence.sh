# round-2 evidence: ncu --set full captures of the kernels bench.py's rooflines name, launch list of the final tree, sanitizer runs
T=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
X3=1 timeout 300 $NCU -k regex:tc_gemm -s 3 -c 1 -f -o gpurun_out/${T}_conv224_x3 python tools/prof_gemm_one.py conv224 > gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:tc_gemm -s 3 -c 1 -f -o gpurun_out/${T}_conv_deep python tools/prof_gemm_one.py conv_deep >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:tc_gemm -s 3 -c 1 -f -o gpurun_out/${T}_conv_up2 python tools/prof_gemm_one.py up2 >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:tc_gemm -s 3 -c 1 -f -o gpurun_out/${T}_geglu python tools/prof_gemm_one.py geglu >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:gn_group -s 3 -c 1 -f -o gpurun_out/${T}_gn_group python tools/prof_misc_one.py gn_group >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:attention_kernel -s 3 -c 1 -f -o gpurun_out/${T}_attn_self python tools/prof_misc_one.py attn_self >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:attention_kernel -s 3 -c 1 -f -o gpurun_out/${T}_attn_cross python tools/prof_misc_one.py attn_cross >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:gn_prep_fused -s 3 -c 1 -f -o gpurun_out/${T}_gn_fused python tools/prof_misc_one.py gn_fused >> gpurun_out/${T}_ncu.log 2>&1
timeout 300 $NCU -k regex:prep_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prep python tools/prof_misc_one.py prep >> gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py gpurun_out/${T}_conv224_x3.ncu-rep gpurun_out/${T}_conv_deep.ncu-rep gpurun_out/${T}_conv_up2.ncu-rep gpurun_out/${T}_geglu.ncu-rep gpurun_out/${T}_gn_group.ncu-rep gpurun_out/${T}_attn_self.ncu-rep gpurun_out/${T}_attn_cross.ncu-rep gpurun_out/${T}_gn_fused.ncu-rep gpurun_out/${T}_prep.ncu-rep > gpurun_out/${T}_ncu_set_full_summary.txt 2>&1
head -60 gpurun_out/${T}_ncu_set_full_summary.txt
# launch list of the default bench step (cold caches, serialised: shares)
UPGPT_CALIBRATE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|gn_group|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_hot_path.py > gpurun_out/${T}_prof.log 2>&1
tail -1 gpurun_out/${T}_prof.log
