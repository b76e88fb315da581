T=${1:-final}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -W ignore 2>&1 | tail -4) > gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 python tools/gpu_breakdown.py > gpurun_out/${T}_breakdown.json 2> /dev/null
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_hot_path.py > gpurun_out/${T}_prof.log 2>&1
python tools/dump_program.py gpurun_out/${T}_launches.csv 2>&1 | grep -v "Warn\|Diffusion\|Autoenc" > gpurun_out/${T}_gemm_shapes.txt
X3=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o gpurun_out/${T}_conv224_x3 python tools/prof_gemm_one.py conv224 > gpurun_out/${T}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_prep_fused -s 4 -c 1 -f -o gpurun_out/${T}_gn_fused python tools/prof_misc_one.py gn_fused > gpurun_out/${T}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 4 -c 1 -f -o gpurun_out/${T}_attn_self python tools/prof_misc_one.py attn_self > gpurun_out/${T}_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 4 -c 1 -f -o gpurun_out/${T}_attn_cross python tools/prof_misc_one.py attn_cross > gpurun_out/${T}_ncu5.log 2>&1
timeout 200 python tools/gpu_hbm_kernels.py > gpurun_out/${T}_hbm.txt 2>&1
cat gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_prof.log; cat gpurun_out/${T}_bench_ref.json | cut -c1-300; cat gpurun_out/${T}_bench.json
