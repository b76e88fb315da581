T=${1:-prof}
mkdir -p gpurun_out
X3=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o gpurun_out/${T}_conv224_x3 python tools/prof_gemm_one.py conv224 > gpurun_out/${T}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o gpurun_out/${T}_conv224_fp16 python tools/prof_gemm_one.py conv224 > gpurun_out/${T}_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_prep_fused -s 4 -c 1 -f -o gpurun_out/${T}_gn_fused python tools/prof_misc_one.py gn_fused > gpurun_out/${T}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 4 -c 1 -f -o gpurun_out/${T}_attn_self python tools/prof_misc_one.py attn_self > gpurun_out/${T}_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 4 -c 1 -f -o gpurun_out/${T}_attn_cross python tools/prof_misc_one.py attn_cross > gpurun_out/${T}_ncu5.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_ncu*.log; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
