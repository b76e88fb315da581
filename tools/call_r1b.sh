mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r1b_pytest.log
timeout 600 python tools/gpu_breakdown.py > gpurun_out/r1b_breakdown.json 2> gpurun_out/r1b_breakdown.err
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/r1b_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/r1b_launches.csv python tools/prof_hot_path.py > gpurun_out/r1b_prof.log 2>&1
cat gpurun_out/r1b_pytest.log gpurun_out/r1b_breakdown.json; tail -3 gpurun_out/r1b_prof.log; tail -5 gpurun_out/r1b_breakdown.err
