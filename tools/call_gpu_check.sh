# usage: bash tools/call_gpu_check.sh <tag>   -- GPU tests + phase breakdown + GEMM timeline + ncu launch list of one eager pass
T=${1:-chk}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/${T}_pytest.log
timeout 600 python tools/gpu_breakdown.py > gpurun_out/${T}_breakdown.json 2> gpurun_out/${T}_breakdown.err
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_gemm|attention_kernel|prep_kernel|gn_stats|gn_prep_fused|layernorm|linear_small|conv_small|timestep_emb|softmax_rows" --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_hot_path.py > gpurun_out/${T}_prof.log 2>&1
cat gpurun_out/${T}_pytest.log gpurun_out/${T}_breakdown.json; tail -5 gpurun_out/${T}_breakdown.err; tail -2 gpurun_out/${T}_prof.log
