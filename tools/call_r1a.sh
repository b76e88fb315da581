mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r1a_pytest.log
timeout 600 python tools/gpu_breakdown.py > gpurun_out/r1a_breakdown.json 2> gpurun_out/r1a_breakdown.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upgpt --csv --log-file gpurun_out/r1a_launches.csv python tools/prof_hot_path.py > gpurun_out/r1a_prof.log 2>&1
KX=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o gpurun_out/r1a_conv224_x3 python tools/prof_gemm_one.py conv224 > gpurun_out/r1a_ncu_full.log 2>&1
timeout 200 python tools/gpu_gemm_timeline.py > gpurun_out/r1a_timeline.txt 2>&1
cat gpurun_out/r1a_pytest.log gpurun_out/r1a_breakdown.json; tail -3 gpurun_out/r1a_prof.log
