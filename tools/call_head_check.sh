# lean check of the sampler / GEGLU changes + a short bench (budget: ~4 min of box time)
T=${1:-head}
mkdir -p gpurun_out
(timeout 280 python -m pytest tests -m gpu -x -q -W ignore --durations=8 -k "ddim_sampler or ddpm_ancestral or parity_mode or end_to_end or geglu or plms" 2>&1 | tail -25) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 170 python bench.py --steps 3 --warmup 3 --no-fast-mode --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
