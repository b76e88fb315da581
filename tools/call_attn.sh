T=${1:-attn}
mkdir -p gpurun_out
(timeout 200 python -m pytest tests -m gpu -x -q -W ignore -k "flash_attention or parity_mode or ddim_sampler_fused or end_to_end or config5 or full_size_properties or zclip" 2>&1 | tail -15) > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
timeout 170 python bench.py --steps 3 --warmup 3 --no-fast-mode --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'P'
import json
for l in open('gpurun_out/attn_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], {k:(v['us_per_launch'], v['frac']) for k,v in d['roofline_attention'].items()})
P
tail -2 gpurun_out/${T}_bench.err
