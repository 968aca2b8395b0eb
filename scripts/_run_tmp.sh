python -m pytest tests/test_ops_gpu.py tests/test_conv_tc_gpu.py tests/test_models_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/t68.log
python scripts/conv_probe.py tf32 1 0 8 128 64 >> gpurun_out/t68.log 2>&1
python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench68.json 2> gpurun_out/bench68.err
cat gpurun_out/t68.log; python -c "
import json
for f in ('bench68',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'))
"
