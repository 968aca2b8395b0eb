python -m pytest tests/test_trainstep_gpu.py tests/test_reference_drivers_gpu.py tests/test_ops_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/t67.log
python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench67.json 2> gpurun_out/bench67.err
cat gpurun_out/t67.log; python -c "
import json
for f in ('bench67',):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'))
"
