python -m pytest tests -q -x -m gpu 2>&1 | tail -4 > gpurun_out/t65.log
python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench65.json 2> gpurun_out/bench65.err
PU_POOL_CODE=0 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench65_nocode.json 2>> gpurun_out/bench65.err
cat gpurun_out/t65.log; python -c "
import json
for f in ('bench65','bench65_nocode'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'))
"
