python -m pytest tests/test_ops_gpu.py -q -x -k "plastic_head or trace or conv1x1" 2>&1 | tail -3 > gpurun_out/t64.log
python -m pytest tests/test_trainstep_gpu.py tests/test_models_gpu.py -q -x 2>&1 | tail -3 >> gpurun_out/t64.log
python scripts/head_probe.py 128 64 >> gpurun_out/t64.log 2>&1
python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench64.json 2> gpurun_out/bench64.err
PU_HEAD_FUSED=0 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench64_unfused.json 2>> gpurun_out/bench64.err
cat gpurun_out/t64.log; python -c "
import json
for f in ('bench64','bench64_unfused'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'))
"
