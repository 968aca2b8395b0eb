python -m pytest tests/test_dp_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/t69_dp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras --no-cpu-baseline --dp-check > gpurun_out/bench69_n2_fused.json 2> gpurun_out/bench69_n2.err
PU_DP_FUSED=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-extras --no-cpu-baseline --dp-check > gpurun_out/bench69_n2_nccl.json 2>> gpurun_out/bench69_n2.err
cat gpurun_out/t69_dp.log; python -c "
import json
for f in ('bench69_n2_fused','bench69_n2_nccl'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))
"
tail -3 gpurun_out/bench69_n2.err
