for cap in 8 4 16 28 64; do echo "[PU_STEM_CAP=$cap]"; PU_STEM_CAP=$cap python scripts/conv_probe.py tf32 1 0 8 128 64; done > gpurun_out/t71.log 2>&1
cat gpurun_out/t71.log
