python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/gpu_probe.py time > gpurun_out/time.log 2>&1
python tools/ncu_case.py 12500 75000 2048 float32 10 rank > gpurun_out/ncu_plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dist_topk -s 1 -c 1 -o gpurun_out/k1_tf32_cfg3k10 python tools/ncu_case.py 12500 75000 2048 float32 10 rank > gpurun_out/ncu_a.log 2>&1
python tools/ncu_case.py 20000 1000000 512 bfloat16 10 rank > gpurun_out/ncu_plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dist_topk -s 1 -c 1 -o gpurun_out/k1_bf16_512 python tools/ncu_case.py 20000 1000000 512 bfloat16 10 rank > gpurun_out/ncu_b.log 2>&1
python tools/ncu_case.py 12500 75000 2048 float32 10 rank 3 > gpurun_out/ncu_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3k10.csv python tools/ncu_case.py 12500 75000 2048 float32 10 rank 3 > gpurun_out/ncu_c.log 2>&1
