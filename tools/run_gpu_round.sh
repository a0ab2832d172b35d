python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/gpu_probe.py time > gpurun_out/time.log 2>&1
python bench.py --steps 3 --no-cpu > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
