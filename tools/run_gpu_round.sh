timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py ab > gpurun_out/ab.log 2>&1
timeout 300 python bench.py --steps 3 --no-cpu --no-e2e > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err
