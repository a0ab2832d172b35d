timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
SBIR_K1_PAIR=2 timeout 600 python -m pytest tests -m gpu -x -q -k "oracle or full_size or sharded" > gpurun_out/pytest_gpu_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_pair.log
timeout 300 python tools/gpu_probe.py time > gpurun_out/time.log 2>&1
