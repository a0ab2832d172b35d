timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_probe.py time > gpurun_out/time.log 2>&1
timeout 300 python tools/e2e_cfg5.py --gallery 20000 --queries 2000 > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err
