# One GPU round-trip (run under gpurun on one B200): tests, smoke, the default bench, the probes.
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python tools/gpu_probe.py time > gpurun_out/time.log 2>&1
timeout 300 python tools/gpu_probe.py diag > gpurun_out/diag.log 2>&1
timeout 300 python tools/e2e_cfg5.py --gallery 25000 --queries 2000 > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err
