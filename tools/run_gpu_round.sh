timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_probe.py bw > gpurun_out/bw.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"l2_normalize|row_norm|topk_merge|triplet_rows" -c 40 --csv --log-file gpurun_out/bw_kernels_ncu.csv python tools/gpu_probe.py bw > gpurun_out/ncu_bw.log 2>&1
timeout 300 python bench.py --workload cfg3 --steps 5 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg3k10 --steps 5 --no-cpu > gpurun_out/bench_cfg3k10.json 2> gpurun_out/bench_cfg3k10.err
timeout 300 python bench.py --workload cfg1 --steps 10 --no-cpu > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
