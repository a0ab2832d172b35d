# Evidence pass (run under gpurun, one GPU): each ncu command only after the plain command exited 0.
set -x
mkdir -p gpurun_out
# (a) launch list of the default bench (share of the step per kernel)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_a_plain.json 2> gpurun_out/ncu_a_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg4.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_a.log 2>&1
# (b) full capture of one K1 launch at cfg4 (DRAM traffic, tensor pipe)
timeout 300 python tools/ncu_case.py 100000 10000000 512 bfloat16 10 rank 2 > gpurun_out/ncu_b_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dist_topk -s 1 -c 1 -o gpurun_out/k1_cfg4_r01b -f \
  python tools/ncu_case.py 100000 10000000 512 bfloat16 10 rank 2 > gpurun_out/ncu_b.log 2>&1
# (c) bandwidth kernels: duration + DRAM bytes
timeout 300 python tools/gpu_probe.py bw > gpurun_out/bw.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"topk_merge|l2_normalize|row_norm|triplet|chunk_min" -c 60 --csv --log-file gpurun_out/bw_kernels_ncu.csv \
  python tools/gpu_probe.py bw > gpurun_out/ncu_c.log 2>&1
# (d) launch list of the fp32 top-100 workload
timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_d_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg3.csv \
  python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_d.log 2>&1
