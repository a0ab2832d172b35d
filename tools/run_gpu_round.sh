SBIR_K1_PAIR=1 timeout 300 python bench.py --steps 3 --no-cpu --no-e2e > gpurun_out/bench_cfg4_pair.json 2> gpurun_out/bench_cfg4_pair.err
SBIR_K1_PAIR=0 timeout 300 python bench.py --steps 3 --no-cpu --no-e2e > gpurun_out/bench_cfg4_single.json 2> gpurun_out/bench_cfg4_single.err
SBIR_K1_PAIR=1 timeout 300 python bench.py --steps 3 --no-cpu --no-e2e > gpurun_out/bench_cfg4_pair2.json 2> gpurun_out/bench_cfg4_pair2.err
SBIR_K1_PAIR=0 timeout 300 python bench.py --steps 3 --no-cpu --no-e2e > gpurun_out/bench_cfg4_single2.json 2> gpurun_out/bench_cfg4_single2.err
