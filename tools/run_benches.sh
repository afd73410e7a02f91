set -x
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > gpurun_out/r02_final_pytest.log
python bench.py --config 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err
python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1_n1.json 2> /dev/null
python bench.py --config 3 --steps 2 --warmup 3 --timesteps 200 > gpurun_out/r02_bench_c3_n1_t200.json 2> /dev/null
python bench.py --config 3cfg8 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3cfg8_n1.json 2> /dev/null
python bench.py --config 4 --steps 2 --warmup 3 > gpurun_out/r02_bench_c4_n1.json 2> /dev/null
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2> /dev/null
python bench.py --impl reference --config 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_reference.json 2> /dev/null
