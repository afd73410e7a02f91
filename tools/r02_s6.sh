set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_s6_pytest.log
B=64 python tools/step_profile.py > $O/r02_s6_prof.jsonl 2>/dev/null
python bench.py --config 2 --steps 3 --warmup 3 > $O/r02_s6_bench_c2.json 2> $O/r02_s6_bench_c2.err
python bench.py --config 1 --steps 5 --warmup 3 > $O/r02_s6_bench_c1.json 2> /dev/null
python bench.py --config 3 --steps 2 --warmup 3 --timesteps 200 > $O/r02_s6_bench_c3_t200.json 2> /dev/null
python bench.py --config 3cfg8 --steps 3 --warmup 3 > $O/r02_s6_bench_c3cfg8.json 2> /dev/null
python bench.py --config 4 --steps 2 --warmup 3 > $O/r02_s6_bench_c4.json 2> /dev/null
python bench.py --config 5 --steps 10 --warmup 3 > $O/r02_s6_bench_c5.json 2> /dev/null
