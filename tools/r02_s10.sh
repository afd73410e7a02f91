set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_s10_pytest.log
B=64 DUMP=1 python tools/step_profile.py > $O/r02_s10_prof.jsonl 2>/dev/null
B=4 DUMP=1 python tools/step_profile.py > $O/r02_s10_prof_b4.jsonl 2>/dev/null
python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s10_bench_c2_t100.json 2> /dev/null
