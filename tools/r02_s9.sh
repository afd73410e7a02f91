set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_s9_pytest.log
B=64 DUMP=1 python tools/step_profile.py > $O/r02_s9_prof.jsonl 2>/dev/null
python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s9_bench_c2_t100.json 2> /dev/null
python bench.py --config 5 --steps 10 --warmup 3 > $O/r02_s9_bench_c5.json 2> /dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_s9_smoke.log 2>&1
