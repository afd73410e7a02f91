# Final single-GPU pass of round 2 (run through gpurun): GPU tests, smoke(), per-launch step profiles, one bench line per config.
# Knob A/B runs of the round used the same tools with MF_ROW_PATCH / MF_SPLIT_FILL (env, medfusion_b200/_lib.py) or
# SPLIT_FILL / BLOCK_N (tools/step_profile.py) set.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_final_smoke.log 2>&1
python bench.py --config 2 --steps 2 --warmup 3 > $O/r02_final_bench_c2.json 2> $O/r02_final_bench_c2.err
python bench.py --config 1 --steps 5 --warmup 3 > $O/r02_final_bench_c1.json 2> /dev/null
python bench.py --config 3 --steps 2 --warmup 3 --timesteps 200 > $O/r02_final_bench_c3_t200.json 2> /dev/null
python bench.py --config 3cfg8 --steps 3 --warmup 3 > $O/r02_final_bench_c3cfg8.json 2> /dev/null
python bench.py --config 4 --steps 2 --warmup 3 > $O/r02_final_bench_c4.json 2> /dev/null
python bench.py --config 5 --steps 10 --warmup 3 > $O/r02_final_bench_c5.json 2> /dev/null
