set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > $O/r02_s5_pytest.log
B=64 python tools/step_profile.py > $O/r02_s5_prof.jsonl 2>/dev/null
MF_ROW_PATCH=0 B=64 python tools/vae_ops.py > $O/r02_s5_vae_off.log 2>&1
MF_ROW_PATCH=1 B=64 python tools/vae_ops.py > $O/r02_s5_vae_on.log 2>&1
python bench.py --config 5 --steps 10 --warmup 3 > $O/r02_s5_bench_c5.json 2> $O/r02_s5_bench_c5.err
python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s5_bench_c2_t100.json 2> /dev/null
python bench.py --config 4 --steps 2 --warmup 3 --timesteps 50 > $O/r02_s5_bench_c4_t50.json 2> /dev/null
B=64 ncu --set full --clock-control none -k regex:'conv_tc|gn_apply' -s 22 -c 22 -f -o /tmp/r02_vae_decode2 python tools/vae_once.py 2 > $O/r02_s5_ncu_vae.log 2>&1
ncu -i /tmp/r02_vae_decode2.ncu-rep --page raw --csv > $O/r02_s5_vae_decode_raw.csv 2>/dev/null
