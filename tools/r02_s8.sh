set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_s8_pytest.log
B=64 python tools/step_profile.py > $O/r02_s8_prof.jsonl 2>/dev/null
SANITIZE=vae_canon_b1,unet_canon_b5,pipe_b3_graph timeout 1200 python tests/ragged_check.py > $O/r02_s8_memcheck.log 2>&1
ncu --set full --clock-control none -k regex:'conv_tc|conv_simt|gn_apply|head1x1|linear_small|pack_nchw' -s 100 -c 100 -f -o /tmp/r02_unet_step2 \
    python tools/unet_once.py 3 > $O/r02_s8_ncu_unet.log 2>&1
ncu -i /tmp/r02_unet_step2.ncu-rep --page raw --csv > $O/r02_s8_unet_step_raw.csv 2>/dev/null
python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s8_bench_c2_t100.json 2> /dev/null
