# Round-2 evidence pass on one B200 (run through gpurun): GPU tests, the ncu launch list of the bench command, one
# `--set full` capture of a whole canonical UNet step + VAE decode + the attention core, racecheck / synccheck.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 > $O/r02_s2_pytest.log
# (1) launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_c2.csv \
    python bench.py --config 2 --steps 1 --warmup 1 --timesteps 4 > $O/r02_ncu_bench.log 2>&1
# (2) full capture: one UNet step at B=64 (all launches of the 2nd forward), raw page as csv
ncu --set full --clock-control none -k regex:'conv_tc|conv_simt|gn_apply|head1x1|linear_small|pack_nchw' -s 100 -c 100 -f -o /tmp/r02_unet_step \
    python tools/unet_once.py 3 > $O/r02_ncu_unet.log 2>&1
ncu -i /tmp/r02_unet_step.ncu-rep --page raw --csv > $O/r02_unet_step_raw.csv 2>/dev/null
# (3) full capture with source: the 32x32-level conv, the attention core
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 1 -f -o $O/r02_conv32_after \
    python tools/gpu_probe.py --case perf_32_256 > $O/r02_ncu_conv32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o $O/r02_attn_tc \
    python tools/attn_bench.py > $O/r02_ncu_attn.log 2>&1
# (4) VAE decode at B=64: full capture of the 22 launches
B=64 ncu --set full --clock-control none -k regex:'conv_tc|gn_apply|pack_nchw' -s 30 -c 30 -f -o /tmp/r02_vae_decode \
    python tools/vae_once.py 3 > $O/r02_ncu_vae.log 2>&1
ncu -i /tmp/r02_vae_decode.ncu-rep --page raw --csv > $O/r02_vae_decode_raw.csv 2>/dev/null
python tools/step_profile.py > $O/r02_s2_step_profile.jsonl 2> $O/r02_s2_step_profile.err
# (5) racecheck + synccheck on small cases
for tool in racecheck synccheck; do
  for c in unet_small_b3 unet_attn_b3 vae_small_b3_u8 pipe_b3_graph; do
    timeout 240 compute-sanitizer --tool $tool --print-limit 5 python tests/ragged_check.py --case $c 2>&1 | grep -E "=========|$c" | tail -12 > $O/r02_${tool}_$c.log
  done
done
ls -la $O | tail -30
