python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/r02_chk_c1.json 2> gpurun_out/r02_chk_c1.err
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r02_chk_c5.json 2> /dev/null
python bench.py --steps 1 --warmup 3 --timesteps 20 > gpurun_out/r02_chk_c2.json 2> gpurun_out/r02_chk_c2.err
for c in unet_canon_b5 vae_canon_b1; do
  timeout 150 compute-sanitizer --tool synccheck --print-limit 5 python tests/ragged_check.py --case $c 2>&1 | grep -E "=========|$c" | tail -6 > gpurun_out/r02_final_synccheck_$c.log
  timeout 150 compute-sanitizer --tool racecheck --print-limit 3 python tests/ragged_check.py --case $c 2>&1 | grep -E "=========|$c" | tail -14 > gpurun_out/r02_final_racecheck_$c.log
done
