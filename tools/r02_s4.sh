# session-4 GPU pass: new GroupNorm-apply kernel, coalesced stream-K scratch, cost-model tile/split choice
set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -8 > $O/r02_s4_pytest.log
python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k heavy_tailed 2>&1 | grep -E "heavy-tailed|passed|failed" > $O/r02_s4_heavy.log
rm -f $O/r02_s4_prof.jsonl
for b in 1 4 16 32 64; do
  for sf in 0 8; do
    B=$b SPLIT_FILL=$sf python tools/step_profile.py 2>/dev/null | sed "s/^/{\"B\": $b, \"split_fill\": $sf} /" >> $O/r02_s4_prof.jsonl
  done
done
B=4 SPLIT_FILL=8 BLOCK_N=256 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": 4, \"split_fill\": 8, \"block_n\": 256} /" >> $O/r02_s4_prof.jsonl
B=4 SPLIT_FILL=8 BLOCK_N=64 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": 4, \"split_fill\": 8, \"block_n\": 64} /" >> $O/r02_s4_prof.jsonl
B=4 SPLIT_FILL=4 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": 4, \"split_fill\": 4} /" >> $O/r02_s4_prof.jsonl
B=4 SPLIT_FILL=16 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": 4, \"split_fill\": 16} /" >> $O/r02_s4_prof.jsonl
python bench.py --config 1 --steps 5 --warmup 3 > $O/r02_s4_bench_c1.json 2> $O/r02_s4_bench_c1.err
python bench.py --config 3cfg8 --steps 3 --warmup 3 > $O/r02_s4_bench_c3cfg8.json 2> /dev/null
MF_SPLIT_FILL=0 python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s4_bench_c2_t100_sf0.json 2> /dev/null
python bench.py --config 2 --steps 2 --warmup 3 --timesteps 100 > $O/r02_s4_bench_c2_t100_sf8.json 2> /dev/null
python bench.py --config 5 --steps 10 --warmup 3 > $O/r02_s4_bench_c5.json 2> /dev/null
