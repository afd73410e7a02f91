# Small-batch fill (mf_set_split_fill) A/B on one B200: per-launch step profiles at B=4 / B=32, then the bench lines
set -x
O=gpurun_out
for b in 4 32; do
  for sf in 0 8 4 16; do
    B=$b SPLIT_FILL=$sf python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": $b, \"split_fill\": $sf} /" >> $O/r02_smallbatch_prof.jsonl
  done
  B=$b SPLIT_FILL=8 BLOCK_N=128 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": $b, \"split_fill\": 8, \"block_n\": 128} /" >> $O/r02_smallbatch_prof.jsonl
  B=$b SPLIT_FILL=8 BLOCK_N=64 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": $b, \"split_fill\": 8, \"block_n\": 64} /" >> $O/r02_smallbatch_prof.jsonl
done
B=64 SPLIT_FILL=8 python tools/step_profile.py 2>/dev/null | head -1 | sed "s/^/{\"B\": 64, \"split_fill\": 8} /" >> $O/r02_smallbatch_prof.jsonl
python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6 > $O/r02_s3_pytest.log
python bench.py --config 1 --steps 5 --warmup 3 > $O/r02_s3_bench_c1.json 2> $O/r02_s3_bench_c1.err
python bench.py --config 3cfg8 --steps 3 --warmup 3 > $O/r02_s3_bench_c3cfg8.json 2> /dev/null
