set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_s7_pytest.log
B=64 python tools/vae_ops.py > $O/r02_s7_vae.log 2>&1
B=64 python tools/step_profile.py > $O/r02_s7_prof.jsonl 2>/dev/null
B=4 python tools/step_profile.py > $O/r02_s7_prof_b4.jsonl 2>/dev/null
