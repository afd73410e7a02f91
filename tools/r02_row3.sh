set -x
O=gpurun_out
MF_ROW_PATCH=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_tc_matches" 2>&1 | tail -8 > $O/r02_row3_v1.log
MF_ROW_PATCH=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_tc_matches" 2>&1 | tail -8 > $O/r02_row3_v2.log
MF_ROW_PATCH=1 B=64 python tools/vae_ops.py > $O/r02_row3_vae_on.log 2>&1
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_row3_pytest.log
MF_ROW_PATCH=2 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > $O/r02_row3_pytest_v2.log
