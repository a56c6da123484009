python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
python tools/cat_sweep.py 4096 16384 65536 > gpurun_out/r2n_cat_fused.json 2>&1
CATB200_CAT_FUSED=0 python tools/cat_sweep.py 4096 16384 65536 > gpurun_out/r2n_cat_2launch.json 2>&1
tail -1 gpurun_out/r2n_cat_fused.json; tail -1 gpurun_out/r2n_cat_2launch.json
python tools/kernel_shares.py > gpurun_out/r2n_shares.txt 2>&1; head -22 gpurun_out/r2n_shares.txt | tail -20
