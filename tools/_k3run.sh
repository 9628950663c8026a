python -m pytest tests/test_gpu_warp.py -x -q 2>&1 | tail -4
python tools/k3time.py 2>&1 | tail -1
for v in k3m3; do IMGCORR_LIB=$PWD/variants/$v.so python tools/k3time.py 2>&1 | tail -1; done
