cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_oracle_fullsize.py -x -q 2>&1 | tail -3
bash tools/ab.sh 2 build/lib_v8.so build/lib_v11.so > gpurun_out/r2b_ab8.txt 2>&1
cat gpurun_out/r2b_ab8.txt
DESMAN_B200_LIB=build/libdesman_b200_kprof.so timeout 200 python tools/kprof.py > gpurun_out/r2b_kprof11.txt 2>&1
grep "^mu_\|^draw\|^maintain\|^ll_\|^finalize\|^tau" gpurun_out/r2b_kprof11.txt | tail -9
