set -x
cd $GRAFT_REPO_ROOT
for v in r1 r2 r3; do
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python scripts/march_probe.py 20 > gpurun_out/r2e_probe_$v.json 2>&1
  NFA_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sampling or traverse or full_size" 2>&1 | tail -n 3 > gpurun_out/r2e_pytest_$v.log
done
NFA_LIB=$PWD/gpurun_variants/lib_r3.so ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 1 -f -o gpurun_out/r2e_march_r3 python scripts/profile_kernels.py step > gpurun_out/r2e_prof.log 2>&1
python scripts/host_overhead.py > gpurun_out/r2e_host.log 2>&1
head -n 1 gpurun_out/r2e_probe_*.json; tail -n 3 gpurun_out/r2e_pytest*.log; head -n 6 gpurun_out/r2e_host.log
