set -x
cd $GRAFT_REPO_ROOT
python scripts/composite_probe.py 30 > gpurun_out/r2g_comp_base.json 2>&1
for v in c5 c6; do NFA_LIB=$PWD/gpurun_variants/lib_$v.so python scripts/composite_probe.py 30 > gpurun_out/r2g_comp_$v.json 2>&1; done
NFA_EXTRA_ONLY=scans,f3,f4 timeout 600 python scripts/extra_configs.py reference-cuda > gpurun_out/r2g_extra_ref.log 2>&1
cat gpurun_out/r2g_comp_*.json; tail -n 4 gpurun_out/r2g_extra_ref.log
