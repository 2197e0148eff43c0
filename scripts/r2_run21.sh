set -x
cd $GRAFT_REPO_ROOT
for V in one per_slot; do
  if [ $V = per_slot ]; then export NFA_LIB=$GRAFT_REPO_ROOT/build_variants/libnfa_atomic_per_slot.so; else unset NFA_LIB; fi
  for S in 8 16; do
    NFA_MARCH_SLOTS=$S python scripts/march_probe.py 20 > gpurun_out/r2r_probe_${V}_s$S.txt 2>&1
    echo "== $V slots $S"; cat gpurun_out/r2r_probe_${V}_s$S.txt
  done
done
unset NFA_LIB
NFA_MARCH_SLOTS=8 python scripts/march_trace.py > gpurun_out/r2r_trace_one_s8.txt 2>&1
cat gpurun_out/r2r_trace_one_s8.txt
