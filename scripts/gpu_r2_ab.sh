# A/B bench lines on one box. VARIANTS: space-separated "name[@ENV=VAL[,ENV=VAL]]"; name "cur" = the in-tree build,
# anything else = ab_builds/<name>.so. ARGS: extra bench.py arguments per line set (";"-separated).
#   gpurun -- 'TAG=s3 VARIANTS="cur cur@B9_STATIC_ROUNDS=0 nopdl" bash scripts/gpu_r2_ab.sh'
mkdir -p gpurun_out
TAG=${TAG:-ab}
REPS=${REPS:-2}
IFS=';' read -ra ARGSETS <<< "${ARGS:---adversarial 0.01;--adversarial 0}"
# TESTS=1 runs the GPU suite first; TESTS_LIB=<name> runs it against ab_builds/<name>.so instead of the in-tree build
if [ -n "$TESTS" ]; then ( [ -n "$TESTS_LIB" ] && export B9GPU_LIB=$PWD/ab_builds/$TESTS_LIB.so; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_tests.log ); fi
for rep in $(seq 1 $REPS); do
  for v in ${VARIANTS:-cur}; do
    lib=${v%%@*}; envs=""; [ "$v" != "$lib" ] && envs=${v#*@}
    k=0
    for a in "${ARGSETS[@]}"; do
      k=$((k+1))
      ( if [ "$lib" != cur ]; then export B9GPU_LIB=$PWD/ab_builds/$lib.so; fi
        for e in ${envs//,/ }; do export "$e"; done
        timeout 200 python bench.py --no-cpu-baseline --e2e-steps 2 $a > "gpurun_out/${TAG}_${v//[@=,]/_}_a${k}_$rep.json" 2> "gpurun_out/${TAG}_${v//[@=,]/_}.err" )
    done
  done
done
python scripts/bench_lines.py gpurun_out/${TAG}_*_a*.json
