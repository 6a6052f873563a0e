# Round-2 iteration run (one GPU): parity tests on the current build, A/B bench lines against other builds
# in ab_builds/ (B9GPU_LIB), one full ncu capture of the identity main kernel. Usage:
#   gpurun -- 'TAG=s1 AB="r1" bash scripts/gpu_r2_step.sh'
set -x
mkdir -p gpurun_out
TAG=${TAG:-step}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_tests.log
for rep in 1 2; do
  for lib in cur ${AB}; do
    if [ "$lib" = cur ]; then unset B9GPU_LIB; else export B9GPU_LIB=$PWD/ab_builds/$lib.so; fi
    timeout 200 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/${TAG}_${lib}_adv1_$rep.json 2> gpurun_out/${TAG}_${lib}.err
    timeout 200 python bench.py --no-cpu-baseline --e2e-steps 2 --adversarial 0 > gpurun_out/${TAG}_${lib}_adv0_$rep.json 2>> gpurun_out/${TAG}_${lib}.err
  done
done
unset B9GPU_LIB
timeout 200 python bench.py --no-cpu-baseline --e2e-steps 2 --handler vadd_f32 > gpurun_out/${TAG}_cur_vadd.json 2> gpurun_out/${TAG}_vadd.err
if [ -n "$AB" ]; then B9GPU_LIB=$PWD/ab_builds/${AB%% *}.so timeout 200 python bench.py --no-cpu-baseline --e2e-steps 2 --handler vadd_f32 > gpurun_out/${TAG}_${AB%% *}_vadd.json 2>> gpurun_out/${TAG}_vadd.err; fi
if [ -z "$NO_NCU" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain3_kernel -s 2 -c 1 -o gpurun_out/${TAG}_ncu_identity_main -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:drain_slow -s 2 -c 1 -o gpurun_out/${TAG}_ncu_identity_slow -f python bench.py --no-cpu-baseline --e2e-steps 1 --steps 2 --warmup 1 > gpurun_out/ncu_b.log 2>&1
fi
for f in gpurun_out/${TAG}_*_adv*.json gpurun_out/${TAG}_*_vadd.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "value %.3g  kernel_ms %.4f  frac %.3f  e2e %.3g" % (d["value"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
