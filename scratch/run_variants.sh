for v in rng rngpos smem; do
  echo "== parity $v"; CELERITAS_B200_LIB=celeritas_b200/variants/lib_$v.so python -m pytest tests/test_gpu_testem3.py tests/test_gpu_simple_compton.py tests/test_gpu_field.py -x -q 2>&1 | tail -2
done
for v in default rng rngpos smem; do
  lib=celeritas_b200/variants/lib_$v.so; [ $v = default ] && lib=celeritas_b200/libceleritas_b200.so
  CELERITAS_B200_LIB=$lib python bench.py --no-extra --no-cpu-baseline --steps 4 > gpurun_out/bench_var_$v.json 2>gpurun_out/bench_var_$v.err
  python -c "
import json,sys
r=json.loads(open('gpurun_out/bench_var_$v.json').read().strip().splitlines()[-1])
pa=r['roofline']['per_action_seconds']
print('$v', 'value %.4g ms %.2f' % (r['value'], r['ms_per_step']), {k[:12]: round(v*1e3,2) for k,v in pa.items()})
"
done
for v in default rng rngpos smem; do
  lib=celeritas_b200/variants/lib_$v.so; [ $v = default ] && lib=default
  bash scratch/ncu_layout.sh $lib $v
done
