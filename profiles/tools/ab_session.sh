# quick A/B of bench variants on one GPU (debug tool): bash profiles/tools/ab_session.sh "<variant flags>;<variant flags>;..."
IFS=';' read -ra CFGS <<< "${1:-;--no-pipeline}"
for cfg in "${CFGS[@]}"; do
  for v in 8 1; do
    tag=$(echo "$cfg" | tr -d ' -')
    timeout 300 python bench.py --views $v --steps 64 --warmup 8 --no-cpu-baseline --no-ref-cuda $cfg > gpurun_out/t_v${v}_$tag.json 2> gpurun_out/t_err.txt
    echo "views $v cfg [$cfg]: $(python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['config']['step_ms'], {k:round(x,3) for k,x in d['roofline']['kernels_ms_per_step'].items()})
" gpurun_out/t_v${v}_$tag.json 2>&1 | tail -1)"
    tail -2 gpurun_out/t_err.txt | cut -c1-300
  done
done
