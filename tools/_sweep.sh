for g in 1 2 4 8; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tune roi_gpc=$g 2>/dev/null | tail -1 > gpurun_out/sw.json
  python -c "
import json; d=json.loads(open('gpurun_out/sw.json').read()); print('gpc', $g, d['ms_per_step'], {k:v.get('ms') for k,v in d['kernels'].items() if 'roi' in k})"
done
