nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|^CPU(s)\|Socket\|Model name" 
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')"
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
for nb in 0 1; do
  if [ $nb = 1 ]; then export CDDMSL_NO_NUMA_BIND=1; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$nb bench.py --gpus 8 --no-cpu-baseline --steps 10 2>/dev/null | tail -1 > gpurun_out/b8_$nb.json
  python -c "
import json; d=json.loads(open('gpurun_out/b8_$nb.json').read()); print('nobind=$nb', d['value'], d['e2e']['value'], d['e2e'].get('numa_bound_cpus'))"
done
