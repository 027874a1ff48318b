# 8 ranks: topology, then the end-to-end step with and without binding each rank to its GPU's NUMA node
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|model name" | head -8
for flag in "--no-numa-bind" ""; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 --e2e-steps 2 $flag 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flag [$flag]', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'numa', d['e2e'].get('numa_node'))"
done
