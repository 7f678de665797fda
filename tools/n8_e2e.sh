TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/pcie_bw.py
$TR --nproc-per-node 8 --master-port 29512 tools/pcie_bw.py 2>/dev/null
for S in 1 2 4; do
$TR --nproc-per-node 8 --master-port 2952$S bench.py --gpus 8 --steps 5 --warmup 3 --streams $S --skip-extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams',$S,'value',round(d['value']),'e2e',round(d['e2e']['value']), 'ms', d['e2e']['ms_per_step'])"
done
