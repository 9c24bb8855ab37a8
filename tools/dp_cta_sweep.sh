#!/bin/bash
# N-GPU training step with the gradient communicator capped to a few CTAs (NFDPM_NCCL_MAX_CTAS); usage: dp_cta_sweep.sh N
N=${1:-2}
for c in ${CTAS:-default 16 8 4 2}; do
  if [ "$c" = default ]; then export NFDPM_NCCL_MAX_CTAS=0; else export NFDPM_NCCL_MAX_CTAS=$c; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
    bench.py --gpus $N --only-train --steps 30 --warmup 5 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); t=d['train']; print('max_ctas=$c', 'N=%d' % d['n_gpus'], 'ms/step=%.3f' % t['ms_per_step'], 'img/s=%.0f' % t['value'], 'exposed_us=%.0f' % (t['exposed_comm_us_per_step'] or 0))"
done
