#!/bin/bash
# forward+inverse at config 2 with the deep-level StepFlow chains split into concurrent sub-batches (NFDPM_DEEP_STREAMS=n1:n2)
for spec in ${SPECS:-1:1 2:2 2:4 4:4 4:8 2:8}; do
  NFDPM_DEEP_STREAMS=$spec python bench.py --no-train --no-cpu-baseline --no-eager-gpu --steps 30 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); m=d['modes']
print('deep_streams=$spec', 'fp32-faithful: step %.3f ms fwd %.3f inv %.3f' % (d['ms_per_step'], d['directions']['forward']['ms'], d['directions']['inverse']['ms']), '| bf16: step %.3f ms' % m['bf16']['ms_per_step'], '| recon %.2e z_rel %.2e' % (d['checks']['recon_max_abs_err'], d['checks'].get('z_rel_l2_vs_oracle', -1)))"
done
