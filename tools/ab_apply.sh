for cfg in "GG_APPLY_PER_SM=6" "GG_APPLY_PER_SM=4" "GG_APPLY_PER_SM=3" "GG_APPLY_PER_SM=8" "GG_APPLY_PER_SM=2 GG_APPLY_UB=8" "GG_APPLY_PER_SM=4 GG_APPLY_UB=8"; do
  env $cfg timeout 200 python bench.py --no-cpu-baseline --no-extra --steps 20 --repeats 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
h={x['kernel'][:22]:x['ms'] for x in d['roofline']['hbm_kernels']}
print('$cfg', round(d['ms_per_step'],4), {k:v for k,v in h.items() if k.startswith('bn_')})"
done
