show() { python -c "
import sys,json
s=open('$1').read(); d=json.loads(s[s.index('{'):])
print('$2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'redo', d['steps_redone_literal'], 'alone', d['roofline'].get('alone',{}).get('ms_per_launch'), {k:v for k,v in d['queries_flagged'].items() if k!='what'}, 'rescored', d['rescored_candidates'])
"; }
timeout 600 python -m pytest tests/test_gpu_listmajor.py -x -q > gpurun_out/t_lm.log 2>&1; tail -5 gpurun_out/t_lm.log
for r in 512 256; do
VS_LM_SEED_ROWS=$r timeout 100 python bench.py --no-cpu-baseline --no-extra --steps 40 --warmup 6 > gpurun_out/ab_seed$r.json 2>> gpurun_out/ab_err.log; show gpurun_out/ab_seed$r.json seed$r
done
tail -3 gpurun_out/ab_err.log | cut -c1-300
