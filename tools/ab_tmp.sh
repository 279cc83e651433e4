timeout 300 python -m pytest tests/test_gpu_listmajor.py tests/test_gpu_search.py -m gpu -q -x 2>&1 | tail -2
show() { python -c "
import sys,json
s=open('$1').read(); d=json.loads(s[s.index('{'):])
print('$2', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'], d['steps_redone_literal'], d.get('rescored_candidates'), {k:v for k,v in d['queries_flagged'].items() if k!='what'})
"; }
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-extra --contexts 1 > gpurun_out/ab1.json 2>gpurun_out/ab_err.log; show gpurun_out/ab1.json fork_1ctx
VS_LM_SEED_STAGE=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --contexts 1 > gpurun_out/ab2.json 2>>gpurun_out/ab_err.log; show gpurun_out/ab2.json stageseed_1ctx
done
timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/ab3.json 2>>gpurun_out/ab_err.log; show gpurun_out/ab3.json fork_4ctx
VS_LM_SEED_STAGE=1 timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/ab4.json 2>>gpurun_out/ab_err.log; show gpurun_out/ab4.json stageseed_4ctx
tail -3 gpurun_out/ab_err.log
