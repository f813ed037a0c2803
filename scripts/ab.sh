# A/B of the dedupe strategies + sub-batch sweep; prints value and per-stage ms
run() { python bench.py --steps 5 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.0f  e2e %.0f  stages %s  pipeline_frac %.3f' % (d['value'], d['e2e']['value'], {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, r['pipeline']['frac']))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export XCB200_SUB_BATCH=16
echo "match/match"; run
echo "hist=tag lwa=match"; XCB200_HIST_DEDUP=t run
echo "hist=match lwa=tag"; XCB200_LWA_DEDUP=t run
for sb in 8 32; do echo "sub $sb"; XCB200_SUB_BATCH=$sb run; done
