cd $GRAFT_REPO_ROOT
for W in 17 16; do B200_WIRE_WINDOW=$W timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-roofline --no-uniform 2>/dev/null | cut -c1-110; done
B200_WIRE_WINDOW=18 timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('w18', d['value'], d['e2e']['value'], d['uniform_variant'], d['config']['key'])"
timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('w20', d['value'], d['e2e']['value'], d['uniform_variant'], d['config']['key'])"
