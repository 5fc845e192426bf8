"""Dev helper: run one test-world scenario and print a summary (used under gpurun)."""
import sys, time, json
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from _launch import run_world
n = int(sys.argv[1]); scenario = sys.argv[2]; args = sys.argv[3:]
t = time.time()
import os
res = run_world(n, scenario, args=args, timeout=int(os.environ.get('WORLD_TIMEOUT','900')))
ok = all(r.get('ok') and r.get('returncode') == 0 for r in res)
print("WORLD n=%d %s %s -> %s in %.1fs" % (n, scenario, ' '.join(args), 'OK' if ok else 'FAIL', time.time() - t))
for r in res:
    print("  rank", r['rank'], {k: v for k, v in r.items() if k not in ('log',)})
    if not r.get('ok'):
        print(r['log'][-2500:])
sys.exit(0 if ok else 1)
