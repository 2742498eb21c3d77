mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --reads 50000000 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$*', round(d['value']/1e9,1), round(d['e2e']['value']/1e9,2), d['e2e']['ms_per_step'])
"; }
run A=1
run B200SK_PINNED_WC=1
run B200SK_SUB_BYTES=201326592
run B200SK_SUB_BYTES=805306368
run A=2
