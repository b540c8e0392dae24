for d in 0 1 2; do MEMOTR_WINDOW_DEBUG=$d timeout 300 python tools/micro_msda.py --quick 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('debug', '$d', d['case'], 'win', round(d['window_us'],1))
"; done
