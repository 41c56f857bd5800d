for cfg in "1 128" "2 128" "2 64" "2 32" "3 32" "4 32" "2 16" "4 16"; do set -- $cfg; python bench.py --steps 6 --warmup 3 --lanes $1 --chunk-pairs $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lanes $1 chunk $2', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))"; done
