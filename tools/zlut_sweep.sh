for cfg in "4 0" "3 0" "5 0" "6 0" "8 0" "4 2" "3 2" "4 1"; do set -- $cfg; echo "S=$1 CT=$2: $(VSB_PACK_STAGES=$1 VSB_PACK_CTAS=$2 python bench.py --workload zlut --no-e2e --no-cpu-baseline --no-extras --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']/32, {k:v['ms_per_layer'] for k,v in d['stages'].items()})")"; done
