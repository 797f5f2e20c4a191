# QuantumVolume(30) through bench.py under kernel / planner debug knobs: "label env..." per line of VARIANTS
run() { label=$1; shift; env "$@" timeout 100 python bench.py --qubits 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['ms_per_step'],1), round(d['roofline']['launch_ms'],3), d['config']['passes'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
if [ $# -eq 0 ]; then set -- "default A=1" "all_threads_poll QSV_DEBUG_FLAGS=256"; fi
for v in "$@"; do run $v; done
