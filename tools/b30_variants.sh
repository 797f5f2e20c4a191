run() { label=$1; shift; env "$@" timeout 100 python bench.py --qubits 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['ms_per_step'],1), round(d['roofline']['launch_ms'],3), d['config']['passes'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
run default A=1
run early128 QSV_DEBUG_FLAGS=128
run low3 QSV_LOW_QUBITS=3
run low3_early QSV_LOW_QUBITS=3 QSV_DEBUG_FLAGS=128
