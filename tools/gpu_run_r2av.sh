O=gpurun_out/r2av; mkdir -p $O
python -m pytest tests/test_gpu_lpc.py -x -q 2>&1 | tail -n 4
for a in wide1 wide4; do echo analysis=$a; B200RATE_LPC_ANALYSIS=$a python tools/lpc_probe.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('  ',d['workload'][:30],d['gpu_ms'],d['bit_exact_sample'])
    except Exception: print(l.rstrip()[:200])
"; done
python tools/lpc_probe.py > $O/lpc_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lpc_' -c 4 --csv --log-file $O/lpc_launches.csv python tools/lpc_probe.py > /dev/null 2>&1
cut -d, -f5,15- $O/lpc_launches.csv | tail -4 | cut -c1-160
