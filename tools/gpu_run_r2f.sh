set -x
O=gpurun_out/r2f; mkdir -p $O
CMD="python bench.py --workload cfg4 --streams 256 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
for G in 4 5 6; do B200RATE_DFT_GROUPS=$G $CMD > $O/cfg4x256_g$G.json 2> $O/cfg4x256_g$G.err; done
for G in 5 6; do B200RATE_DFT_GROUPS=$G python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_matches or batch_cases or fixture or many_identical or fuzz" > $O/pytest_g$G.log 2>&1; echo "rc=$?" >> $O/pytest_g$G.log; tail -n 3 $O/pytest_g$G.log; done
B200RATE_DFT_GROUPS=5 python tools/stage_probe.py > $O/stage_probe_g5.txt 2>&1
B200RATE_DFT_GROUPS=6 python tools/stage_probe.py > $O/stage_probe_g6.txt 2>&1
export B200RATE_DFT_GROUPS=5
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dftp_kernel' -c 1 -s 4 -o $O/prof_dftp_g5 -f $CMD > $O/ncu_f.log 2>&1
for f in $O/*.err; do tail -n 3 "$f"; done | tail -n 20
for G in 4 5 6; do cut -c1-200 $O/cfg4x256_g$G.json; python -c "
import json; d=json.load(open('$O/cfg4x256_g$G.json')); print('G=$G', round(d['value']), d['roofline']['stage_ms'])"; done
grep -v 'stage ' $O/stage_probe_g5.txt; grep -v 'stage ' $O/stage_probe_g6.txt
