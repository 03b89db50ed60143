set -x
mkdir -p gpurun_out/r1b
python bench.py > gpurun_out/r1b/bench_cfg4.json 2> gpurun_out/r1b/bench_cfg4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b/bench_reference_cfg4.json 2> gpurun_out/r1b/bench_reference.err
for w in cfg1 cfg1x256 cfg2 cfg3 cfg5; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1b/bench_$w.json 2> gpurun_out/r1b/bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b/launches_cfg4x256.csv python bench.py --workload cfg4 --streams 256 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dftp_kernel|poly0_pair' -c 2 -s 8 -o gpurun_out/r1b/prof_cfg4x256 -f python bench.py --workload cfg4 --streams 256 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'halfband_pair' -c 2 -s 4 -o gpurun_out/r1b/prof_cfg5_halfband -f python tools/stage_probe.py > gpurun_out/r1b/ncu_h.log 2>&1
python tools/stage_probe.py > gpurun_out/r1b/stage_probe.txt 2>&1
for f in gpurun_out/r1b/*.err; do tail -n 2 "$f"; done | tail -n 30
