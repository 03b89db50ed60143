# Round-2 final single-GPU set: tests, bench lines, launch list, full ncu captures of the dominant kernels, probes.

O=gpurun_out/r2final3; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -n 3 $O/pytest.log
python bench.py > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_cfg4.json 2> $O/bench_reference.err
for w in cfg3 cfg5; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/stage_probe.py > $O/stage_probe.txt 2>&1
python tools/stream_probe.py > $O/stream_probe.txt 2>&1
for args in "44100 48000 2 65536 float" "44100 48000 2 65536 double" "44100 48000 2 262144 float" "192000 44100 8 65536 double"; do B200RATE_TRACE_STREAM=1 tools/bin/stream_lat $args >> $O/stream_latency.txt 2>&1; done
CMD="python bench.py --workload cfg4 --streams 256 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4x256.csv $CMD > $O/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dftp_kernel|poly0_pair' -c 2 -s 8 -o $O/prof_cfg4x256 -f $CMD > $O/ncu_f.log 2>&1
CMD3="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
ncu --set full --clock-control none --import-source on -k regex:'dft64|halfband_kernel|poly0_dual' -c 3 -s 12 -o $O/prof_cfg3 -f $CMD3 > $O/ncu_f3.log 2>&1
# summaries on the box (the reports themselves are too large to travel back)
python tools/ncu_summary.py $O/prof_cfg4x256.ncu-rep > $O/ncu_full_cfg4x256.txt 2>&1
python tools/ncu_summary.py $O/prof_cfg3.ncu-rep > $O/ncu_full_cfg3.txt 2>&1
mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/foo_dsp_resampler_b200/csrc/engine.o > /dev/null && cuobjdump -xelf all $GRAFT_REPO_ROOT/foo_dsp_resampler_b200/csrc/dft64.o > /dev/null; ls /tmp/cub)
python tools/ncu_by_line.py $O/prof_cfg4x256.ncu-rep /tmp/cub/engine.sm_100a.cubin _ZN8b200rate11dftp_kernelILi0ELi10ELi11ELb1ELi5EEEvNS_11DftPkParamsEx 30 dftp_kernel > $O/ncu_by_line_dftp.txt 2>&1
python tools/ncu_by_line.py $O/prof_cfg4x256.ncu-rep /tmp/cub/engine.sm_100a.cubin _ZN8b200rate18poly0_pair2_kernelILi24ELi2EEEvNS_15Poly0PairParamsEx 24 poly0_pair2 > $O/ncu_by_line_poly0_pair2.txt 2>&1
python tools/ncu_by_line.py $O/prof_cfg3.ncu-rep /tmp/cub/dft64.sm_100a.cubin _ZN8b200rate12dft64_kernelILb0EEEvNS_11Dft64ParamsEx 30 dft64_kernel > $O/ncu_by_line_dft64_cfg3.txt 2>&1
rm -f $O/*.ncu-rep
for f in $O/*.err; do tail -n 2 "$f"; done | tail -n 12
for f in $O/bench_*.json; do cut -c1-160 $f; done
python tools/lpc_probe.py > $O/lpc_probe.txt 2> $O/lpc_probe.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
