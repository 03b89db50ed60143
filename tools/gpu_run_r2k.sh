set -x
O=gpurun_out/r2k; mkdir -p $O
python tools/d64_probe.py all > $O/probe_default.txt 2>&1
B200RATE_D64_LANE_MAJOR=1 python tools/d64_probe.py all > $O/probe_lane_major.txt 2>&1
for LM in 0 1; do for GT in 64 128; do for G in 2 3 4 6; do echo "LM=$LM GT=$GT G=$G" >> $O/probe_sweep.txt; if [ $LM = 1 ]; then export B200RATE_D64_LANE_MAJOR=1; else unset B200RATE_D64_LANE_MAJOR; fi; B200RATE_D64_GT=$GT B200RATE_D64_GROUPS=$G python tools/d64_probe.py 2>&1 | cut -c1-60 >> $O/probe_sweep.txt; done; done; done
unset B200RATE_D64_LANE_MAJOR
ncu --set full --clock-control none --import-source on -k regex:'dft64' -c 2 -s 2 -o $O/prof_dft64 -f python tools/d64_probe.py > $O/ncu_f.log 2>&1
cat $O/probe_default.txt $O/probe_lane_major.txt; cat $O/probe_sweep.txt
