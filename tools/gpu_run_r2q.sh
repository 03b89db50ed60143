set -x
O=gpurun_out/r2q; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64_batch" > $O/memcheck_fp64.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck_fp64.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64_batch and (44100-48000 or 192000-44100 or 384000-48000 or 32000-24000-p50-q0-2ch)" > $O/racecheck_fp64.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck_fp64.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp64_batch and (44100-48000 or 192000-44100)" > $O/synccheck_fp64.log 2>&1; echo "synccheck rc=$?" >> $O/synccheck_fp64.log
tail -n 6 $O/memcheck_fp64.log $O/racecheck_fp64.log $O/synccheck_fp64.log
