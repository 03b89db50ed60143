# 8 GPUs of one box: multi-GPU C ABI (NCCL gather), strong-scaling cfg4, 10-hour cfg5, C harness
set -x
O=gpurun_out/r2final2_8gpu; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log; tail -n 5 $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/scale_cfg4_n8.json 2> $O/scale_cfg4_n8.err
$TR bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 3 > $O/scale_cfg5_n8.json 2> $O/scale_cfg5_n8.err
$TR bench.py --gpus 8 --weak --steps 3 --warmup 3 --no-e2e > $O/scale_cfg4_weak_n8.json 2> $O/scale_cfg4_weak_n8.err
./examples/harness_multi batch 48000 44100 2 1024 10 8 > $O/harness_multi.txt 2>&1
./examples/harness_multi batch 48000 44100 2 1024 10 1 >> $O/harness_multi.txt 2>&1
./examples/harness_multi stream 384000 48000 8 1 300 8 >> $O/harness_multi.txt 2>&1
./examples/harness_multi stream 384000 48000 8 1 300 1 >> $O/harness_multi.txt 2>&1
for f in $O/*.err; do tail -n 3 "$f"; done | tail -n 20
cut -c1-500 $O/scale_cfg4_n8.json; cut -c1-500 $O/scale_cfg5_n8.json; cat $O/harness_multi.txt
