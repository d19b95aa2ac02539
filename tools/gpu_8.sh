#!/bin/bash
# 8-GPU visit: bench.py at N = 8 (torchrun; carries c5_strong), C5 strong scaling through the C ABI alone (sp_render_shards),
# raw PCIe ceiling at 1 / 2 / 4 / 8 ranks.  usage (under gpurun --gpus 8): bash tools/gpu_8.sh TAG
TAG=${1:-r2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_8gpu_$TAG.json 2> $OUT/bench_8gpu_$TAG.err; echo "rc=$?" >> $OUT/bench_8gpu_$TAG.err
for g in 8 4 2; do C5_GPUS=$g C5_SAMPLES=$((1<<33)) timeout -s KILL 600 python tools/c5_shards.py >> $OUT/c5_shards_$TAG.jsonl 2>> $OUT/c5_shards_$TAG.err; done
for n in 1 2 4 8; do timeout -s KILL 300 $TR --nproc-per-node $n --master-port $((29520+n)) tools/pcie_ceiling.py >> $OUT/pcie_ceiling_$TAG.jsonl 2>> $OUT/pcie_ceiling_$TAG.err; done
timeout -s KILL 300 python -m pytest tests -m gpu -q -k "render_shards or multi_device" > $OUT/pytest_8gpu_$TAG.log 2>&1
cut -c1-300 $OUT/bench_8gpu_$TAG.json; tail -2 $OUT/bench_8gpu_$TAG.err; cat $OUT/c5_shards_$TAG.jsonl | cut -c1-400; cat $OUT/pcie_ceiling_$TAG.jsonl; tail -2 $OUT/pytest_8gpu_$TAG.log
