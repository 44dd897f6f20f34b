#!/bin/bash
# 8-GPU box: config 2 (weak scaling, as the driver runs it), config 4 strong scaling at 2/4/8, configs 3 and 5 at 8, CLI on 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MFKC_BENCH_NO_CPU=1 MFKC_BENCH_NO_INGEST=1
timeout 600 $TR --nproc-per-node 8 --master-port 29501 bench.py --gpus 8 --steps 9 --warmup 3 > gpurun_out/n8_cfg2.json 2> gpurun_out/n8_cfg2.err; echo "rc=$?" >> gpurun_out/n8_cfg2.err
timeout 600 $TR --nproc-per-node 4 --master-port 29502 bench.py --gpus 4 --steps 9 --warmup 3 > gpurun_out/n4_cfg2.json 2> gpurun_out/n4_cfg2.err; echo "rc=$?" >> gpurun_out/n4_cfg2.err
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --config 4 --reads 200000000 --steps 3 --warmup 2 > gpurun_out/n${n}_cfg4.json 2> gpurun_out/n${n}_cfg4.err; echo "rc=$?" >> gpurun_out/n${n}_cfg4.err
done
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --config 3 --steps 2 --warmup 1 > gpurun_out/n8_cfg3.json 2> gpurun_out/n8_cfg3.err; echo "rc=$?" >> gpurun_out/n8_cfg3.err
timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --config 5 --reads 80000000 --steps 3 --warmup 2 > gpurun_out/n8_cfg5.json 2> gpurun_out/n8_cfg5.err; echo "rc=$?" >> gpurun_out/n8_cfg5.err
CLI=metafast_b200/bin/mfkc_cli; D=/tmp/cli8; mkdir -p $D
for s in 0 1 2; do $CLI gen-reads $D/s$s.fastq 400000 $s > /dev/null 2>&1; done
( $CLI -t kmer-counter-many -k 31 -b 2 -i $D/s0.fastq $D/s1.fastq $D/s2.fastq -w $D/one --gpus 1 > $D/one.out 2> $D/one.err
  $CLI -t kmer-counter-many -k 31 -b 2 -i $D/s0.fastq $D/s1.fastq $D/s2.fastq -w $D/shard --gpus 8 --gpu-mode shard > $D/shard.out 2> $D/shard.err
  $CLI -t kmer-counter-many -k 31 -b 2 -i $D/s0.fastq $D/s1.fastq $D/s2.fastq -w $D/samples --gpus 3 --gpu-mode samples > $D/samples.out 2> $D/samples.err
  for m in shard samples; do for s in 0 1 2; do cmp $D/one/kmers/s$s.kmers.bin $D/$m/kmers/s$s.kmers.bin && cmp $D/one/stats/s$s.stat.txt $D/$m/stats/s$s.stat.txt && echo "cli --gpu-mode $m: s$s.kmers.bin and s$s.stat.txt identical to the one-GPU run"; done; done
  tail -n 3 $D/shard.err ) > gpurun_out/n8_cli.log 2>&1
cat gpurun_out/n8_cli.log
for f in n8_cfg2 n4_cfg2 n8_cfg4 n4_cfg4 n2_cfg4 n8_cfg3 n8_cfg5; do echo "== $f $(tail -n 1 gpurun_out/$f.err)"; done
