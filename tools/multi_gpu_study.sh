#!/bin/bash
# Multi-GPU end-to-end study (run on an 8-GPU box): raw copy ceilings per direction at N = 2/4/8, the headline
# bench at N = 2/4/8, the N = 8 experiments of DESIGN.md section 5 (huge-page pinned memory, fewer slots, staggered
# ranks) and BASELINE config 5 on 8 GPUs.  Output: gpurun_out/mg_*.
out=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $out/mg_topo.txt 2>&1
(lscpu | head -25; echo; numactl -H 2>/dev/null; echo; cat /sys/kernel/mm/transparent_hugepage/enabled; grep -i huge /proc/meminfo) > $out/mg_host.txt 2>&1
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29500 + n)) tools/pcie_ceiling.py > $out/mg_pcie_n$n.log 2>&1
done
D2PC_PINNED_THP=1 $TR --nproc-per-node 8 --master-port 29520 tools/pcie_ceiling.py > $out/mg_pcie_n8_thp.log 2>&1
rm -f $out/mg_scale.jsonl $out/mg_exp.jsonl
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29530 + n)) bench.py --gpus $n --config 4 --steps 3 --e2e-frames 512 --no-cpu \
      --append $out/mg_scale.jsonl > $out/mg_bench_n$n.log 2>&1
done
D2PC_PINNED_THP=1 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --config 4 --steps 2 --e2e-frames 512 --no-cpu \
    --append $out/mg_exp.jsonl > $out/mg_bench_n8_thp.log 2>&1
$TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --config 4 --steps 2 --e2e-frames 512 --no-cpu --slots 2 --no-ceiling \
    --append $out/mg_exp.jsonl > $out/mg_bench_n8_slots2.log 2>&1
$TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --config 4 --steps 2 --e2e-frames 512 --no-cpu --stagger-us 3000 --no-ceiling \
    --append $out/mg_exp.jsonl > $out/mg_bench_n8_stagger.log 2>&1
$TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --config 5 --steps 3 --no-cpu \
    --append $out/mg_config5_n8.jsonl > $out/mg_bench_c5_n8.log 2>&1
tail -3 $out/mg_pcie_n8.log
