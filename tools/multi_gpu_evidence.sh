#!/bin/bash
# One 8-GPU box: the concurrent device->host probe at 1 / 2 / 4 / 8 ranks with and without NUMA-local CPU binding, the
# topology, and the bench at N = 8 (and N = 2).   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_evidence.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
probe=$out/${tag}_d2h_probe.txt
: > $probe
nvidia-smi topo -m >> $probe 2>&1
echo "host: $(nproc) CPUs, $(grep -c processor /proc/cpuinfo) logical, NUMA nodes: $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l)" >> $probe
for n in ${PROBE_RANKS:-1 2 4 8}; do
  for mode in nobind bind; do
    echo "--- $n ranks, $mode" >> $probe
    run $n tools/d2h_probe.py $mode 2>/dev/null | grep -E "^rank" | sort >> $probe
  done
done
if [ -z "$SKIP_PROBE_BENCH" ]; then
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --bind-cpus 0 > $out/${tag}_bench_n8_nobind.json 2> $out/${tag}_bench_n8_nobind.err
fi
run 8 bench.py --gpus 8 --steps 20 --warmup 3 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
run 4 bench.py --gpus 4 --steps 20 --warmup 3 > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err
run 2 bench.py --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 3 --no-configs --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python - <<PY
import json
for n in ("n1", "n2", "n4", "n8", "n8_nobind"):
    try:
        d = json.load(open("$out/${tag}_bench_%s.json" % n))
        print(n, "value %.3e kernel_ms %.3f e2e %.3e resident %.3e d2h %.1f of %.1f GB/s per GPU" % (
            d["value"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["obs_resident_in_hbm"]["value"],
            d["e2e"]["roofline"]["achieved"], d["e2e"]["roofline"]["peak"]), d["roofline"]["kernel"][:40])
    except Exception as e:
        print(n, "failed", e)
PY
cat $probe | grep -v "^$" | tail -60
