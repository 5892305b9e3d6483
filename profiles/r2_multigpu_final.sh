#!/bin/bash
# Final multi-GPU records of round 2 (one 8-GPU box):  gpurun --gpus 8 -- bash profiles/r2_multigpu_final.sh
set -u
OUT=gpurun_out/r2_multigpu_final
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/bench_weak_n8.json 2> $OUT/bench_weak_n8.err
$TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline > $OUT/bench_strong_n8.json 2> $OUT/bench_strong_n8.err
GL="-m holodeck_b200.librarian.gen_lib PS_Classic_Phenom_Uniform"
rm -rf /tmp/lib0 /tmp/lib1 /tmp/lib8
python $GL /tmp/lib0 -n 20 -r 100 -l 5 --gwb --ss --params --seed 1 --no-combine > /dev/null 2>&1      # warm the on-disk geometry cache
python $GL /tmp/lib1 -n 250 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_1gpu.log 2>&1
$TR --nproc-per-node 8 --master-port 29520 $GL /tmp/lib8 -n 2000 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_8gpu.log 2>&1
grep -h "library:\|combined" $OUT/genlib_1gpu.log $OUT/genlib_8gpu.log
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], d["n_gpus"], d["scaling"], "value %.4e" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4e" % d["e2e"]["value"],
          "lib %.1f/s" % d["library_sample"]["samples_per_s"])
except Exception as err:
    print(sys.argv[1], "unreadable:", err)
PY
done
