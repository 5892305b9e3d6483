#!/bin/bash
# Round-2 multi-GPU records (one 8-GPU box): weak- and strong-scaling bench lines, and BASELINE configs[4]
# (library generation: 2000 samples x 100 realizations, PS_Classic with its 0.3 dex scatter) at 1 and 8 GPUs.
#   gpurun --gpus 8 -- bash profiles/r2_multigpu.sh
set -u
OUT=gpurun_out/r2_multigpu
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# (the driver records the weak-scaling curve itself at round end: SCALE_r02.json; here the STRONG-scaling line)
for N in ${HOLO_WEAK_NS:-8}; do
  $TR --nproc-per-node $N --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_weak_n$N.json 2> $OUT/bench_weak_n$N.err
done
for N in ${HOLO_STRONG_NS:-8}; do
  $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline > $OUT/bench_strong_n$N.json 2> $OUT/bench_strong_n$N.err
done
GL="-m holodeck_b200.librarian.gen_lib PS_Classic_Phenom_Uniform"
rm -rf /tmp/lib1 /tmp/lib8 /tmp/lib8w1 /tmp/lib1ref
python $GL /tmp/lib1ref -n 60 -r 100 -l 5 --gwb --ss --params --seed 1 --no-streaming > $OUT/genlib_1gpu_reference_fileplane.log 2>&1
python $GL /tmp/lib1 -n 250 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_1gpu.log 2>&1
$TR --nproc-per-node 8 --master-port 29520 $GL /tmp/lib8 -n 2000 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_8gpu.log 2>&1
$TR --nproc-per-node 8 --master-port 29521 $GL /tmp/lib8w1 -n 2000 -r 100 -l 5 --gwb --ss --params --seed 1 --no-pipeline --no-combine > $OUT/genlib_8gpu_nopipeline.log 2>&1
ls -la /tmp/lib8 /tmp/lib8/library_store | head -20 >> $OUT/genlib_8gpu.log
grep -h "library:\|combined\|rank 0" $OUT/genlib_*.log
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
