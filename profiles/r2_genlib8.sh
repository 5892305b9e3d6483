#!/bin/bash
# library generation only (BASELINE configs[4]) at 1 and 8 GPUs:  gpurun --gpus 8 -- bash profiles/r2_genlib8.sh
OUT=gpurun_out/r2_multigpu; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GL="-m holodeck_b200.librarian.gen_lib PS_Classic_Phenom_Uniform"
rm -rf /tmp/lib1 /tmp/lib8
python $GL /tmp/lib0 -n 20 -r 100 -l 5 --gwb --ss --params --seed 1 --no-combine > /dev/null 2>&1      # warm the on-disk geometry cache
python $GL /tmp/lib1 -n 250 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_1gpu_final.log 2>&1
$TR --nproc-per-node 8 --master-port 29520 $GL /tmp/lib8 -n 2000 -r 100 -l 5 --gwb --ss --params --seed 1 > $OUT/genlib_8gpu_final.log 2>&1
grep -h "library:\|combined" $OUT/genlib_1gpu_final.log $OUT/genlib_8gpu_final.log
nproc; python -c "
import numpy as np
a=np.load('/tmp/lib8/sam-library.npz'); print({k: a[k].shape for k in a.files if not k.startswith('attrs')}, 'nan rows', int(np.isnan(a['gwb']).any(axis=(1,2)).sum()))"
