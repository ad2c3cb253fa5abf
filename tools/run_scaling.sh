#!/bin/bash
# usage: tools/run_scaling.sh N "merge modes" "spp values"   -> gpurun_out/scale_n${N}_${merge}_${spp}.json (one bench line each)
N=$1; MODES=${2:-scatter}; SPPS=${3:-1024}
mkdir -p gpurun_out
for m in $MODES; do for spp in $SPPS; do
  out=gpurun_out/scale_n${N}_${m}_${spp}
  if [ "$N" = 1 ]; then python bench.py --steps 5 --warmup 3 --spp $spp --no-cpu-baseline > $out.json 2> $out.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --spp $spp --merge $m --no-cpu-baseline > $out.json 2> $out.err; fi
  python - $out.json $m $spp <<'PY'
import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if line.startswith('{'):
        j=json.loads(line)
        print(sys.argv[2], sys.argv[3], "N=%d value %.2f G ms/step %.2f kernel_ms %.2f e2e %.2f G" % (j["n_gpus"], j["value"]/1e9, j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["value"]/1e9))
PY
done; done
