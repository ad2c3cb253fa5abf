#!/bin/bash
# usage: tools/run_scaling.sh TAG N [bench.py arguments...]   -> gpurun_out/TAG.json (one bench line) + a one-line summary
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
out=gpurun_out/$TAG
if [ "$N" = 1 ]; then python bench.py "$@" > $out.json 2> $out.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $out.json 2> $out.err; fi
python - $out.json $TAG <<'PY'
import json,sys
ok = False
for line in open(sys.argv[1]):
    line=line.strip()
    if line.startswith('{'):
        j=json.loads(line); ok = True
        ex = j.get("film_exchange_ms") or {}
        print(sys.argv[2], "N=%d %s value %.2f G ms/step %.3f kernel_ms %.3f e2e %.2f G" % (j["n_gpus"], j["scaling"], j["value"]/1e9, j["ms_per_step"], j["roofline"]["kernel_ms"], j["e2e"]["value"]/1e9),
              "rmse_ok", (j.get("image_rmse") or {}).get("ok"), "other", {k: round(v["value"]/1e9, 2) for k, v in j.items() if k in ("weak", "strong") and isinstance(v, dict)},
              "exchange", {k: round(v["max_over_ranks"], 3) for k, v in ex.items()})
if not ok: print(sys.argv[2], "NO LINE; stderr tail:"); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
