#!/bin/bash
# bench line + autotuner choices (stderr); optional env passed through
mkdir -p gpurun_out
tag=${1:-cur}
A3D_AUTOTUNE_VERBOSE=1 timeout 400 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench $tag rc=$?"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_$tag.json
python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])
print(tag, "ms/step", round(d["ms_per_step"],4), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["conv_tensor_tflops"],1), "frac", round(d["roofline"]["conv_tensor_frac_of_burst"],3))
ops=json.load(open(f"gpurun_out/bench_ops_{tag}.json"))["ops"]
for o in ops:
    if o["op"].startswith(("a3d_conv2d","a3d_dense")): print(f'{o["op"]:28s} {o["detail"]:40s} {o["ms"]*1e3:7.1f} us')
PY
grep -- "-> candidate" gpurun_out/bench_$tag.err
