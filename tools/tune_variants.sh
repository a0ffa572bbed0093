#!/bin/bash
# Times bench.py's headline workload with every build/tune/libmems_*.so variant in place of the library (builder tool:
# compile-time tuning constants, e.g. -DMEMS_GROUP_LANES / -DMEMS_GROUP_BUDGET of the walk kernels).  Run under gpurun.
cp libmems_b200/libmems_b200.so build/tune/libmems_main.so
for f in build/tune/libmems_*.so; do
  cp "$f" libmems_b200/libmems_b200.so
  python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/tune.json 2> /dev/null || { echo "$f FAILED"; continue; }
  python - "$f" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/tune.json").read().strip().splitlines()[-1])
k = d["kernels"]
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "kernel %.3f" % d["roofline"]["kernel_ms_per_step"],
      " ".join("%s %.3f" % (n, k[n]["ms_per_step"]) for n in (__import__("os").environ.get("TUNE_KERNELS") or "walk_right walk_left long_walk_right long_walk_left giant_walk_right giant_walk_left").split() if n in k),
      "matches", d["matches_per_step"])
PY
done
cp build/tune/libmems_main.so libmems_b200/libmems_b200.so
