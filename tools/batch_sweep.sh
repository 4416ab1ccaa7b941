#!/bin/bash
# bench.py at several batch sizes (stereo pairs per launch): device-resident / e2e frames per second and the stage times
for b in ${@:-32 64 128 256 512}; do
  python bench.py --batch "$b" --no-knn --no-match --no-cpu-baseline 2>/dev/null | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("batch", d["config"]["stereo_pairs_per_step_per_gpu"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      {k: round(v, 3) for k, v in d["roofline"]["stage_ms_left_images"].items()})'
done
