python bench.py --no-knn --no-cpu-baseline 2>/dev/null | python -c '
import json, sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
m = d["match"]; b = d["bow"]
print("SearchByProjection ms", round(m["ms_per_batch"], 4), "local map", {k: round(v, 4) for k, v in m["local_map"].items() if isinstance(v, float) and "ms" in k},
      "bow", {k: round(v, 4) for k, v in b.items() if isinstance(v, float) and "ms" in k})'
