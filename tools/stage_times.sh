python bench.py --no-knn --no-match --no-cpu-baseline "$@" 2>/dev/null | python -c '
import json, sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["roofline"]["stage_ms_left_images"].items()})'
