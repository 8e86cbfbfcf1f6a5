"""stdin: one JSON line of tools/big_run.py; prints it with every phase list reduced to (count, sum, median)."""
import json, sys
import statistics as st
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        print(line[:300]); continue
    d = json.loads(line)
    d["phases_ms"] = {k: (v[0] if len(v) == 1 else {"n": len(v), "sum": round(sum(v), 2), "median": round(st.median(v), 3)})
                      for k, v in d.get("phases_ms", {}).items()}
    print(json.dumps(d))
