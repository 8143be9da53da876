"""Prints the few fields of a bench.py JSON line that matter when iterating (ms/step, phases, collectives)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as ex:  # noqa: BLE001
        print(path, "unreadable:", ex)
        continue
    ph = {k: round(v, 3) for k, v in (d.get("phase_ms") or {}).items()}
    print(json.dumps({"file": path, "n_gpus": d.get("n_gpus"), "ms_per_step": round(d.get("ms_per_step", 0), 3),
                      "iterations": d.get("iterations"), "max_residual": d.get("max_residual"), "phase_ms": ph,
                      "collectives": d.get("collectives_per_solve"), "launches": d.get("gpu_launches"),
                      "transport": d.get("transport"), "parity": d.get("parity_check"),
                      "e2e": (d.get("e2e") or {}).get("value"),
                      "other": {k: {"ms": round(v.get("ms_per_step") or 0, 3), "iters": v.get("iterations"),
                                    "it/s": round(v.get("iterations_per_s") or 0, 1),
                                    "ok": (v.get("parity_check") or {}).get("pass"), "err": v.get("error")}
                                for k, v in (d.get("other_configs") or {}).items()}}))
