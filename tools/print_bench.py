import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        p = d["phase_ms_per_step"]
        print(f, "gpus", d["n_gpus"], "it/s %.1f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], {k: round(v, 3) for k, v in p.items()})
    except Exception as e:
        print(f, "unreadable:", e); print(open(f).read()[-600:])
