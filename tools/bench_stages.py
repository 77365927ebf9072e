import json,sys
for f in sys.argv[1:]:
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        s=l["config"]["stage_ms"]
        print(f"{f}: value {l['value']:.0f} ms/step {l['ms_per_step']:.3f} vis {s['visibility_ms']:.3f} shade {s['shading_ms']:.3f} e2e {l['e2e']['value']:.0f}")
    except Exception as e:
        print(f, "ERR", e)
